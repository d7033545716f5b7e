// Weight-gradient reduction for the bf16 mode:
//   dW[o][i][ky][kx] += sum_pixels dZ[pix][o] * cat(x,h)[pix + (ky-pad, kx-pad)][i]      
// accumulated in fp32 straight into the reference OIHW layout (`conv.weight.grad`).
#pragma once
#include "../../include/plc.h"
#include "conv_simt.cuh"

namespace plc {

inline int launch_wgrad_tc(const PlcCellDesc* d, const void* x, const void* h_prev, const void* dz, float* dW,
                           float* db, int num_sms, cudaStream_t st) {
  const int M = d->B * d->H * d->W;
  WgradParams w;
  memset(&w, 0, sizeof(w));
  w.B = d->B; w.H = d->H; w.W = d->W; w.M = M;
  w.ksize = d->k; w.pad = d->k / 2;
  w.C0 = d->Cin; w.C1 = d->Ch; w.K = d->k * d->k * (d->Cin + d->Ch); w.N = 4 * d->Ch;
  w.src0 = x; w.src1 = h_prev; w.dz = dz; w.dW = dW; w.db = db;
  const int tiles = ((w.N + SBN - 1) / SBN) * ((w.K + SBM - 1) / SBM);
  int splits = (num_sms * 4 + tiles - 1) / tiles;
  if (splits < 1) splits = 1;
  int ppb = (((M + splits - 1) / splits + SBK - 1) / SBK) * SBK;
  if (ppb < SBK) ppb = SBK;
  w.pix_per_block = ppb;
  dim3 grid((w.N + SBN - 1) / SBN, (w.K + SBM - 1) / SBM, (M + ppb - 1) / ppb);
  wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(w);
  return cudaGetLastError() == cudaSuccess ? PLC_OK : PLC_ERR_CUDA;
}

}  // namespace plc
