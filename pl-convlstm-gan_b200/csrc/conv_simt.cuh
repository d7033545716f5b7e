// fp32 SIMT kernels: the "fp32 validation mode" of the ConvLSTM cell (north_star: <= 1e-5 vs the
// reference) and the weight-gradient reduction.  Same algorithm as the tensor-core path (implicit
// GEMM over (src, tap, channel), fused gate epilogue) with FFMA accumulation and exact expf/tanhf.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace plc {

enum : int { SEPI_LSTM_FWD = 0, SEPI_LSTM_BWD_GATES = 1, SEPI_PLAIN = 2 };

struct ConvSimtParams {
  int B, H, W, M;            // M = B*H*W
  int ksize, pad;
  int C0, C1;                // channels of source 0 (x) and source 1 (h)
  int K;                     // ksize^2 * (C0 + C1)
  int N;                     // output columns (4Ch for LSTM, Cin+Ch for PLAIN)
  int Ch, Cin;
  const float* src0;         // [M, C0] fp32 NHWC
  const float* src1;         // [M, C1]
  const float* w;            // packed [K][N]; LSTM: column n' = ch*4 + gate
  const float* bias;         // [4Ch] reference order or nullptr
  const float* c_prev;
  float* c_out;
  float* h_out;
  float* gates_out;          // [M, 4Ch] reference gate order, or nullptr
  const float* dh;
  const float* dh2;
  const float* dc_next;
  float* dc_prev;
  float* dz;                 // [M, 4Ch] reference gate order
  float* out0;               // PLAIN: [M, Cin]
  float* out1;               // PLAIN: [M, N - Cin]
};

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int SBM = 64, SBN = 64, SBK = 16;

// K index -> (source, tap, channel); order: source 0 taps/channels first, then source 1.
__device__ __forceinline__ void decode_k(const ConvSimtParams& p, int kidx, int& src, int& tap, int& c) {
  const int k0 = p.ksize * p.ksize * p.C0;
  if (kidx < k0) {
    src = 0; tap = kidx / p.C0; c = kidx - tap * p.C0;
  } else {
    const int r = kidx - k0;
    src = 1; tap = r / p.C1; c = r - tap * p.C1;
  }
}

template <int EPI>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tm = tid >> 4, tn = tid & 15;
  const int m0 = blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;

  // A-load assignment: pixel lm, K rows lk + 4*r
  const int lm = tid & 63, lk = tid >> 6;
  const int gm = m0 + lm;
  int pb = 0, py = 0, px = 0;
  const bool m_ok = gm < p.M;
  if (m_ok) {
    px = gm % p.W;
    const int t = gm / p.W;
    py = t % p.H;
    pb = t / p.H;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += SBK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kr = lk + 4 * r;
      const int kidx = k0 + kr;
      float v = 0.f;
      if (m_ok && kidx < p.K) {
        int src, tap, c;
        decode_k(p, kidx, src, tap, c);
        const int yy = py + tap / p.ksize - p.pad;
        const int xx = px + tap % p.ksize - p.pad;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
          const size_t pix = (static_cast<size_t>(pb) * p.H + yy) * p.W + xx;
          v = src ? p.src1[pix * p.C1 + c] : p.src0[pix * p.C0 + c];
        }
      }
      As[kr][lm] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int kr = lk + 4 * r;
      const int kidx = k0 + kr;
      const int n = n0 + lm;
      Bs[kr][lm] = (kidx < p.K && n < p.N) ? p.w[static_cast<size_t>(kidx) * p.N + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tm * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tn * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm * 4 + i;
    if (m >= p.M) continue;
    if constexpr (EPI == SEPI_PLAIN) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tn * 4 + j;
        if (n >= p.N) continue;
        if (n < p.Cin) {
          if (p.out0) p.out0[static_cast<size_t>(m) * p.Cin + n] = acc[i][j];
        } else {
          if (p.out1) p.out1[static_cast<size_t>(m) * (p.N - p.Cin) + (n - p.Cin)] = acc[i][j];
        }
      }
    } else {
      const int ch = (n0 + tn * 4) >> 2;  // packed column n' = ch*4 + gate
      if (ch >= p.Ch) continue;
      const size_t off = static_cast<size_t>(m) * p.Ch + ch;
      const float bi = p.bias ? p.bias[0 * p.Ch + ch] : 0.f;
      const float bf = p.bias ? p.bias[1 * p.Ch + ch] : 0.f;
      const float bo = p.bias ? p.bias[2 * p.Ch + ch] : 0.f;
      const float bg = p.bias ? p.bias[3 * p.Ch + ch] : 0.f;
      const float ig = sigmoid_exact(acc[i][0] + bi);   // convlstm.py:21
      const float fg = sigmoid_exact(acc[i][1] + bf);   // convlstm.py:22
      const float og = sigmoid_exact(acc[i][2] + bo);   // convlstm.py:23
      const float gt = tanhf(acc[i][3] + bg);           // convlstm.py:24
      const float cp = p.c_prev[off];
      const float c2 = fg * cp + ig * gt;               // convlstm.py:26
      const float tc = tanhf(c2);
      if constexpr (EPI == SEPI_LSTM_FWD) {
        p.c_out[off] = c2;
        p.h_out[off] = og * tc;                         // convlstm.py:27
        if (p.gates_out) {
          float* g = p.gates_out + static_cast<size_t>(m) * 4 * p.Ch + ch;
          g[0 * p.Ch] = ig; g[1 * p.Ch] = fg; g[2 * p.Ch] = og; g[3 * p.Ch] = gt;
        }
      } else {
        float dh_ = p.dh[off];
        if (p.dh2) dh_ += p.dh2[off];
        const float dcn = p.dc_next ? p.dc_next[off] : 0.f;
        const float dc = dcn + dh_ * og * (1.f - tc * tc);
        p.dc_prev[off] = dc * fg;
        float* z = p.dz + static_cast<size_t>(m) * 4 * p.Ch + ch;
        z[0 * p.Ch] = dc * gt * ig * (1.f - ig);
        z[1 * p.Ch] = dc * cp * fg * (1.f - fg);
        z[2 * p.Ch] = dh_ * tc * og * (1.f - og);
        z[3 * p.Ch] = dc * ig * (1.f - gt * gt);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ wgrad
// dW[o][i][ky][kx] += sum_pixels dZ[pix][o] * comb[pix + (ky-pad, kx-pad)][i]   (reference OIHW layout)
// Output tile 64 (o) x 64 (k index = (src, tap, c)); each block reduces a contiguous pixel chunk
// and atomically adds its partial tile.  TIn = float (validation mode) or __nv_bfloat16.
struct WgradParams {
  int B, H, W, M;
  int ksize, pad;
  int C0, C1, K;          // K = ksize^2*(C0+C1)
  int N;                  // 4Ch
  int pix_per_block;
  const void* src0;       // [M, C0]
  const void* src1;       // [M, C1]
  const void* dz;         // [M, N] reference gate order
  float* dW;              // [N][C0+C1][k][k]
  float* db;              // [N] or nullptr
};

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

template <typename TIn>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradParams p) {
  __shared__ float Zs[SBK][SBN + 4];  // [pixel][o]
  __shared__ float As[SBK][SBM + 4];  // [pixel][k]
  const int tid = threadIdx.x;
  const int to = tid >> 4, tk = tid & 15;
  const int o0 = blockIdx.x * SBN;
  const int kbase = blockIdx.y * SBM;
  const int m_begin = blockIdx.z * p.pix_per_block;
  const int m_end = min(p.M, m_begin + p.pix_per_block);
  const TIn* s0 = reinterpret_cast<const TIn*>(p.src0);
  const TIn* s1 = reinterpret_cast<const TIn*>(p.src1);
  const TIn* dz = reinterpret_cast<const TIn*>(p.dz);

  // this thread loads column lc of both tiles for pixel rows lr + 4*r
  const int lc = tid & 63, lr = tid >> 6;
  const int kidx = kbase + lc;
  int ksrc = 0, ktap = 0, kc = 0;
  const bool k_ok = kidx < p.K;
  if (k_ok) {
    const int k0n = p.ksize * p.ksize * p.C0;
    if (kidx < k0n) { ksrc = 0; ktap = kidx / p.C0; kc = kidx - ktap * p.C0; }
    else { const int r = kidx - k0n; ksrc = 1; ktap = r / p.C1; kc = r - ktap * p.C1; }
  }
  const int kdy = ktap / p.ksize - p.pad, kdx = ktap % p.ksize - p.pad;
  const bool o_ok = (o0 + lc) < p.N;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;  // db partial: threads with blockIdx.y == 0 && lr == 0.. handled below

  for (int mb = m_begin; mb < m_end; mb += SBK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int pr = lr + 4 * r;
      const int m = mb + pr;
      float zv = 0.f, av = 0.f;
      if (m < m_end) {
        if (o_ok) zv = ld_as_float<TIn>(dz + static_cast<size_t>(m) * p.N + o0 + lc);
        if (k_ok) {
          const int x = m % p.W;
          const int t = m / p.W;
          const int y = t % p.H;
          const int b = t / p.H;
          const int yy = y + kdy, xx = x + kdx;
          if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
            const size_t pix = (static_cast<size_t>(b) * p.H + yy) * p.W + xx;
            av = ksrc ? ld_as_float<TIn>(s1 + pix * p.C1 + kc) : ld_as_float<TIn>(s0 + pix * p.C0 + kc);
          }
        }
      }
      Zs[pr][lc] = zv;
      As[pr][lc] = av;
      bsum += zv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < SBK; ++pp) {
      float a[4], z[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = Zs[pp][to * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] = As[pp][tk * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(z[i], a[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int ctot = p.C0 + p.C1;
  const int kk = p.ksize * p.ksize;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kj = kbase + tk * 4 + j;
    if (kj >= p.K) continue;
    int src, tap, c;
    const int k0n = kk * p.C0;
    if (kj < k0n) { src = 0; tap = kj / p.C0; c = kj - tap * p.C0; }
    else { const int r = kj - k0n; src = 1; tap = r / p.C1; c = r - tap * p.C1; }
    const int ic = src ? p.C0 + c : c;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = o0 + to * 4 + i;
      if (o >= p.N) continue;
      atomicAdd(&p.dW[(static_cast<size_t>(o) * ctot + ic) * kk + tap], acc[i][j]);
    }
  }
  if (p.db && blockIdx.y == 0 && o_ok) {
    // 4 threads (lr = 0..3) share column lc
    atomicAdd(&p.db[o0 + lc], bsum);
  }
}

// ------------------------------------------------------------------------------------ packing
// fp32 forward image: Wf[k=(src,tap,c)][n' = ch*4 + gate] = w[gate*Ch + ch][i][ky][kx]
__global__ void pack_w_f32_fwd_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Ch,
                                      int ksize) {
  const int kk = ksize * ksize, ctot = Cin + Ch, K = kk * ctot, N = 4 * Ch;
  const size_t total = static_cast<size_t>(K) * N;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = idx % N, kidx = idx / N;
    const int ch = n >> 2, gate = n & 3;
    int src, tap, c;
    const int k0n = kk * Cin;
    if (kidx < k0n) { src = 0; tap = kidx / Cin; c = kidx - tap * Cin; }
    else { const int r = kidx - k0n; src = 1; tap = r / Ch; c = r - tap * Ch; }
    const int ic = src ? Cin + c : c;
    out[idx] = w[(static_cast<size_t>(gate * Ch + ch) * ctot + ic) * kk + tap];
  }
}
// fp32 dgrad image: source = dZ (4Ch channels, reference order), outputs c in [0, Cin+Ch)
//   Wd[k=(tap', n)][c] = w[n][c][k-1-ty'][k-1-tx']
__global__ void pack_w_f32_dgrad_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Ch,
                                        int ksize) {
  const int kk = ksize * ksize, ctot = Cin + Ch, N4 = 4 * Ch;
  const size_t total = static_cast<size_t>(kk) * N4 * ctot;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = idx % ctot;
    const int kidx = idx / ctot;
    const int tap = kidx / N4, n = kidx - tap * N4;
    const int fy = ksize - 1 - tap / ksize, fx = ksize - 1 - tap % ksize;
    out[idx] = w[(static_cast<size_t>(n) * ctot + c) * kk + fy * ksize + fx];
  }
}

}  // namespace plc
