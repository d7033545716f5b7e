// First layer of the clip discriminator: a 3x3 (optionally stride-2) convolution straight from the fp32 frames
// [N, Cf, H, W] (Cf = 1..4 rain channels) to NHWC bf16 features, bias + LeakyReLU fused -- and its backward.
//
// With Cf = 1 the contraction is 9 multiply-adds per output value: on the implicit-GEMM core the layer was all padding
// (input zero-padded to 8 bf16 channels = 4x the frame bytes, K padded 9 -> 192, N = 32 of a 128-wide MMA; wgrad
// 18 TFLOP/s "executed").  Its roofline is HBM: frames in (4 B / input pixel) + features out (2 * Cout B / output
// pixel).  Three SIMT kernels, all reading the frames where they lie (no layout / pad / cast pass):
//   frameconv_fwd_kernel   : 8 channels x 4 consecutive output pixels per thread, coalesced NHWC stores
//   frameconv_wgrad_kernel : dW [Cout, Cf, 3, 3] and db: 72 register accumulators per thread (9 taps x 8 channels),
//                            warp-shuffle + smem block reduction, one atomicAdd per output per block
//   frameconv_dgrad_kernel : dframes (fp32, the generator's adversarial gradient): per output row, the channel
//                            contraction once per output pixel into a shared-memory ring, then a gather per input pixel
// The activation derivative is folded into both backward kernels (dZ = dY * (Y > 0 ? 1 : slope) on the fly), so the
// layer needs no gradient-mask pass either.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "frame_io.cuh"

namespace plc {

struct FrameConvParams {
  const float* frames;        // [N, Cf, H, W] fp32
  const float* w;             // [Cout, Cf, 3, 3] fp32 (torch OIHW)
  const float* bias;          // [Cout] or nullptr
  int N, Cf, H, W, Ho, Wo, Cout, stride;
  int act;                    // 0 none, 1 ReLU, 2 LeakyReLU(slope)
  float slope;
  __nv_bfloat16* out;         // fwd: [N, Ho, Wo, Cout]
  const __nv_bfloat16* y;     // bwd: forward output (nullptr when act == 0)
  const __nv_bfloat16* dy;    // bwd: [N, Ho, Wo, Cout]
  float* dframes;             // bwd: [N, Cf, H, W] fp32 (dgrad kernel)
  float* dW;                  // bwd: [Cout, Cf, 3, 3] fp32, accumulated
  float* db;                  // bwd: [Cout] fp32, accumulated, or nullptr
};

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&u[j]);
    f[2 * j] = __low2float(t);
    f[2 * j + 1] = __high2float(t);
  }
}

// dZ of 8 channels: dY masked by the activation derivative at Y (act: 0 none, 1 ReLU, 2 LeakyReLU(slope)).  The mask is
// applied on the packed bf16 pairs (factor = 1 or slope, exact in the select; the product rounds to bf16 exactly like
// the dZ tensor plc_convnd_grad_mask stores for the other layers).
__device__ __forceinline__ void load_dz8(const FrameConvParams& p, size_t off, float (&dz)[8]) {
  uint4 d = __ldg(reinterpret_cast<const uint4*>(p.dy + off));
  if (p.act != 0) {
    const uint4 yv = __ldg(reinterpret_cast<const uint4*>(p.y + off));
    const __nv_bfloat162 one = __float2bfloat162_rn(1.f), neg = __float2bfloat162_rn(p.act == 2 ? p.slope : 0.f);
    const __nv_bfloat162 zero = __float2bfloat162_rn(0.f);
    uint32_t* dw = reinterpret_cast<uint32_t*>(&d);
    const uint32_t* yw = reinterpret_cast<const uint32_t*>(&yv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 m = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&yw[j]), zero);   // 1.0 / 0.0 per half
      const __nv_bfloat162 f = __hfma2(m, __hsub2(one, neg), neg);                                // 1 or slope
      const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&dw[j]), f);
      dw[j] = *reinterpret_cast<const uint32_t*>(&r);
    }
  }
  unpack8(d, dz);
}

// Thread layout of all three kernels: 256-thread blocks; G = Cout/8 consecutive lanes own the 8-channel groups of one
// pixel (their 16-byte accesses form one contiguous Cout*2-byte run), the 32/G "pixel lanes" of a warp walk along ONE
// image row, and warps grid-stride over rows -- the (image, row) decode happens once per row and nothing in the inner
// loops divides (a first version that decoded a flat 64-bit pixel index per iteration spent most of its time in the
// division routine: 419 / 1012 / 910 us for fwd / wgrad / dgrad on 2560 frames of 128x128).
constexpr int kFcPx = 4;

__global__ void __launch_bounds__(256) frameconv_fwd_kernel(const FrameConvParams p) {
  extern __shared__ float fc_ws[];   // [9 * Cf][Cout] + bias [Cout]
  const int C = p.Cout, cin = p.Cf, s = p.stride;
  const int nw = cin * 9 * C;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) {
    const int co = i % C, r = i / C, tap = r / cin, ci = r % cin;
    fc_ws[i] = p.w[(static_cast<size_t>(co) * cin + ci) * 9 + tap];
  }
  float* bs = fc_ws + nw;
  for (int i = threadIdx.x; i < C; i += blockDim.x) bs[i] = p.bias ? p.bias[i] : 0.f;
  __syncthreads();
  const int G = C >> 3, PL = 32 / G, lane = threadIdx.x & 31;
  const int c0 = (lane % G) * 8, pl = lane / G;
  const int qpr = (p.Wo + kFcPx - 1) / kFcPx;      // quads of 4 consecutive output pixels per row
  const unsigned rows = static_cast<unsigned>(p.N) * p.Ho;
  const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
  const float neg = p.act == 2 ? p.slope : 0.f;
  constexpr int kSpan = (kFcPx - 1) * 2 + 3;      // input columns a quad touches at stride 2 (stride 1: kFcPx + 2)
  for (unsigned row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += nwarps) {
    const unsigned n = row / p.Ho;
    const int yo = static_cast<int>(row - n * p.Ho);
    for (int qx = pl; qx < qpr; qx += PL) {
      const int xo0 = qx * kFcPx;
      float acc[kFcPx][8];
#pragma unroll
      for (int px = 0; px < kFcPx; ++px)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[px][j] = bs[c0 + j];
      const int xi0 = xo0 * s - 1;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = yo * s + dy - 1;
        if (yy < 0 || yy >= p.H) continue;           // zero padding
        for (int ci = 0; ci < cin; ++ci) {
          const float* rowp = p.frames + ((static_cast<size_t>(n) * cin + ci) * p.H + yy) * p.W;
          float v[kSpan];
#pragma unroll
          for (int i = 0; i < kSpan; ++i) {
            const int xx = xi0 + i;
            v[i] = (xx >= 0 && xx < p.W && i < (kFcPx - 1) * s + 3) ? __ldg(rowp + xx) : 0.f;
          }
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4* wr = reinterpret_cast<const float4*>(fc_ws + ((dy * 3 + dx) * cin + ci) * C + c0);
            const float4 w0 = wr[0], w1 = wr[1];
#pragma unroll
            for (int px = 0; px < kFcPx; ++px) {
              const float a = s == 2 ? v[2 * px + dx] : v[px + dx];
              acc[px][0] = fmaf(a, w0.x, acc[px][0]); acc[px][1] = fmaf(a, w0.y, acc[px][1]);
              acc[px][2] = fmaf(a, w0.z, acc[px][2]); acc[px][3] = fmaf(a, w0.w, acc[px][3]);
              acc[px][4] = fmaf(a, w1.x, acc[px][4]); acc[px][5] = fmaf(a, w1.y, acc[px][5]);
              acc[px][6] = fmaf(a, w1.z, acc[px][6]); acc[px][7] = fmaf(a, w1.w, acc[px][7]);
            }
          }
        }
      }
      const size_t pix0 = static_cast<size_t>(row) * p.Wo + xo0;
#pragma unroll
      for (int px = 0; px < kFcPx; ++px) {
        if (xo0 + px < p.Wo) {
          if (p.act != 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[px][j] = acc[px][j] > 0.f ? acc[px][j] : neg * acc[px][j];
          }
          store8<__nv_bfloat16>(p.out + (pix0 + px) * C + c0, acc[px]);
        }
      }
    }
  }
}

// grid (blocks, Cf): thread = 8 channels of one output pixel at a time, 9 taps x 8 channels of partial dW for input
// channel blockIdx.y in registers (+ 8 of db when blockIdx.y == 0).
__global__ void __launch_bounds__(256, 2) frameconv_wgrad_kernel(const FrameConvParams p) {
  __shared__ float red[8][80];         // per warp: 72 dW + 8 db partials of ONE channel group at a time
  const int C = p.Cout, s = p.stride, ci = blockIdx.y;
  const int G = C >> 3, PL = 32 / G, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (lane % G) * 8, pl = lane / G;
  const unsigned rows = static_cast<unsigned>(p.N) * p.Ho;
  const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
  float acc[9][8], dbv[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) dbv[j] = 0.f;
  for (unsigned row = blockIdx.x * (blockDim.x >> 5) + warp; row < rows; row += nwarps) {
    const unsigned n = row / p.Ho;
    const int yo = static_cast<int>(row - n * p.Ho);
    const float* img = p.frames + (static_cast<size_t>(n) * p.Cf + ci) * p.H * p.W;
    const float* r0 = img + static_cast<ptrdiff_t>(yo * s - 1) * p.W;      // input rows yo*s - 1 .. yo*s + 1
    const bool v0 = yo * s - 1 >= 0, v2 = yo * s + 1 < p.H;              // (row yo*s itself always exists)
    for (int xo = pl; xo < p.Wo; xo += PL) {
      float dz[8];
      load_dz8(p, (static_cast<size_t>(row) * p.Wo + xo) * C + c0, dz);
#pragma unroll
      for (int j = 0; j < 8; ++j) dbv[j] += dz[j];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = xo * s + dx - 1;
        const bool vx = xx >= 0 && xx < p.W;
        const float a0 = (vx && v0) ? __ldg(r0 + xx) : 0.f;
        const float a1 = vx ? __ldg(r0 + p.W + xx) : 0.f;
        const float a2 = (vx && v2) ? __ldg(r0 + 2 * p.W + xx) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[dx][j] = fmaf(a0, dz[j], acc[dx][j]);
          acc[3 + dx][j] = fmaf(a1, dz[j], acc[3 + dx][j]);
          acc[6 + dx][j] = fmaf(a2, dz[j], acc[6 + dx][j]);
        }
      }
    }
  }
  // reduce over the pixel lanes: xor-shuffles over lane bits >= log2(G) stay inside a channel group
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      for (int o = 16; o >= G; o >>= 1) acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], o);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    for (int o = 16; o >= G; o >>= 1) dbv[j] += __shfl_xor_sync(0xffffffffu, dbv[j], o);
  // lanes 0..G-1 of every warp now hold the warp totals of channel groups 0..G-1
  const int nw = blockDim.x >> 5;
  for (int g = 0; g < G; ++g) {
    __syncthreads();
    if (lane == g) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][t * 8 + j] = acc[t][j];
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][72 + j] = dbv[j];
    }
    __syncthreads();
    if (threadIdx.x < 80) {
      float sum = 0.f;
      for (int w = 0; w < nw; ++w) sum += red[w][threadIdx.x];
      const int co = g * 8 + (threadIdx.x & 7);
      if (threadIdx.x < 72) atomicAdd(p.dW + (static_cast<size_t>(co) * p.Cf + ci) * 9 + (threadIdx.x >> 3), sum);
      else if (p.db && ci == 0) atomicAdd(p.db + co, sum);
    }
  }
}

// dframes in two stages per output row, one block per frame:
//   stage 1  t[xo][tap, ci] = sum_co dZ[yo, xo, co] * w[co, ci, tap]: G lanes per output pixel contract their 8 channels
//            (coalesced 16-byte loads, dZ formed once per output pixel) and meet in a G-lane shuffle reduction; the row
//            of t lands in a shared-memory ring of R rows (R = 2 at stride 2, 3 at stride 1);
//   stage 2  every input pixel of block-row Yb = yo - 1 sums the <= R*R ring entries whose tap reaches it
//            (ky = iy + 1 - (row offset) * S resolved per thread) and stores its Cf values, coalesced along x.
// The (output pixel, tap) products are shared by all input pixels they feed, so the 32-channel contraction runs once
// per OUTPUT pixel -- a per-input-pixel gather repeated it up to four times and was instruction-bound.
template <int S, int CF>
__global__ void __launch_bounds__(256) frameconv_dgrad_kernel(const FrameConvParams p) {
  extern __shared__ float fc_sm[];   // weights [9][CF][Cout] | ring t[R][Wo][9 * CF]
  constexpr int R = S == 2 ? 2 : 3, O0 = S == 2 ? 0 : -1, NT = 9 * CF;
  const int C = p.Cout;
  float* ws = fc_sm;
  float* ring = fc_sm + 9 * CF * C;
  for (int i = threadIdx.x; i < 9 * CF * C; i += blockDim.x) {
    const int co = i % C, r = i / C, ci = r % CF, tap = r / CF;
    ws[i] = p.w[(static_cast<size_t>(co) * CF + ci) * 9 + tap];
  }
  const int G = C >> 3, g = threadIdx.x % G, c0 = g * 8;
  const int PB = blockDim.x / G;               // output pixels per pass of stage 1
  const int Hb = (p.H + S - 1) / S;
  for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
    for (int yo = 0; yo <= p.Ho; ++yo) {
      __syncthreads();                         // ring row yo % R is free again (and, first time, the weights are staged)
      if (yo < p.Ho) {
        float* trow = ring + (yo % R) * p.Wo * NT;
        for (int xo0 = 0; xo0 < p.Wo; xo0 += PB) {         // all lanes stay in the loop: the shuffles need full warps
          const int xo = xo0 + threadIdx.x / G;
          float t[NT];
#pragma unroll
          for (int k = 0; k < NT; ++k) t[k] = 0.f;
          if (xo < p.Wo) {
            float dz[8];
            load_dz8(p, ((static_cast<size_t>(n) * p.Ho + yo) * p.Wo + xo) * C + c0, dz);
#pragma unroll
            for (int k = 0; k < NT; ++k) {
              const float4 w0 = *reinterpret_cast<const float4*>(ws + k * C + c0);
              const float4 w1 = *reinterpret_cast<const float4*>(ws + k * C + c0 + 4);
              t[k] = dz[0] * w0.x + dz[1] * w0.y + dz[2] * w0.z + dz[3] * w0.w + dz[4] * w1.x + dz[5] * w1.y +
                     dz[6] * w1.z + dz[7] * w1.w;
            }
          }
#pragma unroll
          for (int k = 0; k < NT; ++k)
            for (int o = 1; o < G; o <<= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
          if (xo < p.Wo) {
#pragma unroll
            for (int k = 0; k < NT; ++k)
              if (k % G == g) trow[xo * NT + k] = t[k];     // lane g of the group stores entries g, g + G, ...
          }
        }
      }
      __syncthreads();
      const int Yb = yo - 1;                   // block-row whose last contributing output row is yo
      if (Yb < 0 || Yb >= Hb) continue;
      for (int i = threadIdx.x; i < S * p.W; i += blockDim.x) {
        const int iy = i / p.W, x = i - iy * p.W;
        const int y = Yb * S + iy;
        if (y >= p.H) continue;
        const int Xb = x / S, ix = x - Xb * S;
        float acc[CF];
#pragma unroll
        for (int ci = 0; ci < CF; ++ci) acc[ci] = 0.f;
#pragma unroll
        for (int ry = 0; ry < R; ++ry) {
          const int yr = Yb + O0 + ry, ky = iy + 1 - (O0 + ry) * S;
          if (yr < 0 || yr >= p.Ho || ky < 0 || ky > 2) continue;
          const float* trow = ring + (yr % R) * p.Wo * NT;
#pragma unroll
          for (int rx = 0; rx < R; ++rx) {
            const int xr = Xb + O0 + rx, kx = ix + 1 - (O0 + rx) * S;
            if (xr < 0 || xr >= p.Wo || kx < 0 || kx > 2) continue;
#pragma unroll
            for (int ci = 0; ci < CF; ++ci) acc[ci] += trow[xr * NT + (ky * 3 + kx) * CF + ci];
          }
        }
#pragma unroll
        for (int ci = 0; ci < CF; ++ci) p.dframes[((static_cast<size_t>(n) * CF + ci) * p.H + y) * p.W + x] = acc[ci];
      }
    }
  }
}

}  // namespace plc
