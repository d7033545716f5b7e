// Frame front-end and output head of the recurrence (HBM-bound SIMT kernels).
//   frontend: relu(init_conv(add_coord_channels(frame)))   -- generator.py:166-168, coordconv.py:3-10
//             frames [N, Cf, H, W] fp32 NCHW  ->  features [N, H, W, C] (bf16 or fp32) NHWC
//   head    : 1x1 conv C -> 1 (+bias) on the top layer's h (north_star encoder-forecaster extension)
//             h [N, H, W, C] NHWC -> frames [N, 1, H, W] fp32
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace plc {

template <typename TOut> __device__ __forceinline__ void store8(TOut* dst, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* dst, const float (&v)[8]) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(dst) = o;
}

// Persistent blocks: block (G = C/8, P); threadIdx.x = 8-channel group, threadIdx.y = pixel-quad lane.  Each thread
// computes 8 output channels for 4 CONSECUTIVE x pixels (register blocking: one pair of 16-byte weight loads feeds
// 32 FMAs), and the G lanes of one pixel write one contiguous C*sizeof(TOut) run (coalesced NHWC stores).
// Weights are staged in smem ONCE per block (as [tap*(Cf+2)+ci][C]); the block grid-strides over pixel quads.
// w: init_conv.weight [C, Cf+2, 3, 3] (reference OIHW).
constexpr int kFrontPx = 4;
template <typename TOut>
__global__ void __launch_bounds__(256) frontend_kernel(const float* __restrict__ frames, const float* __restrict__ w,
                                                       const float* __restrict__ bias, TOut* __restrict__ out, int N,
                                                       int Cf, int H, int W, int C, int C_out_stride) {
  extern __shared__ float ws[];  // [(Cf+2)*9][C] + bias[C]
  const int cin = Cf + 2;
  const int nw = cin * 9 * C;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  for (int i = tid; i < nw; i += nthr) {
    const int co = i % C, r = i / C;       // r = tap*cin + ci
    const int tap = r / cin, ci = r % cin;
    ws[i] = w[(static_cast<size_t>(co) * cin + ci) * 9 + tap];
  }
  float* bs = ws + nw;
  for (int i = tid; i < C; i += nthr) bs[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int c0 = threadIdx.x * 8;
  const float inv_h = H > 1 ? 1.f / (H - 1) : 0.f, inv_w = W > 1 ? 1.f / (W - 1) : 0.f;
  const int qpr = (W + kFrontPx - 1) / kFrontPx;                     // quads per image row
  const size_t nquads = static_cast<size_t>(N) * H * qpr;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.y;
  for (size_t q = static_cast<size_t>(blockIdx.x) * blockDim.y + threadIdx.y; q < nquads; q += stride) {
    const int x0 = static_cast<int>(q % qpr) * kFrontPx;
    const size_t t = q / qpr;
    const int y = static_cast<int>(t % H);
    const size_t n = t / H;
    float acc[kFrontPx][8];
#pragma unroll
    for (int px = 0; px < kFrontPx; ++px)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[px][j] = bs[c0 + j];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= H) continue;     // zero padding (applies to the coord channels too)
      for (int ci = 0; ci < cin; ++ci) {
        float v[kFrontPx + 2];             // input values at x0-1 .. x0+kFrontPx
#pragma unroll
        for (int i = 0; i < kFrontPx + 2; ++i) {
          const int xx = x0 - 1 + i;
          float val = 0.f;
          if (xx >= 0 && xx < W) {
            if (ci < Cf) val = __ldg(frames + ((n * Cf + ci) * H + yy) * W + xx);
            else if (ci == Cf) val = yy * inv_h;   // row channel: linspace(0,1,H)   (coordconv.py:7)
            else val = xx * inv_w;                 // col channel: linspace(0,1,W)   (coordconv.py:8)
          }
          v[i] = val;
        }
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4* wr = reinterpret_cast<const float4*>(ws + ((dy * 3 + dx) * cin + ci) * C + c0);
          const float4 w0 = wr[0], w1 = wr[1];
#pragma unroll
          for (int px = 0; px < kFrontPx; ++px) {
            const float a = v[px + dx];
            acc[px][0] = fmaf(a, w0.x, acc[px][0]); acc[px][1] = fmaf(a, w0.y, acc[px][1]);
            acc[px][2] = fmaf(a, w0.z, acc[px][2]); acc[px][3] = fmaf(a, w0.w, acc[px][3]);
            acc[px][4] = fmaf(a, w1.x, acc[px][4]); acc[px][5] = fmaf(a, w1.y, acc[px][5]);
            acc[px][6] = fmaf(a, w1.z, acc[px][6]); acc[px][7] = fmaf(a, w1.w, acc[px][7]);
          }
        }
      }
    }
    const size_t pix0 = (n * H + y) * W + x0;
#pragma unroll
    for (int px = 0; px < kFrontPx; ++px) {
      if (x0 + px < W) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[px][j] = fmaxf(acc[px][j], 0.f);   // F.relu (generator.py:168)
        store8<TOut>(out + (pix0 + px) * C_out_stride + c0, acc[px]);
      }
    }
  }
}

// G = C/8 lanes per pixel (G a power of two <= 32): each lane reads 16 B, partial dot products are
// combined with shuffles -> one coalesced C*2-byte read per pixel.
template <int G>
__global__ void head_kernel_bf16_coalesced(const __nv_bfloat16* __restrict__ h, const float* __restrict__ w,
                                           const float* __restrict__ bias, float* __restrict__ out, size_t npix) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t p = gid / G;
  const int g = gid % G;
  float acc = 0.f;
  if (p < npix) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(h + p * (G * 8) + g * 8));
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&u[j]);
      acc = fmaf(__low2float(t), __ldg(w + g * 8 + 2 * j), acc);
      acc = fmaf(__high2float(t), __ldg(w + g * 8 + 2 * j + 1), acc);
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (p < npix && g == 0) out[p] = acc + (bias ? bias[0] : 0.f);
}

template <typename TIn>
__global__ void head_kernel(const TIn* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
                            float* __restrict__ out, size_t npix, int C) {
  const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (p >= npix) return;
  float acc = bias ? bias[0] : 0.f;
  const TIn* src = h + p * C;
  for (int c = 0; c < C; c += 8) {
    if constexpr (sizeof(TIn) == 2) {
      uint4 v = *reinterpret_cast<const uint4*>(src + c);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&u[j]);
        acc = fmaf(__low2float(t), __ldg(w + c + 2 * j), acc);
        acc = fmaf(__high2float(t), __ldg(w + c + 2 * j + 1), acc);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(static_cast<float>(src[c + j]), __ldg(w + c + j), acc);
    }
  }
  out[p] = acc;
}

// frames [B, T, Cf, H, W] fp32 (reference layout of rain_lr) -> [T*B, H, W, Cp] bf16, T-major, with the two coordinate
// planes of coordconv.py:3-10 appended (row = y/(H-1), col = x/(W-1)) and zero padding up to Cp channels (Cp % 8 == 0).
__global__ void __launch_bounds__(256) frames_to_nhwc_kernel(const float* __restrict__ frames,
                                                             __nv_bfloat16* __restrict__ out, int B, int T, int Cf,
                                                             int H, int W, int Cp) {
  const size_t npix = static_cast<size_t>(B) * T * H * W;
  const float inv_h = H > 1 ? 1.f / (H - 1) : 0.f, inv_w = W > 1 ? 1.f / (W - 1) : 0.f;
  for (size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; p < npix;
       p += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(p % W);
    size_t r = p / W;
    const int y = static_cast<int>(r % H);
    r /= H;
    const int b = static_cast<int>(r % B);
    const int t = static_cast<int>(r / B);
    const float* src = frames + ((static_cast<size_t>(b) * T + t) * Cf) * H * W + static_cast<size_t>(y) * W + x;
    for (int c0 = 0; c0 < Cp; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        v[j] = c < Cf ? __ldg(src + static_cast<size_t>(c) * H * W) : (c == Cf ? y * inv_h : (c == Cf + 1 ? x * inv_w : 0.f));
      }
      store8<__nv_bfloat16>(out + p * Cp + c0, v);
    }
  }
}

// Backward of the 1x1 head: y[p] = sum_c h[p][c] * w[c] + b
//   dh[p][c] = dy[p] * w[c]  (bf16, coalesced: G = C/8 lanes per pixel)
//   dw[c]   += sum_p dy[p] * h[p][c],   db += sum_p dy[p]       (block partial sums -> atomics)
template <int G>
__global__ void __launch_bounds__(256) head_bwd_kernel_bf16(const __nv_bfloat16* __restrict__ h,
                                                            const float* __restrict__ w, const float* __restrict__ dy,
                                                            __nv_bfloat16* __restrict__ dh, float* __restrict__ dw,
                                                            float* __restrict__ db, size_t npix) {
  __shared__ float red[256 * 8];
  const int g = threadIdx.x % G;
  float wv[8], acc[8], bsum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { wv[j] = __ldg(w + g * 8 + j); acc[j] = 0.f; }
  const size_t stride = static_cast<size_t>(gridDim.x) * (blockDim.x / G);
  for (size_t p = static_cast<size_t>(blockIdx.x) * (blockDim.x / G) + threadIdx.x / G; p < npix; p += stride) {
    const float d = __ldg(dy + p);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(h + p * (G * 8) + g * 8));
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&u[j]);
      acc[2 * j] = fmaf(d, __low2float(t), acc[2 * j]);
      acc[2 * j + 1] = fmaf(d, __high2float(t), acc[2 * j + 1]);
      const __nv_bfloat162 r = __floats2bfloat162_rn(d * wv[2 * j], d * wv[2 * j + 1]);
      o[j] = *reinterpret_cast<const uint32_t*>(&r);
    }
    *reinterpret_cast<uint4*>(dh + p * (G * 8) + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    if (g == 0) bsum += d;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  // threads 0 .. C-1 reduce channel c = t over the blockDim/G pixel lanes
  const int C = G * 8;
  if (threadIdx.x < C) {
    const int gg = threadIdx.x / 8, jj = threadIdx.x % 8;
    float s = 0.f;
    for (int l = 0; l < blockDim.x / G; ++l) s += red[(l * G + gg) * 8 + jj];
    atomicAdd(dw + threadIdx.x, s);
  }
  if (db) {
    for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    if ((threadIdx.x & 31) == 0 && bsum != 0.f) atomicAdd(db, bsum);
  }
}

}  // namespace plc
