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

// grid: (ceil(W/P), H, N); block: (G = C/8, P).  threadIdx.x = 8-channel group, threadIdx.y = pixel, so the
// G lanes of one pixel write one contiguous C*sizeof(TOut) run (coalesced NHWC stores).
// w: init_conv.weight [C, Cf+2, 3, 3] (reference OIHW), staged in smem as [tap*(Cf+2)+ci][C].
template <typename TOut>
__global__ void frontend_kernel(const float* __restrict__ frames, const float* __restrict__ w,
                                const float* __restrict__ bias, TOut* __restrict__ out, int Cf, int H, int W, int C,
                                int C_out_stride) {
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
  const int x = blockIdx.x * blockDim.y + threadIdx.y, y = blockIdx.y, n = blockIdx.z;
  if (x >= W) return;
  const int c0 = threadIdx.x * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = bs[c0 + j];
  const float inv_h = H > 1 ? 1.f / (H - 1) : 0.f, inv_w = W > 1 ? 1.f / (W - 1) : 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;  // zero padding applies to the coord channels too
    for (int ci = 0; ci < cin; ++ci) {
      float v;
      if (ci < Cf) v = __ldg(frames + ((static_cast<size_t>(n) * Cf + ci) * H + yy) * W + xx);
      else if (ci == Cf) v = yy * inv_h;      // row channel: linspace(0,1,H)   (coordconv.py:7)
      else v = xx * inv_w;                    // col channel: linspace(0,1,W)   (coordconv.py:8)
      const float4* wr = reinterpret_cast<const float4*>(ws + (tap * cin + ci) * C + c0);
      const float4 w0 = wr[0], w1 = wr[1];
      acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
      acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
      acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
      acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);   // F.relu (generator.py:168)
  store8<TOut>(out + ((static_cast<size_t>(n) * H + y) * W + x) * C_out_stride + c0, acc);
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

}  // namespace plc
