// Frame front-end on the tensor cores with the im2col done by the producer warps:
//     x_t = relu(init_conv(add_coord_channels(frame_t)))          generator.py:166-168, coordconv.py:3-10
// for all T frames of a batch in ONE launch, frames [B, T, CF, H, W] fp32 (reference layout) -> [T*B, H, W, 64] bf16.
//
// Why a dedicated kernel: with CF + 2 = 3 input channels the generic implicit GEMM spends its time on operand plumbing
// (one TMA box or one shifted UMMA view per tap, 9 short MMAs per tile behind a generic issue loop: ~2700 issue cycles
// per 128 pixels, 570 us for 320 frames of 128 x 128).  Here K = 9 * (CF + 2) <= 64 fits ONE 128-byte row: eight
// producer warps read the raw frames, generate the two coordinate planes analytically, and write each pixel's 9-tap
// row straight into a SWIZZLE_128B K-major smem tile; the MMA warp needs K/16 <= 4 aligned MMAs per tile and the
// kernel becomes output-bound (16 KB written per tile through a TMA tensor store).  Weights (<= 8 KB) stay resident.
//
// warp 0: MMA issuer + TMEM owner | warps 1-8: producers (two groups of 128 threads, one tile each) |
// warps 9-12: epilogue (bias + ReLU + bf16, swizzled staging, TMA store)
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "plc_ptx.cuh"

namespace plc {

struct FrontendTcParams {
  int B, T, H, W;
  int num_tiles;               // ceil(T*B*H*W / 128)
  const float* frames;         // [B, T, CF, H, W]
  const float* w;              // init_conv.weight [64, CF+2, 3, 3] fp32
  const float* bias;           // [64] or nullptr
};

constexpr int kFeStages = 6;                 // A tiles in flight (16 KB each)
constexpr int kFeAcc = 4;                    // accumulator stages (64 TMEM columns each)
constexpr int kFeN = 64;                     // output channels
constexpr int kFeThreads = 32 * 13;
constexpr int kFeSmemBytes = 1024 + kFeStages * 16384 + 8192 /*B*/ + 2 * 16384 /*staging*/ + 256 /*bias*/ + 256;

__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}

template <int CF>
__global__ void __launch_bounds__(kFeThreads, 1)
frontend_tc_kernel(const FrontendTcParams p, const __grid_constant__ CUtensorMap tmap_out) {
  constexpr int CI = CF + 2;                 // channels per tap: frame channels, row coordinate, column coordinate
  constexpr int K = 9 * CI;
  constexpr int KSTEPS = (K + 15) / 16;
  static_assert(K <= 64, "one 128-byte K row per pixel");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;                                    // [kFeStages][128 rows][128 B]
  uint8_t* smem_b = smem + kFeStages * 16384;                // [64 rows][128 B]
  uint8_t* stage_out = smem_b + 8192;                        // [2][128 rows][128 B]
  float* bias_s = reinterpret_cast<float*>(stage_out + 2 * 16384);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + 2 * 16384 + 256);
  uint64_t* full_bar = bars;                                 // [kFeStages] 4 producer-warp arrivals
  uint64_t* empty_bar = bars + kFeStages;                    // [kFeStages] tcgen05.commit
  uint64_t* tmem_full = bars + 2 * kFeStages;                // [kFeAcc]
  uint64_t* tmem_empty = bars + 2 * kFeStages + kFeAcc;      // [kFeAcc] 4 epilogue-warp arrivals
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * kFeStages + 2 * kFeAcc);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int my_tiles = p.num_tiles > static_cast<int>(blockIdx.x)
                           ? (p.num_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kFeStages; ++s) { mbar_init(&full_bar[s], 4); mbar_init(&empty_bar[s], 1); }
      for (int s = 0; s < kFeAcc; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
      fence_mbar_init();
    }
    tmem_alloc<1>(tmem_ptr_s, kFeAcc * kFeN);
  }
  // resident weight tile: B[n][k], k = tap * CI + c (tap = ky * 3 + kx), K-major, SWIZZLE_128B, zero beyond K
  for (int i = threadIdx.x; i < kFeN * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    float v = 0.f;
    if (k < K) {
      const int tap = k / CI, c = k - tap * CI;
      v = p.w[(n * CI + c) * 9 + tap];
    }
    reinterpret_cast<__nv_bfloat16*>(smem_b)[n * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = __float2bfloat16(v);
  }
  if (threadIdx.x < kFeN) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;
  const int HW = p.H * p.W;
  const int total = p.B * p.T * HW;

  if (warp == 0) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, kFeN, 0, 0);
    const uint64_t adesc0 = make_smem_desc(smem_u32(smem_a), 0, 1024);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(smem_b), 0, 1024);
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % kFeStages, acc = it % kFeAcc;
      mbar_wait(&tmem_empty[acc], ((it / kFeAcc) & 1) ^ 1);
      mbar_wait(&full_bar[st], (it / kFeStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = adesc0 + st * (16384 >> 4);
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k)
          umma_bf16<1>(tmem_base + acc * kFeN, adesc + 2 * k, bdesc0 + 2 * k, idesc, k != 0);
        umma_commit<1>(&empty_bar[st]);
        umma_commit<1>(&tmem_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp <= 8) {
    // ===================================================================== producers: im2col rows into smem
    const int grp = (warp - 1) >> 2;                         // two groups alternate tiles
    const int row = ((warp - 1) & 3) * 32 + lane;            // pixel row of the tile
    const float inv_h = p.H > 1 ? 1.f / (p.H - 1) : 0.f, inv_w = p.W > 1 ? 1.f / (p.W - 1) : 0.f;
    for (int it = grp; it < my_tiles; it += 2) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int st = it % kFeStages;
      const int P = tile * 128 + row;
      float v[64];
#pragma unroll
      for (int k = K; k < 64; ++k) v[k] = 0.f;
      if (P < total) {
        const int n = P / HW, rem = P - n * HW;              // n = t * B + b (T-major output)
        const int y = rem / p.W, x = rem - y * p.W;
        const int t = n / p.B, b = n - t * p.B;
        const float* src = p.frames + (static_cast<size_t>(b) * p.T + t) * CF * HW;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            const bool in = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;      // zero "same" padding of init_conv
            const int k0 = (ky * 3 + kx) * CI;
#pragma unroll
            for (int c = 0; c < CF; ++c) v[k0 + c] = in ? __ldg(src + static_cast<size_t>(c) * HW + yy * p.W + xx) : 0.f;
            v[k0 + CF] = in ? yy * inv_h : 0.f;              // coordconv.py:3-10: row plane, then column plane
            v[k0 + CF + 1] = in ? xx * inv_w : 0.f;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = 0.f;
      }
      mbar_wait(&empty_bar[st], ((it / kFeStages) & 1) ^ 1);
      const uint32_t dst = smem_u32(smem_a) + st * 16384 + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        st_shared_v4(dst + ((j ^ (row & 7)) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]),
                     pack_bf16x2(v[8 * j + 2], v[8 * j + 3]), pack_bf16x2(v[8 * j + 4], v[8 * j + 5]),
                     pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
      fence_proxy_async_smem();                              // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[st]);
    }
  } else {
    // ===================================================================== epilogue
    const int q = warp & 3;                                  // TMEM lane quadrant == warp_idx % 4
    const int row = q * 32 + lane;
    const bool issuer = warp == 9 && lane == 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int acc = it % kFeAcc;
      mbar_wait(&tmem_full[acc], (it / kFeAcc) & 1);
      tc_fence_after();
      if (issuer) tma_store_wait_read_n<1>();                // the store that last read this staging buffer is done
      named_bar_sync(1, 128);
      const uint32_t so = smem_u32(stage_out) + (it & 1) * 16384 + row * 128;
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kFeN;
#pragma unroll
      for (int cc = 0; cc < kFeN / 16; ++cc) {
        uint32_t r[16];
        tmem_ld16(t_acc + cc * 16, r);
        tmem_ld_wait();
        if (cc == kFeN / 16 - 1) {                           // accumulator drained: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = fmaxf(__uint_as_float(r[hlf * 8 + e]) + bias_s[cc * 16 + hlf * 8 + e], 0.f);
          st_shared_v4(so + (((cc * 2 + hlf) ^ (row & 7)) << 4), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                       pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer) {                                          // rows past the tensor end are clipped by the TMA unit
        tma_store_2d(&tmap_out, smem_u32(stage_out) + (it & 1) * 16384, 0, tile * 128);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kFeAcc * kFeN);
  }
}

}  // namespace plc
