// Probe: issue-to-completion cost of short tcgen05.mma instructions (M=128, K=16) that accumulate into the SAME TMEM
// tile (a dependent chain) versus round-robin over several independent accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../pl-convlstm-gan_b200/csrc -o umma_latency_probe \
//        umma_latency_probe.cu && ./umma_latency_probe
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "plc_ptx.cuh"

using namespace plc;

template <int N>
__global__ void __launch_bounds__(128) probe(long long* out, int n_mma, int shifted) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<1>(tptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
  const uint64_t da = make_smem_desc(smem_u32(smem), 0, 1024);
  const uint64_t db = make_smem_desc(smem_u32(smem) + 16384, 0, 1024);
  int phase = 0;
  int slot = 0;
  for (int accs = 1; accs <= 512 / N && accs <= 8; accs *= 2) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        if (shifted == 1) {   // conv patch views: A start shifted by (ky * 10 + kx) rows, 8-row groups 1280 B apart
          const int tap = i % 9;
          const uint64_t dav = make_smem_desc(smem_u32(smem) + ((tap / 3) * 10 + tap % 3) * 128, 0, 1280);
          umma_bf16<1>(tmem + (i % accs) * N, dav, db + 2 * (i & 3), idesc, i >= accs);
        } else if (shifted == -1) {   // MN-major A and B (the wgrad operand form): atoms 8 KB apart along M / N
          umma_bf16<1>(tmem + (i % accs) * N, make_smem_desc(smem_u32(smem), 8192, 1024) + 128 * (i & 3),
                       make_smem_desc(smem_u32(smem) + 16384, 8192, 1024) + 128 * (i & 3),
                       make_idesc_bf16(128, N, 1, 1), i >= accs);
        } else if (shifted >= 2) {   // one fixed view: start row = shifted >> 16, SBO = shifted & 0xffff
          const uint64_t dav = make_smem_desc(smem_u32(smem) + (shifted >> 16) * 128, 0, shifted & 0xffff);
          umma_bf16<1>(tmem + (i % accs) * N, dav + 2 * (i & 3), db + 2 * (i & 3), idesc, i >= accs);
        } else {
          umma_bf16<1>(tmem + (i % accs) * N, da + 2 * (i & 3), db + 2 * (i & 3), idesc, i >= accs);
        }
      }
      const long long t1 = clock64();
      umma_commit<1>(bar);
      mbar_wait(bar, phase);
      const long long t2 = clock64();
      out[slot * 2] = t1 - t0;
      out[slot * 2 + 1] = t2 - t0;
    }
    phase ^= 1;
    ++slot;
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tmem, 512);
}

template <int N>
void run(int n_mma, int shifted = 0) {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  cudaMemset(d, 0, 64 * sizeof(long long));
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024 + 1024);
  probe<N><<<1, 128, 50 * 1024 + 1024>>>(d, n_mma, shifted);
  probe<N><<<1, 128, 50 * 1024 + 1024>>>(d, n_mma, shifted);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
  long long h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int slot = 0;
  for (int accs = 1; accs <= 512 / N && accs <= 8; accs *= 2, ++slot) {
    if (shifted == -1) printf("[MN-major A and B] ");
    if (shifted >= 2) printf("[start row %d, SBO %d] ", shifted >> 16, shifted & 0xffff);
    printf("%sM=128 N=%3d K=16, %3d MMAs round-robin over %d accumulator(s): issue %6.1f cyc/MMA, to completion %6.1f "
           "cyc/MMA (ideal tensor time %d)\n", shifted == 1 ? "[shifted patch views] " : "", N, n_mma, accs, double(h[slot * 2]) / n_mma,
           double(h[slot * 2 + 1]) / n_mma, 128 * N * 16 * 2 / 8192);
  }
  cudaFree(d);
}

int main() {
  run<64>(64);
  run<128>(64);
  run<256>(64);
  run<64>(63, 1);
  run<128>(63, 1);
  run<256>(63, 1);
  run<64>(64, -1);
  run<128>(64, -1);
  run<256>(64, -1);
  // which property of a view is slow: the unaligned start row, or the group stride?
  const int views[][2] = {{0, 1024}, {0, 1280}, {0, 2048}, {1, 1024}, {1, 2048}, {8, 1024}, {10, 1280}, {16, 2048}, {17, 2048}};
  for (auto& v : views) run<128>(64, (v[0] << 16) | v[1]);
  return 0;
}
