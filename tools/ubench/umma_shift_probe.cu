// Probe: can a K-major SWIZZLE_128B UMMA operand descriptor start at a row that is NOT a multiple of 8 rows (1024 B),
// and with a group stride (SBO) that is not a multiple of 1024 B?  That is what an implicit-GEMM conv needs to read all
// k*k taps of a tile from ONE haloed activation patch in shared memory (shifted views) instead of k*k separate tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../../pl-convlstm-gan_b200/csrc -o umma_shift_probe \
//        umma_shift_probe.cu && ./umma_shift_probe
// A patch: rows p = 0..319 of 64 bf16, written with the TMA SWIZZLE_128B pattern of a 1024-aligned box
// (16-byte chunk c of row p lives at chunk c ^ (p & 7)).  B = 64x64 identity, so D[m][n] = A[row(m)][n]:
// even columns encode which patch row was read, odd columns must equal their own index (chunk integrity).
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "plc_ptx.cuh"

using namespace plc;

struct Variant { int shift_rows, sbo_bytes, base_offset; };
constexpr int kRows = 320;

__global__ void __launch_bounds__(128) probe(const Variant* vs, int nv, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                    // kRows x 128 B
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem + kRows * 128);     // 64 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kRows * 128 + 64 * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kRows * 64; i += 128) {
    const int p = i >> 6, k = i & 63;
    float v;
    if (k & 1) v = static_cast<float>(k);
    else if ((k & 3) == 0) v = static_cast<float>(p & 15);
    else v = static_cast<float>(p >> 4);
    A[p * 64 + (((k >> 3) ^ (p & 7)) << 3) + (k & 7)] = __float2bfloat16(v);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int n = i >> 6, k = i & 63;
    Bm[n * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<1>(tptr, 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  for (int v = 0; v < nv; ++v) {
    if (tid == 0) {
      const Variant var = vs[v];
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = make_smem_desc(smem_u32(A) + var.shift_rows * 128 + kk * 32, 0, var.sbo_bytes);
        da |= static_cast<uint64_t>(var.base_offset & 7) << 49;
        const uint64_t db = make_smem_desc(smem_u32(Bm) + kk * 32, 0, 1024);
        umma_bf16<1>(tmem, da, db, idesc, kk > 0);
      }
      umma_commit<1>(bar);
    }
    mbar_wait(bar, v & 1);
    tc_fence_after();
    for (int c = 0; c < 64; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[(static_cast<size_t>(v) * 128 + tid) * 64 + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc<1>(tmem, 64);
}

int main() {
  const Variant h[] = {
      {0, 1024, 0},                                   // sanity: the standard aligned tile
      {1, 1024, 0}, {1, 1024, 1},                     // contiguous rows, start shifted by one row
      {3, 1024, 0}, {3, 1024, 3},
      {8, 1024, 0},                                   // shifted by a whole swizzle atom (must work)
      {3, 1280, 0}, {3, 1280, 3},                     // 10-pixel patch rows (8-wide tile + halo), shift ky*10+kx
      {11, 1280, 0}, {11, 1280, 3}, {22, 1280, 0}, {22, 1280, 6},
      {1, 2048, 0}, {1, 2048, 1},                     // 16-pixel (padded) patch rows: group stride = 2 atoms
      {17, 2048, 0}, {17, 2048, 1}, {34, 2048, 0}, {34, 2048, 2},
  };
  const int nv = sizeof(h) / sizeof(h[0]);
  Variant* dv; float* dout;
  cudaMalloc(&dv, sizeof(h)); cudaMemcpy(dv, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, sizeof(float) * nv * 128 * 64);
  cudaMemset(dout, 0xff, sizeof(float) * nv * 128 * 64);
  const int smem = kRows * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dv, nv, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  float* o = static_cast<float*>(malloc(sizeof(float) * nv * 128 * 64));
  cudaMemcpy(o, dout, sizeof(float) * nv * 128 * 64, cudaMemcpyDeviceToHost);
  for (int v = 0; v < nv; ++v) {
    int rows_ok = 0, chunks_ok = 0, consistent = 0;
    int got_first[12];
    for (int m = 0; m < 128; ++m) {
      const float* d = o + (static_cast<size_t>(v) * 128 + m) * 64;
      const int want = h[v].shift_rows + (m >> 3) * (h[v].sbo_bytes / 128) + (m & 7);
      const int got = static_cast<int>(d[0]) + 16 * static_cast<int>(d[2]);
      bool cons = true, chunks = true;
      for (int k = 0; k < 64; ++k) {
        if (k & 1) chunks &= d[k] == static_cast<float>(k);
        else if ((k & 3) == 0) cons &= d[k] == d[0];
        else cons &= d[k] == d[2];
      }
      rows_ok += got == want && cons;
      chunks_ok += chunks;
      consistent += cons;
      if (m < 12) got_first[m] = cons ? got : -1;
    }
    printf("shift %2d rows, SBO %4d B, base_offset %d : rows as wanted %3d/128, row-consistent %3d/128, chunk order ok "
           "%3d/128 | rows read by m=0..11:", h[v].shift_rows, h[v].sbo_bytes, h[v].base_offset, rows_ok, consistent,
           chunks_ok);
    for (int m = 0; m < 12; ++m) printf(" %d", got_first[m]);
    printf("\n");
  }
  return 0;
}
