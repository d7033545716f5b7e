// Micro-benchmark: per-SM throughput of the transcendental ops the LSTM epilogue can use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) {  // tanh.approx.f16x2 : two results per instruction
    unsigned u = __float_as_uint(x), v;
    asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(v) : "r"(u));
    y = __uint_as_float(v);
  }
  if (OP == 4) y = fmaf(x, 1.0001f, 0.5f);
  if (OP == 6) {  // register-register FFMA (no immediates)
    asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=f"(y) : "f"(x));
  }
  return y;
}

template <int OP>
__global__ void k(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<OP>(a[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed fp32: fma.rn.f32x2 does two FMAs per instruction (Blackwell FFMA2)
__global__ void k_f32x2(float* out, int iters) {
  unsigned long long a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (unsigned long long)__float_as_uint(threadIdx.x * 1e-3f + i * 0.1f) * 0x100000001ull;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(a[i]));
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s);
}
void run_x2(int warps_per_sm) {
  int sms = 148, iters = 4096;
  float* out;
  cudaMalloc(&out, sms * warps_per_sm * 32 * sizeof(float));
  k_f32x2<<<sms, warps_per_sm * 32>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_f32x2<<<sms, warps_per_sm * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ops = double(sms) * warps_per_sm * 32 * iters * 8;
  printf("%-18s warps/SM %2d : %.2f thread-INSTR/clk/SM (x2 results each) at %d MHz nominal\n", "fma.rn.f32x2", warps_per_sm,
         ops / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree(out);
}

template <int OP>
void run(const char* name, int warps_per_sm) {
  int sms = 148, iters = 4096;
  float* out;
  cudaMalloc(&out, sms * warps_per_sm * 32 * sizeof(float));
  k<OP><<<sms, warps_per_sm * 32>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<sms, warps_per_sm * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ops = double(sms) * warps_per_sm * 32 * iters * 8;
  printf("%-18s warps/SM %2d : %.1f Gop/s  -> %.2f thread-ops/clk/SM at %d MHz (nominal)\n", name, warps_per_sm,
         ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) {
    run<0>("tanh.approx.f32", w);
    run<3>("tanh.approx.f16x2", w);
    run<1>("ex2.approx.f32", w);
    run<2>("rcp.approx.f32", w);
    run<4>("ffma (imm)", w);
    run<6>("ffma (reg)", w);
    run_x2(w);
  }
  return 0;
}
