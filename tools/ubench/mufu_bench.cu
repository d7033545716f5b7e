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
    run<4>("ffma", w);
  }
  return 0;
}
