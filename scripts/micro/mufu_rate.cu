// MUFU.EX2 throughput microbenchmark (B200): warps per SM swept, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void k(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-6f + i * 0.1f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ex2(a[i]) - 1.0f;
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int threads : {128, 256, 512, 1024}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<<<148, threads>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * threads * 8.0 * iters;
    printf("threads/SM %4d: %.1f Gexp/s  -> %.2f exp/clk/SM at 1.9 GHz (%.3f ms)\n", threads, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9, ms);
  }
  return 0;
}
