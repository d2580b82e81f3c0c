// MUFU.EX2 throughput microbenchmark (B200): f32 vs packed f16x2 / bf16x2, 8 independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8]; uint32_t u[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-6f + i * 0.1f; u[i] = 0x38003400u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]) - 1.0f;
      if (MODE == 1) u[i] = ex2_h2(u[i]) ^ 0x04000400u;
      if (MODE == 2) u[i] = ex2_b2(u[i]) ^ 0x00800080u;
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* out) {
  for (int threads : {256, 512, 1024}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148, threads>>>(out, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * threads * 8.0 * iters * (MODE ? 2.0 : 1.0);
    printf("%s threads/SM %4d: %.1f Gexp/s  -> %.2f exp/clk/SM at 1.9 GHz (%.3f ms)\n", name, threads, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9, ms);
  }
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  run<0>("f32   ", out); run<1>("f16x2 ", out); run<2>("bf16x2", out);
  return 0;
}
