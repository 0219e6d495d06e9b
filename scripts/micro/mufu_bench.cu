// Microbenchmark: MUFU ex2 throughput per SM for f32, f16x2 and bf16x2 operands (sm_100a).  Build: nvcc -arch=sm_100a -O3.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = a + 0.5f, c = a + 0.25f, d = a + 0.75f;
  unsigned ua = 0x3c003800u + threadIdx.x, ub = ua + 7, uc = ua + 11, ud = ua + 13;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ua));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ub));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(uc));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ud));
    } else {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ua));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ub));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(uc));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ud));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ua ^ ub ^ uc ^ ud);
}
template <int MODE>
void run(const char* name, int elems) {
  float* o;
  cudaMalloc(&o, 148 * 8 * 256 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(o, 100);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(o, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = 148.0 * 8 * 256 * 4.0 * iters;
  printf("%-8s %.3f ms  %.1f G thread-instr/s  %.1f G elements/s  (%.2f elements/clk/SM at 1.9 GHz)\n", name, ms, instr / ms / 1e6,
         instr * elems / ms / 1e6, instr * elems / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(o);
}
int main() {
  run<0>("f32", 1);
  run<1>("f16x2", 2);
  run<2>("bf16x2", 2);
  return 0;
}
