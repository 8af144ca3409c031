// Microbenchmark: MUFU.EX2 / FFMA / tcgen05.ld throughput per SM on this GPU (evidence for the roofline of the exp-bound kernels).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// packed 16-bit variants: two exponentials per instruction in PTX (SASS: one MUFU.EX2.BF16 / .F16 per half)
__device__ __forceinline__ float ex2_bf16x2(float x) { unsigned u = __float_as_uint(x), y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(u)); return __uint_as_float(y); }
__device__ __forceinline__ float ex2_f16x2(float x) { unsigned u = __float_as_uint(x), y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(u)); return __uint_as_float(y); }
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 1e-3f + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);                        // MUFU only
      if (MODE == 1) a[i] = fmaf(a[i], 0.999f, 1e-3f);        // FFMA only
      if (MODE == 2) a[i] = ex2(fmaf(a[i], 0.999f, -1e-3f));  // FFMA + MUFU
      if (MODE == 3) { a[i] = ex2(fmaf(a[i], 0.999f, -1e-3f)); a[(i + 1) & 15] += a[i]; }  // + FADD
      if (MODE == 4) a[i] = ex2_bf16x2(a[i]);                 // 2 bf16 exponentials per op
      if (MODE == 5) a[i] = ex2_f16x2(a[i]);                  // 2 f16 exponentials per op
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int threads) {
  int sms = 148, iters = 4096;
  float* out; cudaMalloc(&out, sms * threads * 4);
  k<MODE><<<sms, threads>>>(out, 16, 0.5f);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<sms, threads>>>(out, iters, 0.5f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)sms * threads * iters * 16;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-14s threads/SM=%4d: %.3f ms  %.2f Gop/s  = %.2f op/clk/SM (at %d MHz nominal)\n", name, threads, ms, ops / ms / 1e6, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}
int main() {
  for (int t : {128, 256, 512, 1024}) { run<0>("ex2", t); run<1>("ffma", t); run<2>("ffma+ex2", t); run<3>("ffma+ex2+fadd", t); run<4>("ex2.bf16x2 (x2)", t); run<5>("ex2.f16x2 (x2)", t); }
  return 0;
}
