// micro-benchmark: MUFU throughput of ex2.approx.ftz.f32 vs the packed ex2.approx.ftz.bf16x2 / ex2.approx.f16x2 on sm_100a
// (exponentials per SM per clock).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ex2 ex2.cu && ./ex2
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096, CH = 8;
__global__ void k_f32(float* out) {
  float v[CH];
  for (int i = 0; i < CH; ++i) v[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
  float s = 0;
  for (int i = 0; i < CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_bf16x2(float* out) {
  unsigned v[CH];
  for (int i = 0; i < CH; ++i) v[i] = 0xBC00BC00u + threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
  unsigned s = 0;
  for (int i = 0; i < CH; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
__global__ void k_f16x2(float* out) {
  unsigned v[CH];
  for (int i = 0; i < CH; ++i) v[i] = 0xB000B000u + threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
  unsigned s = 0;
  for (int i = 0; i < CH; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int which = 0; which < 3; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_f32<<<148 * 8, 256>>>(d); else if (which == 1) k_bf16x2<<<148 * 8, 256>>>(d); else k_f16x2<<<148 * 8, 256>>>(d);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double inst = 148.0 * 8 * 256 * ITER * CH;
      const double el = inst * (which ? 2 : 1);
      if (rep == 2) printf("%s: %.3f ms  %.2f exponentials / SM / clock (at the %d kHz boost clock), %.2f MUFU lane-ops / SM / clock\n",
                           which == 0 ? "ex2.f32   " : which == 1 ? "ex2.bf16x2" : "ex2.f16x2 ", ms, el / (ms * 1e-3) / 148 / (clk * 1e3), clk, inst / (ms * 1e-3) / 148 / (clk * 1e3));
    }
  }
  return 0;
}
