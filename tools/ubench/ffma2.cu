// micro-benchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096, CH = 8;
__global__ void scalar(float* out, float a, float b) {
  float v[2 * CH];
  for (int i = 0; i < 2 * CH; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) v[i] = fmaf(v[i], a, b);
  float s = 0;
  for (int i = 0; i < 2 * CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void packed(float* out, float a, float b) {
  unsigned long long v[CH], aa, bb;
  asm("mov.b64 %0, {%1,%2};" : "=l"(aa) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1,%2};" : "=l"(bb) : "f"(b), "f"(b));
  for (int i = 0; i < CH; ++i) { float x = threadIdx.x + 2 * i, y = x + 1; asm("mov.b64 %0, {%1,%2};" : "=l"(v[i]) : "f"(x), "f"(y)); }
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(aa), "l"(bb));
  float s = 0;
  for (int i = 0; i < CH; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v[i])); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) scalar<<<148 * 8, 256>>>(d, 0.999f, 0.001f); else packed<<<148 * 8, 256>>>(d, 0.999f, 0.001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 8 * 256 * ITER * 2 * CH;
      if (rep == 2) printf("%s: %.3f ms, %.1f G fp32-FMA/s per SM-clk-lane... = %.2f TFLOP/s\n", which ? "FFMA2" : "FFMA ", ms, fma / ms / 1e6, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
