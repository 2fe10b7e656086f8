// common.cuh -- shared helpers of libsdvar_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include <atomic>

#include "../../include/sdvar_b200.h"

namespace sdvar {

// ---- error plumbing (no exceptions cross the C ABI) -------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_arch();  // SDVAR_OK iff the current device is sm_100
int sm_count();    // multiprocessor count of the current device (cached per thread)

#define SDVAR_REQUIRE(cond, ...)                     \
  do {                                               \
    if (!(cond)) {                                   \
      sdvar::set_error(__VA_ARGS__);                 \
      return SDVAR_ERR_ARG;                          \
    }                                                \
  } while (0)

#define SDVAR_CUDA(call)                                                                       \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      sdvar::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SDVAR_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

#define SDVAR_LAUNCH_CHECK()                                                                   \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) {                                                                  \
      sdvar::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SDVAR_ERR_CUDA;                                                                   \
    }                                                                                          \
    sdvar::count_launch();                                                                     \
  } while (0)

// cudaFuncSetAttribute applies to the CURRENT device only: the opt-in to > 48 KiB of dynamic shared memory is repeated
// once per device (bit d of an atomic mask), so a second GPU driven from the same process gets it too; concurrent
// first calls at worst set the attribute twice.
#define SDVAR_SET_SMEM_ONCE(func, bytes)                                                                      \
  do {                                                                                                        \
    static std::atomic<unsigned long long> mask__{0};                                                         \
    int dev__ = 0;                                                                                            \
    SDVAR_CUDA(cudaGetDevice(&dev__));                                                                        \
    const unsigned long long bit__ = 1ull << (dev__ & 63);                                                    \
    if (!(mask__.load(std::memory_order_acquire) & bit__)) {                                                  \
      SDVAR_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      mask__.fetch_or(bit__, std::memory_order_release);                                                      \
    }                                                                                                         \
  } while (0)

// ---- optional per-family device timing (bench.py's roofline leg).  Disabled by default: zero overhead on the
// product path.  When enabled every entry point brackets its launches with a cudaEvent pair on the launch stream.
enum Family { FAM_GEMM = 0, FAM_ATTN = 1, FAM_LN = 2, FAM_SAMPLE = 3, FAM_VERIFY = 4, FAM_VQ = 5, FAM_EMBED = 6, FAM_MISC = 7, FAM_CONV = 8, FAM_COUNT = 9 };
struct ProfileScope {
  ProfileScope(cudaStream_t st, int family, double work);
  ~ProfileScope();
  cudaStream_t st;
  int slot;
};

struct SegTable {
  int S;
  int begin[SDVAR_MAX_SEG + 1];
  float t1[SDVAR_MAX_SEG];
  float t2[SDVAR_MAX_SEG];
};

__device__ __forceinline__ int seg_of(const SegTable& s, int pos) {
  int j = 0;
#pragma unroll 1
  while (j + 1 < s.S && pos >= s.begin[j + 1]) ++j;
  return j;
}

// ---- streaming 128-bit loads / stores -----------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }

// order-preserving float -> uint32 (oracle/spec_c/sdvar_spec.c:fkey)
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t b = f2u(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) { return u2f((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace sdvar
