// gemm_epilogue.cuh -- fused GEMM epilogues shared by the 1-CTA and 2-CTA tcgen05 kernels.
// A thread owns one output row (TMEM lane) and walks the tile's columns in 32-column TMEM loads.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace sdvar {
namespace gemm {

struct Epi {
  const float* bias;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ldo;
  const float* gate;
  int ld_gate, tokens_per_img;
  const int* slot_map;   // pass-image -> gate row / KV-cache slot (nullptr: identity)
  __nv_bfloat16 *q_out, *k_cache, *vT_cache;
  const float* scale_mul;
  int H, Lq, Lmax, Lmax_pad, kv_off, l2norm, C;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5*x*(1+tanh(sqrt(2/pi)*(x+0.044715x^3)))  (nn.GELU(approximate='tanh'), models/basic_var.py:40)
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    d[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                      pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
}

// Residual epilogue: the fp32 residual tile is known before the accumulator is ready, so the first
// 32-column chunk is loaded while the main loop of the tile is still running, and chunk c+1 is in flight while chunk c
// is combined and stored (the proj GEMM, K = C, is otherwise epilogue-bound on these global round trips).
struct EpiPre {
  float4 x[8];
};
template <int EPI>
__device__ __forceinline__ void epilogue_prefetch(EpiPre& pre, int row, int col_base, int M, int N, const Epi& ep) {
  if constexpr (EPI == SDVAR_EPI_RESID_F32) {
    if (row < M && col_base < N) {
      const float4* x4 = reinterpret_cast<const float4*>(ep.out_f32 + (size_t)row * ep.ldo + col_base);
#pragma unroll
      for (int q = 0; q < 8; ++q) pre.x[q] = x4[q];
    }
  }
}

template <int EPI, int BN>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, int row, int col_base, int M, int N, const Epi& ep, EpiPre& pre) {
  if constexpr (EPI == SDVAR_EPI_RESID_F32) {
    const bool rv = row < M;
    const int gimg = rv ? row / ep.tokens_per_img : 0;
    const float* grow = ep.gate + (size_t)(ep.slot_map != nullptr ? __ldg(ep.slot_map + gimg) : gimg) * ep.ld_gate;
    float* xrow = ep.out_f32 + (size_t)row * ep.ldo;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(taddr + c * 32, r);
      const int col0 = col_base + c * 32, coln = col0 + 32;
      float4 xn[8];
      if (c + 1 < BN / 32 && rv && coln < N) {
        const float4* x4 = reinterpret_cast<const float4*>(xrow + coln);
#pragma unroll
        for (int q = 0; q < 8; ++q) xn[q] = x4[q];
      }
      ptx::tmem_ld_wait();
      if (rv && col0 < N) {
        float4* d = reinterpret_cast<float4*>(xrow + col0);
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
        const float4* g4 = reinterpret_cast<const float4*>(grow + col0);   // one gate row per image: warp-uniform, L1-resident
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = ep.bias != nullptr ? __ldg(b4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 o = pre.x[q];
          const float4 g = __ldg(g4 + q);
          o.x += (__uint_as_float(r[4 * q]) + bb.x) * g.x;
          o.y += (__uint_as_float(r[4 * q + 1]) + bb.y) * g.y;
          o.z += (__uint_as_float(r[4 * q + 2]) + bb.z) * g.z;
          o.w += (__uint_as_float(r[4 * q + 3]) + bb.w) * g.w;
          d[q] = o;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) pre.x[q] = xn[q];
    }
  } else if constexpr (EPI == SDVAR_EPI_QKV) {
    // 64-column groups = one attention head of one of q / k / v
#pragma unroll 1
    for (int g0 = 0; g0 < BN; g0 += 64) {
      uint32_t r0[32], r1[32];
      ptx::tmem_ld_32x32(taddr + g0, r0);
      ptx::tmem_ld_32x32(taddr + g0 + 32, r1);
      ptx::tmem_ld_wait();
      const int col0 = col_base + g0;
      if (row < M && col0 < N) {
        float v[64];
        const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = __ldg(b4 + q), bc = __ldg(b4 + 8 + q);
          v[4 * q] = __uint_as_float(r0[4 * q]) + bb.x; v[4 * q + 1] = __uint_as_float(r0[4 * q + 1]) + bb.y;
          v[4 * q + 2] = __uint_as_float(r0[4 * q + 2]) + bb.z; v[4 * q + 3] = __uint_as_float(r0[4 * q + 3]) + bb.w;
          v[32 + 4 * q] = __uint_as_float(r1[4 * q]) + bc.x; v[32 + 4 * q + 1] = __uint_as_float(r1[4 * q + 1]) + bc.y;
          v[32 + 4 * q + 2] = __uint_as_float(r1[4 * q + 2]) + bc.z; v[32 + 4 * q + 3] = __uint_as_float(r1[4 * q + 3]) + bc.w;
        }
        const int sect = col0 / ep.C, h = (col0 - sect * ep.C) >> 6;
        if (sect < 2 && ep.l2norm) {
          float ss = 0.0f;
#pragma unroll
          for (int i = 0; i < 64; ++i) ss += v[i] * v[i];
          float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (models/basic_var.py:103-104)
          if (sect == 0) inv *= __expf(fminf(__ldg(ep.scale_mul + h), 4.605170185988092f));  // clamp_max(log 100).exp()
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] *= inv;
        }
        const int img = row / ep.Lq, t = row - img * ep.Lq;
        const int simg = ep.slot_map != nullptr ? __ldg(ep.slot_map + img) : img;   // cache slot of this image
        if (sect < 2) {
          __nv_bfloat16* dst = (sect == 0)
              ? ep.q_out + (((size_t)img * ep.H + h) * ep.Lq + t) * 64
              : ep.k_cache + (((size_t)simg * ep.H + h) * ep.Lmax + ep.kv_off + t) * 64;
          uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            d[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                              pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        } else {
          __nv_bfloat16* dst = ep.vT_cache + ((size_t)simg * ep.H + h) * 64 * ep.Lmax_pad + ep.kv_off + t;
#pragma unroll
          for (int i = 0; i < 64; ++i) dst[(size_t)i * ep.Lmax_pad] = __float2bfloat16_rn(v[i]);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(taddr + c0, r);
      ptx::tmem_ld_wait();
      const int col0 = col_base + c0;
      if (row < M && col0 < N) {
        float v[32];
        if (ep.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = __ldg(b4 + q);
            v[4 * q] = __uint_as_float(r[4 * q]) + bb.x; v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bb.y;
            v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bb.z; v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bb.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        }
        if constexpr (EPI == SDVAR_EPI_F32) {
          float4* d = reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * ep.ldo + col0);
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else if constexpr (EPI == SDVAR_EPI_BF16) {
          store_bf16x32(ep.out_bf16 + (size_t)row * ep.ldo + col0, v);
        } else if constexpr (EPI == SDVAR_EPI_GELU_BF16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
          store_bf16x32(ep.out_bf16 + (size_t)row * ep.ldo + col0, v);
        }
      }
    }
  }
}


}  // namespace gemm
}  // namespace sdvar
