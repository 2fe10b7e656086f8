// elementwise.cu -- LayerNorm + adaLN modulate (fp32 -> bf16 GEMM operand), SiLU, casts.
//
// ln_modulate replaces `ln_wo_grad(x).mul(scale.add(1)).add_(shift)` at models/basic_var.py:157-158 and
// :173-174.  HBM-bound: one warp per row, the row is staged once in shared memory (single global read),
// mean and centred variance are reduced with warp shuffles, the bf16 result is written with 8-byte stores.
#include <stdlib.h>

#include "common.cuh"

namespace sdvar {

constexpr int kLnWarps = 4;

__global__ void __launch_bounds__(kLnWarps * 32)
ln_modulate_kernel(const float* __restrict__ x, int M, int C, int tokens_per_img, const float* __restrict__ scale,
                   const float* __restrict__ shift, int ld_mod, const int* __restrict__ slot_map, float eps,
                   __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) float rowbuf[];  // [kLnWarps][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  if (row >= M) return;
  float* buf = rowbuf + (size_t)warp * C;
  const int nvec = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
  float sum = 0.0f;
  for (int i = lane; i < nvec; i += 32) {
    const float4 v = ldg_stream(xr + i);
    reinterpret_cast<float4*>(buf)[i] = v;
    sum += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float mean = sum / (float)C;
  float var = 0.0f;
  for (int i = lane; i < nvec; i += 32) {
    const float4 v = reinterpret_cast<const float4*>(buf)[i];
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    var += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
  const float rstd = rsqrtf(var / (float)C + eps);
  const int img0 = row / tokens_per_img;
  const int img = slot_map != nullptr ? __ldg(slot_map + img0) : img0;
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)img * ld_mod);
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)img * ld_mod);
  uint2* o = reinterpret_cast<uint2*>(out + (size_t)row * C);
  for (int i = lane; i < nvec; i += 32) {
    const float4 v = reinterpret_cast<const float4*>(buf)[i];
    const float4 s = __ldg(sc + i), h = __ldg(sh + i);
    const float y0 = (v.x - mean) * rstd * (1.0f + s.x) + h.x;
    const float y1 = (v.y - mean) * rstd * (1.0f + s.y) + h.y;
    const float y2 = (v.z - mean) * rstd * (1.0f + s.z) + h.z;
    const float y3 = (v.w - mean) * rstd * (1.0f + s.w) + h.w;
    o[i] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
  }
}

// Register-resident variant for the model widths in use (C = 128*NITER): all NITER 16-byte loads of a lane are issued
// back to back (7.7 KB in flight per warp at C=1920), no shared-memory staging, 8 rows per CTA.
template <int NITER>
__global__ void __launch_bounds__(256, NITER > 15 ? 2 : NITER > 12 ? 3 : 4)   // cap registers: ptxas otherwise hoists all scale/shift loads (127 regs, 2 CTAs/SM)
ln_modulate_reg_kernel(const float* __restrict__ x, int M, int tokens_per_img, const float* __restrict__ scale,
                       const float* __restrict__ shift, int ld_mod, const int* __restrict__ slot_map, float eps,
                       __nv_bfloat16* __restrict__ out) {
  constexpr int C = NITER * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * C);
  float4 v[NITER];
#pragma unroll
  for (int i = 0; i < NITER; ++i) v[i] = ldg_stream(xr + i * 32 + lane);
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < NITER; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float mean = sum * (1.0f / (float)C);
  float var = 0.0f;
#pragma unroll
  for (int i = 0; i < NITER; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    var += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
  const float rstd = rsqrtf(var * (1.0f / (float)C) + eps);
  const int img0 = row / tokens_per_img;
  const int img = slot_map != nullptr ? __ldg(slot_map + img0) : img0;
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)img * ld_mod);
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)img * ld_mod);
  uint2* o = reinterpret_cast<uint2*>(out + (size_t)row * C);
#pragma unroll
  for (int i = 0; i < NITER; ++i) {
    const float4 s = __ldg(sc + i * 32 + lane), h = __ldg(sh + i * 32 + lane);
    const float y0 = (v[i].x - mean) * rstd * (1.0f + s.x) + h.x;
    const float y1 = (v[i].y - mean) * rstd * (1.0f + s.y) + h.y;
    const float y2 = (v[i].z - mean) * rstd * (1.0f + s.z) + h.z;
    const float y3 = (v[i].w - mean) * rstd * (1.0f + s.w) + h.w;
    o[i * 32 + lane] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
  }
}

// Persistent variant for large M: one 256-thread CTA per SM, every warp streams ITS rows through a private 3-deep shared-memory
// ring filled by 1-D bulk TMA copies (cp.async.bulk + mbarrier), so 8 x 3 rows (184 KiB at C = 1920) are in flight per SM at
// all times, independent of the register file.  The register-resident kernel above is limited to 24 warps per SM by its 80
// registers and showed long-scoreboard stalls with the issue slots 24 % busy and DRAM at 56 % of peak (profiles/ncu_ln_r02.md).
constexpr int kLnStages = 3;
__device__ __forceinline__ void ln_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void ln_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LN_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LN_DONE;\n"
      "bra LN_WAIT;\n"
      "LN_DONE:\n"
      "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

template <int NITER>
__global__ void __launch_bounds__(256, 1)
ln_modulate_tma_kernel(const float* __restrict__ x, int M, int tokens_per_img, const float* __restrict__ scale,
                       const float* __restrict__ shift, int ld_mod, const int* __restrict__ slot_map, float eps,
                       __nv_bfloat16* __restrict__ out) {
  constexpr int C = NITER * 128;
  constexpr uint32_t kRowBytes = C * 4;
  extern __shared__ __align__(128) unsigned char ln_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(ln_smem) + (size_t)warp * kLnStages * C;          // [kLnStages][C] of this warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + (size_t)8 * kLnStages * kRowBytes) + warp * kLnStages;
  const long long stride = (long long)gridDim.x * 8;
  const long long first = (long long)blockIdx.x * 8 + warp;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kLnStages; ++st)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[st])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int st = 0; st < kLnStages; ++st) {
      const long long r = first + st * stride;
      if (r < M) ln_bulk_g2s(ring + st * C, x + r * C, kRowBytes, &bars[st]);
    }
  }
  __syncwarp();
  uint32_t k = 0;
  for (long long row = first; row < M; row += stride, ++k) {
    const uint32_t st = k % kLnStages;
    ln_mbar_wait(&bars[st], (k / kLnStages) & 1);
    const float4* src = reinterpret_cast<const float4*>(ring + st * C);
    float4 v[NITER];
#pragma unroll
    for (int i = 0; i < NITER; ++i) v[i] = src[i * 32 + lane];
    __syncwarp();                                   // every lane has its values: the stage can be refilled
    if (lane == 0) {
      const long long nr = row + (long long)kLnStages * stride;
      if (nr < M) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        ln_bulk_g2s(ring + st * C, x + nr * C, kRowBytes, &bars[st]);
      }
    }
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NITER; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float mean = sum * (1.0f / (float)C);
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      var += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
    const float rstd = rsqrtf(var * (1.0f / (float)C) + eps);
    const int img0 = (int)(row / tokens_per_img);
    const int img = slot_map != nullptr ? __ldg(slot_map + img0) : img0;
    const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)img * ld_mod);
    const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)img * ld_mod);
    uint2* o = reinterpret_cast<uint2*>(out + (size_t)row * C);
#pragma unroll
    for (int i = 0; i < NITER; ++i) {
      const float4 s4 = __ldg(sc + i * 32 + lane), h4 = __ldg(sh + i * 32 + lane);
      const float y0 = (v[i].x - mean) * rstd * (1.0f + s4.x) + h4.x;
      const float y1 = (v[i].y - mean) * rstd * (1.0f + s4.y) + h4.y;
      const float y2 = (v[i].z - mean) * rstd * (1.0f + s4.z) + h4.z;
      const float y3 = (v[i].w - mean) * rstd * (1.0f + s4.w) + h4.w;
      o[i * 32 + lane] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
    }
  }
}

// Same ring, rows handed out in BLOCKS of (image, 64 consecutive tokens): the image's scale / shift rows are staged in shared
// memory once per block (double-buffered, one barrier per block) instead of being fetched through L1 for every token row -- in
// the kernel above ~30 % of those 2 x C x 4 bytes per row miss L1 and the warp sits on the L2 latency (long_scoreboard 4.3 per
// issue in profiles/ncu_k3_k4_conv_r02.md).  Same per-row arithmetic in the same order: bit-identical output.
template <int NITER>
__global__ void __launch_bounds__(256, 1)
ln_modulate_tma2_kernel(const float* __restrict__ x, int M, int tokens_per_img, const float* __restrict__ scale,
                        const float* __restrict__ shift, int ld_mod, const int* __restrict__ slot_map, float eps,
                        __nv_bfloat16* __restrict__ out) {
  constexpr int C = NITER * 128;
  constexpr uint32_t kRowBytes = C * 4;
  constexpr int kBlkRows = 64;
  extern __shared__ __align__(128) unsigned char ln_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(ln_smem) + (size_t)warp * kLnStages * C;          // [kLnStages][C] of this warp
  float* smod = reinterpret_cast<float*>(ln_smem + (size_t)8 * kLnStages * kRowBytes);      // [2 buffers][scale, shift][C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ln_smem + (size_t)8 * kLnStages * kRowBytes + (size_t)4 * kRowBytes) + warp * kLnStages;
  const int tpi = tokens_per_img;
  const int cpi = (tpi + kBlkRows - 1) / kBlkRows;                                          // blocks per image
  const long long nblk = (((long long)M + tpi - 1) / tpi) * cpi;
  // rows of block b that fall to this warp: r0 + warp + 8 j, j = 0.. while below lim
  auto block_geom = [&](long long b, long long& r0, int& lim, int& img) {
    img = (int)(b / cpi);
    const int ch = (int)(b - (long long)img * cpi);
    r0 = (long long)img * tpi + (long long)ch * kBlkRows;
    long long l = tpi - ch * kBlkRows;
    l = l < kBlkRows ? l : kBlkRows;
    const long long rest = (long long)M - r0;
    lim = (int)(l < rest ? l : rest);
  };
  // issue-side iterator over this warp's rows, in the order the consumer visits them
  long long iblk = blockIdx.x;
  int ij = 0;
  auto next_row = [&]() -> long long {
    while (iblk < nblk) {
      long long r0; int lim, img;
      block_geom(iblk, r0, lim, img);
      const int rr = warp + 8 * ij;
      if (rr < lim) {
        if (++ij == 8) { ij = 0; iblk += gridDim.x; }
        return r0 + rr;
      }
      ij = 0;
      iblk += gridDim.x;
    }
    return -1;
  };
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kLnStages; ++st)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bars[st])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
#pragma unroll
  for (int st = 0; st < kLnStages; ++st) {
    const long long r = next_row();
    if (r >= 0 && lane == 0) ln_bulk_g2s(ring + st * C, x + r * C, kRowBytes, &bars[st]);
  }
  uint32_t k = 0;
  int par = 0;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x, par ^= 1) {
    long long r0; int lim, img0;
    block_geom(blk, r0, lim, img0);
    const int img = slot_map != nullptr ? __ldg(slot_map + img0) : img0;
    float4* md = reinterpret_cast<float4*>(smod + (size_t)par * 2 * C);
    {
      const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)img * ld_mod);
      const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)img * ld_mod);
      for (int i = threadIdx.x; i < 2 * NITER * 32; i += 256) md[i] = i < NITER * 32 ? __ldg(sc + i) : __ldg(sh + (i - NITER * 32));
    }
    __syncthreads();      // buffer `par` is complete; buffer par^1 is free again once every warp is past this barrier
    for (int j = 0; j < 8; ++j, ++k) {
      const int rr = warp + 8 * j;
      if (rr >= lim) break;
      const long long row = r0 + rr;
      const uint32_t st = k % kLnStages;
      ln_mbar_wait(&bars[st], (k / kLnStages) & 1);
      const float4* src = reinterpret_cast<const float4*>(ring + st * C);
      float4 v[NITER];
#pragma unroll
      for (int i = 0; i < NITER; ++i) v[i] = src[i * 32 + lane];
      __syncwarp();                                   // every lane has its values: the stage can be refilled
      {
        const long long nr = next_row();
        if (nr >= 0 && lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          ln_bulk_g2s(ring + st * C, x + nr * C, kRowBytes, &bars[st]);
        }
      }
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      const float mean = sum * (1.0f / (float)C);
      float var = 0.0f;
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        var += (a * a + b * b) + (c * c + d * d);
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) var += __shfl_xor_sync(0xffffffffu, var, off);
      const float rstd = rsqrtf(var * (1.0f / (float)C) + eps);
      uint2* o = reinterpret_cast<uint2*>(out + (size_t)row * C);
#pragma unroll
      for (int i = 0; i < NITER; ++i) {
        const float4 s4 = md[i * 32 + lane], h4 = md[NITER * 32 + i * 32 + lane];
        const float y0 = (v[i].x - mean) * rstd * (1.0f + s4.x) + h4.x;
        const float y1 = (v[i].y - mean) * rstd * (1.0f + s4.y) + h4.y;
        const float y2 = (v[i].z - mean) * rstd * (1.0f + s4.z) + h4.z;
        const float y3 = (v[i].w - mean) * rstd * (1.0f + s4.w) + h4.w;
        o[i * 32 + lane] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
      }
    }
  }
}

__global__ void silu_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    out[i] = __float2bfloat16_rn(v / (1.0f + __expf(-v)));
  }
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(x[i]);
}

// images in [0,1] fp32 (B,3,H,W) -> uint8, as the reference's notebook emits them (`.mul_(255)` ... `.astype(np.uint8)`,
// sdvar_colab_test.py:235-236: truncation), either in place of layout (B,3,H,W) or as (B,H,W,3) rows for PNG / npz writers
__global__ void image_to_u8_kernel(const float* __restrict__ img, int B, int HW, int hwc, uint8_t* __restrict__ out) {
  const long long n = (long long)B * 3 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(img[i], 0.0f), 1.0f) * 255.0f;
    const uint8_t q = (uint8_t)(int)v;
    if (!hwc) {
      out[i] = q;
    } else {
      const long long b = i / (3LL * HW), r = i - b * 3LL * HW;
      const int c = (int)(r / HW), p = (int)(r - (long long)c * HW);
      out[(b * HW + p) * 3 + c] = q;
    }
  }
}

}  // namespace sdvar

using namespace sdvar;

extern "C" int sdvar_image_to_u8(const float* img_B3HW, int B, int H, int W, int hwc, uint8_t* out, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(img_B3HW && out && B > 0 && H > 0 && W > 0, "bad argument");
  const long long n = (long long)B * 3 * H * W;
  const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  ProfileScope prof((cudaStream_t)stream, FAM_MISC, (double)n * 5.0);
  image_to_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img_B3HW, B, H * W, hwc, out);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

extern "C" int sdvar_ln_modulate(const float* x, int M, int C, int tokens_per_img, const float* scale, const float* shift,
                                 int ld_mod, const int* slot_map, float eps, sdvar_bf16* out, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && scale && shift && out, "NULL argument");
  SDVAR_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && tokens_per_img > 0 && ld_mod % 4 == 0, "bad geometry M=%d C=%d", M, C);
  SDVAR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)scale & 15) == 0 && ((uintptr_t)shift & 15) == 0 && ((uintptr_t)out & 7) == 0,
                "alignment");
  const size_t smem = (size_t)kLnWarps * C * sizeof(float);
  SDVAR_REQUIRE(smem <= 48 * 1024, "C=%d too large for ln_modulate", C);
  ProfileScope prof((cudaStream_t)stream, FAM_LN, (double)M * C * 6.0);
  // large M: persistent TMA-ring kernel (same arithmetic, same per-row order => bit-identical output)
  static const int tma_min_rows = [] { const char* e = getenv("SDVAR_LN_TMA_MIN"); return e ? atoi(e) : 4096; }();
  if (M >= tma_min_rows && ((uintptr_t)x & 15) == 0) {
    const int sms = sm_count();
    const int grid = (M + 7) / 8 < sms ? (M + 7) / 8 : sms;
    // blocked variant: images of >= 128 tokens and >= 16384 rows (same-box A/B at C = 1920 / 1024: 256 tokens x 128 images 75.9 ->
    // 70.0 us / 53.7 -> 41.2 us, 169 tokens 50.2 -> 49.5 / 36.0 -> 31.0; 100 tokens and fewer: the strided kernel wins)
    static const bool blocked = getenv("SDVAR_LN_NO_BLOCKS") == nullptr;      // A/B switch
#define SDVAR_LN_TMA(NITER)                                                                                              \
  case NITER * 128: {                                                                                                    \
    const size_t smem2 = (size_t)(8 * kLnStages + 4) * NITER * 128 * 4 + 8 * kLnStages * 8;                              \
    if (blocked && tokens_per_img >= 128 && M >= 16384 && smem2 <= 227 * 1024) {                                         \
      const long long nblk = (((long long)M + tokens_per_img - 1) / tokens_per_img) * ((tokens_per_img + 63) / 64);      \
      SDVAR_SET_SMEM_ONCE(ln_modulate_tma2_kernel<NITER>, smem2);                                                        \
      ln_modulate_tma2_kernel<NITER><<<(int)(nblk < sms ? nblk : sms), 256, smem2, (cudaStream_t)stream>>>(              \
          x, M, tokens_per_img, scale, shift, ld_mod, slot_map, eps, reinterpret_cast<__nv_bfloat16*>(out));             \
      SDVAR_LAUNCH_CHECK();                                                                                              \
      return SDVAR_OK;                                                                                                   \
    }                                                                                                                    \
    const size_t smem = (size_t)8 * kLnStages * NITER * 128 * 4 + 8 * kLnStages * 8;                                     \
    SDVAR_SET_SMEM_ONCE(ln_modulate_tma_kernel<NITER>, smem);                                                            \
    ln_modulate_tma_kernel<NITER><<<grid, 256, smem, (cudaStream_t)stream>>>(x, M, tokens_per_img, scale, shift, ld_mod, slot_map, eps, \
                                                                           reinterpret_cast<__nv_bfloat16*>(out));     \
    SDVAR_LAUNCH_CHECK();                                                                                                \
    return SDVAR_OK;                                                                                                     \
  }
    switch (C) {
      SDVAR_LN_TMA(8) SDVAR_LN_TMA(10) SDVAR_LN_TMA(12) SDVAR_LN_TMA(15) SDVAR_LN_TMA(18)
      default: break;
    }
#undef SDVAR_LN_TMA
  }
#define SDVAR_LN_REG(NITER)                                                                                              \
  case NITER * 128:                                                                                                      \
    ln_modulate_reg_kernel<NITER><<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, M, tokens_per_img, scale, shift, ld_mod, slot_map, eps, \
                                                                                reinterpret_cast<__nv_bfloat16*>(out)); \
    SDVAR_LAUNCH_CHECK();                                                                                                \
    return SDVAR_OK;
  switch (C) {
    SDVAR_LN_REG(8) SDVAR_LN_REG(10) SDVAR_LN_REG(12) SDVAR_LN_REG(15) SDVAR_LN_REG(18) SDVAR_LN_REG(2) SDVAR_LN_REG(3) SDVAR_LN_REG(4)
    default: break;   // other widths: generic shared-memory kernel below
  }
#undef SDVAR_LN_REG
  ln_modulate_kernel<<<(M + kLnWarps - 1) / kLnWarps, kLnWarps * 32, smem, (cudaStream_t)stream>>>(
      x, M, C, tokens_per_img, scale, shift, ld_mod, slot_map, eps, reinterpret_cast<__nv_bfloat16*>(out));
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

extern "C" int sdvar_silu_bf16(const float* x, long long n, sdvar_bf16* out, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && out && n > 0, "bad argument");
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  silu_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, reinterpret_cast<__nv_bfloat16*>(out));
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}
extern "C" int sdvar_f32_to_bf16(const float* x, long long n, sdvar_bf16* out, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && out && n > 0, "bad argument");
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  f32_to_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, reinterpret_cast<__nv_bfloat16*>(out));
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}
