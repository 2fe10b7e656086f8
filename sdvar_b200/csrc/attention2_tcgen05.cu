// attention2_tcgen05.cu -- K2 fast path: one-pass, ping-pong softmax, for l2-normalised attention.
//
// With attn_l2_norm (the reference default, models/basic_var.py:67-70,101-105) q is a unit vector times
// s_h = exp(min(scale_mul_h, ln 100)) and k is a unit vector, so every logit is bounded by |q.k| <= s_h.  Using that bound
// as the softmax reference point, P = exp(S - s_h) never overflows and (for s_h <= 40) never underflows, so no row-maximum
// pass and no accumulator rescaling are needed: one QK^T sweep, P straight to bf16, O += P V.  (Heads with s_h > 40, or
// models without l2 norm, take the two-pass kernel in attention_tcgen05.cu.)
//
// One CTA = 128 queries of one (image, head), 10 warps:
//   warp 0   TMA producer (Q once; K and V^T tiles in 2-deep rings)
//   warp 1   TMEM allocator + MMA issuer: S_b = Q K_j^T into TMEM buffer b = j&1, then O += P_b V_j; S(j+1) is issued before
//            PV(j), so the tensor core works on the next tile while the other softmax group is still exponentiating
//   warps 2-5 / 6-9  two softmax groups (even / odd key tiles), each with its own S buffer in TMEM and P buffer in shared
//            memory (UMMA SWIZZLE_128B layout); ex2.approx on the MUFU; the epilogue is split by head-dim halves.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sdvar {
namespace attn2 {

constexpr int BQ = 128, BKV = 128, D = 64;
constexpr int kThreads = 320;
constexpr int Q_BYTES = BQ * D * 2, K_BYTES = BKV * D * 2, V_BYTES = D * BKV * 2, P_BYTES = BQ * BKV * 2;
constexpr int kTmemCols = 512;  // S0 [0,128) S1 [128,256) O [256,320)
constexpr size_t kSmemBytes = 1024 + Q_BYTES + 2 * K_BYTES + 2 * V_BYTES + 2 * P_BYTES + 1024 /*row sums*/ + 256;

struct Params {
  int H, Lq, kv_off, C;
  float log2e_scale;
  const float* scale_mul;  // [H] raw log-scale parameter
  __nv_bfloat16* out;
  SegTable seg;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
attention_onepass_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + 2 * K_BYTES;
  uint8_t* sP = sV + 2 * V_BYTES;
  float* sL = reinterpret_cast<float*>(sP + 2 * P_BYTES);  // [2][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sL + 256);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;   // [2]
  uint64_t* p_empty = bars + 15;  // [2]
  uint64_t* o_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, img = blockIdx.z;
  const int bh = img * p.H + h;
  const int q0 = qt * BQ;
  const int t_last = min(q0 + BQ, p.Lq) - 1;
  const int max_limit = p.kv_off + p.seg.begin[seg_of(p.seg, t_last) + 1];
  const int nk = (max_limit + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&s_empty[i], 4);
      ptx::mbar_init(&p_full[i], 4); ptx::mbar_init(&p_empty[i], 1);
    }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, Q_BYTES);
      ptx::tma_load_3d(sQ, &tmQ, q_full, 0, q0, bh);
      for (int j = 0; j < nk; ++j) {
        const int b = j & 1;
        const uint32_t par = (uint32_t)((j >> 1) & 1);
        ptx::mbar_wait(&k_empty[b], par ^ 1);
        ptx::mbar_expect_tx(&k_full[b], K_BYTES);
        ptx::tma_load_3d(sK + b * K_BYTES, &tmK, &k_full[b], 0, j * BKV, bh);
        ptx::mbar_wait(&v_empty[b], par ^ 1);
        ptx::mbar_expect_tx(&v_full[b], V_BYTES);
        ptx::tma_load_3d(sV + b * V_BYTES, &tmV, &v_full[b], j * BKV, 0, bh);
        ptx::tma_load_3d(sV + b * V_BYTES + V_BYTES / 2, &tmV, &v_full[b], j * BKV + 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, D);
      const uint32_t q_addr = ptx::smem_u32(sQ);
      auto issue_s = [&](int j) {
        const int b = j & 1;
        const uint32_t par = (uint32_t)((j >> 1) & 1);
        ptx::mbar_wait(&k_full[b], par);
        ptx::mbar_wait(&s_empty[b], par ^ 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(sK + b * K_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_f16(tmem_base + b * 128, ptx::umma_desc_k_sw128(q_addr + k * 32), ptx::umma_desc_k_sw128(k_addr + k * 32), idesc_s,
                        (uint32_t)(k != 0));
        ptx::umma_commit(&k_empty[b]);
        ptx::umma_commit(&s_full[b]);
      };
      ptx::mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) issue_s(j + 1);
        const int b = j & 1;
        const uint32_t par = (uint32_t)((j >> 1) & 1);
        ptx::mbar_wait(&v_full[b], par);
        ptx::mbar_wait(&p_full[b], par);
        ptx::tc_fence_after();
        const uint32_t p_addr = ptx::smem_u32(sP + b * P_BYTES), v_addr = ptx::smem_u32(sV + b * V_BYTES);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_f16(tmem_O, ptx::umma_desc_k_sw128(p_addr + kb * (BQ * 128) + k * 32),
                          ptx::umma_desc_k_sw128(v_addr + kb * (D * 128) + k * 32), idesc_o, (uint32_t)((j | kb | k) != 0));
        ptx::umma_commit(&v_empty[b]);
        ptx::umma_commit(&p_empty[b]);
      }
      ptx::umma_commit(o_full);
    }
  } else {
    const int g = (warp - 2) >> 2;       // softmax group: 0 -> even key tiles, 1 -> odd key tiles
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;   // query row in the tile == TMEM lane
    const int t = q0 + r;
    const bool valid = t < p.Lq;
    const int limit = valid ? p.kv_off + p.seg.begin[seg_of(p.seg, t) + 1] : 0;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float bound = __expf(fminf(__ldg(p.scale_mul + h), 4.605170185988092f));
    const float c = p.log2e_scale, mc = bound * c;
    float l = 0.0f;
    uint8_t* prow = sP + g * P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t tmem_S = tmem_base + g * 128 + lane_addr;
    int n = 0;
    for (int j = g; j < nk; j += 2, ++n) {
      const uint32_t par = (uint32_t)(n & 1);
      ptx::mbar_wait(&s_full[g], par);
      ptx::mbar_wait(&p_empty[g], par ^ 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t s[32];
        ptx::tmem_ld_32x32(tmem_S + c0, s);
        ptx::tmem_ld_wait();
        const int kbase = j * BKV + c0;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float e0 = (kbase + i < limit) ? ex2(fmaf(__uint_as_float(s[i]), c, -mc)) : 0.0f;
          const float e1 = (kbase + i + 1 < limit) ? ex2(fmaf(__uint_as_float(s[i + 1]), c, -mc)) : 0.0f;
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
          l += __low2float(b2) + __high2float(b2);
          pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        uint8_t* blk = prow + (c0 >> 6) * (BQ * 128);
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(blk + (((chunk0 + q) ^ (r & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { ptx::mbar_arrive(&s_empty[g]); ptx::mbar_arrive(&p_full[g]); }
    }
    // combine the two groups' row sums, then each group writes one half of the head dimension
    sL[g * 128 + r] = l;
    named_bar_sync(1, 256);
    const float inv = 1.0f / (sL[r] + sL[128 + r]);
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    uint32_t o[32];
    ptx::tmem_ld_32x32(tmem_O + lane_addr + g * 32, o);
    ptx::tmem_ld_wait();
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)img * p.Lq + t) * p.C + h * D + g * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv),
                            pack_bf16x2(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv),
                            pack_bf16x2(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv),
                            pack_bf16x2(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

int launch_onepass(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, int imgs, int H, int Lq, int kv_off,
                   const int* seg_begin_host, int S, float scale, const float* scale_mul, __nv_bfloat16* out, cudaStream_t st) {
  Params p{};
  p.H = H; p.Lq = Lq; p.kv_off = kv_off; p.C = H * D;
  p.log2e_scale = scale * 1.4426950408889634f;
  p.scale_mul = scale_mul;
  p.out = out;
  p.seg.S = S;
  for (int j = 0; j <= S; ++j) p.seg.begin[j] = seg_begin_host[j];
  static bool attr_set = false;
  if (!attr_set) {
    SDVAR_CUDA(cudaFuncSetAttribute(attention_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  dim3 grid((Lq + BQ - 1) / BQ, H, imgs);
  attention_onepass_kernel<<<grid, kThreads, kSmemBytes, st>>>(tmQ, tmK, tmV, p);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

}  // namespace attn2
}  // namespace sdvar
