// attention_tcgen05.cu -- K2: attention of the current pass's queries over the preallocated KV cache.
//
// Replaces models/basic_var.py:107-117 (torch.cat KV cache + scaled_dot_product_attention).  The cache is a
// fixed ring written in place by the QKV GEMM epilogue (K as (img,h,pos,64), V transposed as (img,h,64,pos)),
// so "append" costs nothing and rollback after a rejected stage is just a smaller kv length.
//
// One CTA = 128 query rows of one (image, head).  Both contractions run on tcgen05 with K-major SWIZZLE_128B
// operands fed by 3-D TMA boxes:  S = Q K^T  (M=128, N=128 keys, K=64) into TMEM columns [0,128);
// O += P V  (M=128, N=64, K=128 keys) into TMEM columns [128,192), P being written to shared memory as bf16
// by the softmax warps in the UMMA swizzle.  Softmax is two-pass (pass 0: row maxima, pass 1: exp / PV), which
// costs a second QK^T sweep (attention is ~3.5% of the model FLOPs) but needs no accumulator rescaling.
// Block-causal masking inside a verify window comes from the stage table (query of stage j sees keys
// < kv_off + end_j, models/var.py:108-113); incremental decode is the S=1 case.
//
// warp 0: TMA producer | warp 1: TMEM alloc + MMA issuer | warps 2-5: softmax / epilogue (TMEM lane quarter
// = warp % 4).
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sdvar {
namespace attn {

constexpr int BQ = 128, BKV = 128, D = 64;
constexpr int kThreads = 192;
constexpr int Q_BYTES = BQ * D * 2;        // 16 KiB
constexpr int K_BYTES = BKV * D * 2;       // 16 KiB per stage
constexpr int V_BYTES = D * BKV * 2;       // 16 KiB (two 64x64 boxes)
constexpr int P_BYTES = BQ * BKV * 2;      // 32 KiB (two 128x64 K-blocks)
constexpr int kKStages = 2;
constexpr int kTmemCols = 256;             // S: [0,128)  O: [128,192)
constexpr size_t kSmemBytes = 1024 + Q_BYTES + kKStages * K_BYTES + V_BYTES + P_BYTES + 256;

struct Params {
  int H, Lq, kv_off, C;
  float scale_log2e;  // softmax scale * log2(e)
  __nv_bfloat16* out;
  const int* slot_map;  // pass-image -> KV-cache slot (nullptr: identity)
  SegTable seg;
};

__global__ void __launch_bounds__(kThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + kKStages * K_BYTES;
  uint8_t* sP = sV + V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, img = blockIdx.z;
  const int bh = img * p.H + h;
  const int kvbh = (p.slot_map != nullptr ? __ldg(p.slot_map + img) : img) * p.H + h;   // (cache slot, head) of K / V
  const int q0 = qt * BQ;
  // keys visible to the last valid query row of this tile bound the key loop
  const int t_last = min(q0 + BQ, p.Lq) - 1;
  const int max_limit = p.kv_off + p.seg.begin[seg_of(p.seg, t_last) + 1];
  const int nk = (max_limit + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kKStages; ++i) { ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1); }
    ptx::mbar_init(v_full, 1);
    ptx::mbar_init(v_empty, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 4);
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(p_empty, 1);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(q_full, Q_BYTES);
      ptx::tma_load_3d(sQ, &tmQ, q_full, 0, q0, bh);
      for (int it = 0; it < 2 * nk; ++it) {
        const int j = it < nk ? it : it - nk;
        const int ks = it & 1;
        ptx::mbar_wait(&k_empty[ks], ((it >> 1) & 1) ^ 1);
        ptx::mbar_expect_tx(&k_full[ks], K_BYTES);
        ptx::tma_load_3d(sK + ks * K_BYTES, &tmK, &k_full[ks], 0, j * BKV, kvbh);
        if (it >= nk) {
          ptx::mbar_wait(v_empty, (j & 1) ^ 1);
          ptx::mbar_expect_tx(v_full, V_BYTES);
          ptx::tma_load_3d(sV, &tmV, v_full, j * BKV, 0, kvbh);
          ptx::tma_load_3d(sV + V_BYTES / 2, &tmV, v_full, j * BKV + 64, 0, kvbh);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, D);
      const uint32_t q_addr = ptx::smem_u32(sQ), p_addr = ptx::smem_u32(sP), v_addr = ptx::smem_u32(sV);
      ptx::mbar_wait(q_full, 0);
      for (int it = 0; it < 2 * nk; ++it) {
        const int j = it < nk ? it : it - nk;
        const int ks = it & 1;
        ptx::mbar_wait(&k_full[ks], (it >> 1) & 1);
        ptx::mbar_wait(s_empty, (it & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(sK + ks * K_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_f16(tmem_S, ptx::umma_desc_k_sw128(q_addr + k * 32), ptx::umma_desc_k_sw128(k_addr + k * 32), idesc_s,
                        (uint32_t)(k != 0));
        ptx::umma_commit(&k_empty[ks]);
        ptx::umma_commit(s_full);
        if (it >= nk) {
          ptx::mbar_wait(v_full, j & 1);
          ptx::mbar_wait(p_full, j & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16(tmem_O, ptx::umma_desc_k_sw128(p_addr + kb * (BQ * 128) + k * 32),
                            ptx::umma_desc_k_sw128(v_addr + kb * (D * 128) + k * 32), idesc_o,
                            (uint32_t)((j | kb | k) != 0));
          ptx::umma_commit(v_empty);
          ptx::umma_commit(p_empty);
        }
      }
      ptx::umma_commit(o_full);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int t = q0 + r;
    const bool valid = t < p.Lq;
    const int limit = valid ? p.kv_off + p.seg.begin[seg_of(p.seg, t) + 1] : 0;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float m = -INFINITY;
    // ---- pass 0: row maxima ----
    for (int it = 0; it < nk; ++it) {
      ptx::mbar_wait(s_full, it & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t s[32];
        ptx::tmem_ld_32x32(tmem_S + lane_addr + c0, s);
        ptx::tmem_ld_wait();
        const int kbase = it * BKV + c0;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (kbase + c < limit) m = fmaxf(m, __uint_as_float(s[c]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_empty);
    }
    const float mc = (m == -INFINITY) ? 0.0f : m * p.scale_log2e;
    float l = 0.0f;
    uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
    // ---- pass 1: P = exp(S - m) as bf16 in the UMMA swizzle, O += P V ----
    for (int j = 0; j < nk; ++j) {
      const int it = nk + j;
      ptx::mbar_wait(s_full, it & 1);
      ptx::mbar_wait(p_empty, (j & 1) ^ 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t s[32];
        ptx::tmem_ld_32x32(tmem_S + lane_addr + c0, s);
        ptx::tmem_ld_wait();
        const int kbase = j * BKV + c0;
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float e0 = (kbase + c < limit) ? exp2f(__uint_as_float(s[c]) * p.scale_log2e - mc) : 0.0f;
          const float e1 = (kbase + c + 1 < limit) ? exp2f(__uint_as_float(s[c + 1]) * p.scale_log2e - mc) : 0.0f;
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
          l += __low2float(b2) + __high2float(b2);
          pk[c >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
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
      if (lane == 0) { ptx::mbar_arrive(s_empty); ptx::mbar_arrive(p_full); }
    }
    // ---- epilogue: O / l -> bf16 ----
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    uint32_t o0[32], o1[32];
    ptx::tmem_ld_32x32(tmem_O + lane_addr, o0);
    ptx::tmem_ld_32x32(tmem_O + lane_addr + 32, o1);
    ptx::tmem_ld_wait();
    if (valid) {
      const float inv = 1.0f / l;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)img * p.Lq + t) * p.C + h * D);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(pack_bf16x2(__uint_as_float(o0[8 * q]) * inv, __uint_as_float(o0[8 * q + 1]) * inv),
                            pack_bf16x2(__uint_as_float(o0[8 * q + 2]) * inv, __uint_as_float(o0[8 * q + 3]) * inv),
                            pack_bf16x2(__uint_as_float(o0[8 * q + 4]) * inv, __uint_as_float(o0[8 * q + 5]) * inv),
                            pack_bf16x2(__uint_as_float(o0[8 * q + 6]) * inv, __uint_as_float(o0[8 * q + 7]) * inv));
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[4 + q] = make_uint4(pack_bf16x2(__uint_as_float(o1[8 * q]) * inv, __uint_as_float(o1[8 * q + 1]) * inv),
                                pack_bf16x2(__uint_as_float(o1[8 * q + 2]) * inv, __uint_as_float(o1[8 * q + 3]) * inv),
                                pack_bf16x2(__uint_as_float(o1[8 * q + 4]) * inv, __uint_as_float(o1[8 * q + 5]) * inv),
                                pack_bf16x2(__uint_as_float(o1[8 * q + 6]) * inv, __uint_as_float(o1[8 * q + 7]) * inv));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace attn
}  // namespace sdvar

namespace sdvar {
namespace attn2 {
int launch_onepass(const void* q, const void* k_cache, const void* vT_cache, int imgs, int H, int Lq, int Lmax, int Lmax_pad, int kv_off,
                   const int* seg_begin_host, int S, float scale, const float* scale_mul, const int* slot_map, int cache_slots,
                   __nv_bfloat16* out, cudaStream_t st);
}
}  // namespace sdvar

using namespace sdvar;

extern "C" int sdvar_attention(const sdvar_bf16* q, const sdvar_bf16* k_cache, const sdvar_bf16* vT_cache, int imgs, int H,
                               int Lq, int Lmax, int Lmax_pad, int kv_off, const int* seg_begin_host, int S, float scale,
                               const float* logit_bound_log, const int* slot_map, int cache_slots, sdvar_bf16* out, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(q && k_cache && vT_cache && out, "NULL argument");
  SDVAR_REQUIRE(imgs > 0 && H > 0 && Lq > 0 && kv_off >= 0 && kv_off + Lq <= Lmax && Lmax <= Lmax_pad, "bad geometry");
  SDVAR_REQUIRE(Lmax_pad % 8 == 0, "Lmax_pad=%d must be a multiple of 8 (TMA stride)", Lmax_pad);
  SDVAR_REQUIRE(S >= 1 && S <= SDVAR_MAX_SEG && seg_begin_host && seg_begin_host[0] == 0 && seg_begin_host[S] == Lq,
                "segment table must cover [0,Lq)");
  SDVAR_REQUIRE(((uintptr_t)out & 15) == 0, "out alignment");
  if (cache_slots <= 0) cache_slots = imgs;
  SDVAR_REQUIRE(slot_map != nullptr || cache_slots == imgs, "cache_slots=%d without a slot map (imgs=%d)", cache_slots, imgs);
  attn::Params p{};
  p.H = H; p.Lq = Lq; p.kv_off = kv_off; p.C = H * attn::D;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.slot_map = slot_map;
  p.seg.S = S;
  for (int j = 0; j <= S; ++j) p.seg.begin[j] = seg_begin_host[j];
  double visible = 0;  // sum over query rows of visible keys
  for (int j = 0; j < S; ++j) visible += (double)(seg_begin_host[j + 1] - seg_begin_host[j]) * (kv_off + seg_begin_host[j + 1]);
  ProfileScope prof((cudaStream_t)stream, FAM_ATTN, 4.0 * 64.0 * visible * imgs * H);
  if (logit_bound_log != nullptr)
    return attn2::launch_onepass(q, k_cache, vT_cache, imgs, H, Lq, Lmax, Lmax_pad, kv_off, seg_begin_host, S, scale, logit_bound_log,
                                 slot_map, cache_slots, reinterpret_cast<__nv_bfloat16*>(out), (cudaStream_t)stream);
  CUtensorMap tmQ, tmK, tmV;
  {
    const uint64_t dq[3] = {64, (uint64_t)Lq, (uint64_t)imgs * H}, sq[2] = {128, (uint64_t)Lq * 128};
    const uint32_t bq[3] = {64, (uint32_t)attn::BQ, 1};
    if (int rc = make_tmap_bf16(&tmQ, q, 3, dq, sq, bq)) return rc;
    const uint64_t dk[3] = {64, (uint64_t)Lmax, (uint64_t)cache_slots * H}, sk[2] = {128, (uint64_t)Lmax * 128};
    const uint32_t bk[3] = {64, (uint32_t)attn::BKV, 1};
    if (int rc = make_tmap_bf16(&tmK, k_cache, 3, dk, sk, bk)) return rc;
    const uint64_t dv[3] = {(uint64_t)Lmax_pad, 64, (uint64_t)cache_slots * H}, sv[2] = {(uint64_t)Lmax_pad * 2, (uint64_t)Lmax_pad * 128};
    const uint32_t bv[3] = {64, 64, 1};
    if (int rc = make_tmap_bf16(&tmV, vT_cache, 3, dv, sv, bv)) return rc;
  }
  SDVAR_SET_SMEM_ONCE(attn::attention_kernel, attn::kSmemBytes);
  dim3 grid((Lq + attn::BQ - 1) / attn::BQ, H, imgs);
  attn::attention_kernel<<<grid, attn::kThreads, attn::kSmemBytes, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}
