// gemm2_tcgen05.cu -- K1, CTA-pair variant: tcgen05.mma.cta_group::2 on a 256 x 256 output tile per cluster of 2 CTAs.
//
// Why: with 128x128 tiles every CTA streams 32 KiB of operands per 64-wide K block for 128x128x64 MACs; at ~1 PFLOP/s that is
// ~15 TB/s of L2->SM traffic, beyond what the L2 delivers, and the 1-CTA kernel saturates near 65% of the cuBLAS rate.  In the
// pair, each CTA loads its own 128 rows of A and HALF of the 256-row W tile per K block (32 KiB per CTA for 128x256x64 MACs):
// half the L2 traffic and half the shared-memory fill per FLOP.  The leader CTA's single MMA lane issues M=256,N=256,K=16
// instructions that read A from both CTAs' shared memory and accumulate into 256 TMEM columns of each CTA (its 128 rows).
//
// Protocol (per CTA: full/empty[kStages], tfull/tempty[2] mbarriers; only the leader's full and tempty are used):
//   producer warp (both CTAs): wait own empty[s] -> arrive.expect_tx(own 32 KiB) on the LEADER's full[s] -> two
//       cp.async.bulk.tensor.cta_group::2 loads into own smem that complete_tx on the leader's full[s]      (count 2)
//   MMA lane (leader): wait full[s] -> 4 x tcgen05.mma.cta_group::2 -> tcgen05.commit multicast to empty[s] of both CTAs;
//       after the last K block: commit multicast to tfull[a] of both CTAs
//   epilogue warps (both CTAs): wait own tfull[a] -> tcgen05.ld own 128 rows -> fused epilogue -> arrive on the leader's
//       tempty[a] (count = epilogue warps x 2 CTAs)
//
// fp32 epilogues (RESID_F32: x += gate * (acc + bias) in place; F32: logits) go through shared memory: each epilogue warp
// owns a ring of 32-row x 32-column fp32 slots (4 KiB, SWIZZLE_128B).  For RESID the residual chunk is TMA-LOADED into the
// slot kEpiAhead chunks ahead (across tile boundaries, so the loads overlap the main loop), combined in place with the
// accumulator, and TMA-STORED back; F32 only stages and stores.  A thread-per-row epilogue issuing 16-byte global accesses
// to 32 different lines per instruction kept the K = C proj GEMM at ~55 % of the cuBLAS rate.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "gemm_epilogue.cuh"

namespace sdvar {
namespace gemm2 {

using gemm::Epi;
using gemm::EpiPre;

constexpr int BM = 128;          // rows per CTA (256 per cluster tile)
constexpr int BN = 256;          // columns per cluster tile; each CTA loads BN/2 rows of W
constexpr int BK = 64, UMMA_K = 16;
constexpr int kAccStages = 2;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2;   // 16 KiB + 16 KiB per CTA per stage
constexpr int kTmemCols = kAccStages * BN;                         // 512
constexpr int kGroupM = 8;                                         // in 256-row cluster tiles
constexpr int kEpiSlots = 4, kEpiSlotBytes = 32 * 128;             // per epilogue warp: ring of 32 x 32 fp32 chunks
constexpr int kEpiAhead = 2;                                       // residual chunks loaded ahead; kEpiSlots - kEpiAhead stores may still be reading
constexpr int kChunks = BN / 32;
constexpr int kBarBytes = 512;
template <int EPI>
struct Cfg {
  static constexpr bool staged = EPI == SDVAR_EPI_RESID_F32 || EPI == SDVAR_EPI_F32;
  static constexpr int stages = staged ? 5 : 6;
  // warps 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, then the epilogue: 4 warps (one per TMEM lane quarter) when staged,
  // otherwise 8 (two column halves x four lane quarters)
  static constexpr int threads = staged ? 256 : 384;
  static constexpr int epi_warps = threads / 32 - 4;
  static constexpr size_t ring_bytes = staged ? (size_t)4 * kEpiSlots * kEpiSlotBytes : 0;
  static constexpr size_t smem_bytes = 1024 + (size_t)stages * (A_BYTES + B_BYTES) + ring_bytes + kBarBytes;
};
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;                        // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive(+expect_tx) on the barrier at the same offset in the LEADER CTA
__device__ __forceinline__ void mbar_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(ptx::smem_u32(bar) & kPeerMask), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ptx::smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   ptx::smem_u32(bar)),
               "h"(mask)
               : "memory");
}

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& mb, int& nb) {
  const int per_group = kGroupM * num_n;
  const int g = t / per_group;
  const int first_m = g * kGroupM;
  const int gsz = min(kGroupM, num_m - first_m);
  const int r = t - g * per_group;
  mb = first_m + r % gsz;
  nb = r / gsz;
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<EPI>::threads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmX, int M, int N, int K, Epi ep) {
  constexpr int kStages = Cfg<EPI>::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)kStages * A_BYTES;
  uint8_t* sX = smem + (size_t)kStages * (A_BYTES + B_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(sX + Cfg<EPI>::ring_bytes);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + kAccStages;
  uint64_t* xfull = tempty + kAccStages;                       // [4 warps][kEpiSlots], staged epilogues only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull + 4 * kEpiSlots);
  static_assert((2 * 6 + 2 * kAccStages + 4 * kEpiSlots) * 8 + 4 <= kBarBytes, "barrier region");

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_m = (M + 2 * BM - 1) / (2 * BM), num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int kblocks = K / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    if constexpr (Cfg<EPI>::staged) ptx::prefetch_tmap(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&full[i], 2); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 2 * Cfg<EPI>::epi_warps); }
    for (int i = 0; i < 4 * kEpiSlots; ++i) ptx::mbar_init(&xfull[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mb, nb;
        tile_coords(t, num_m, num_n, mb, nb);
        const int row0 = (mb * 2 + (int)rank) * BM;            // this CTA's 128 rows of A
        const int col0 = nb * BN + (int)rank * (BN / 2);       // this CTA's half of the W tile
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx_leader(&full[stage], A_BYTES + B_BYTES);
          tma_load_2d_pair(sA + (size_t)stage * A_BYTES, &tmA, &full[stage], kb * BK, row0);
          tma_load_2d_pair(sB + (size_t)stage * B_BYTES, &tmB, &full[stage], kb * BK, col0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(sA + (size_t)stage * A_BYTES);
          const uint32_t b_addr = ptx::smem_u32(sB + (size_t)stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma2_f16(d, ptx::umma_desc_k_sw128(a_addr + k * UMMA_K * 2), ptx::umma_desc_k_sw128(b_addr + k * UMMA_K * 2), idesc,
                      (uint32_t)((kb | k) != 0));
          umma2_commit_mc(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(&tfull[as]);
        if (++as == kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    if constexpr (Cfg<EPI>::staged) {
      constexpr bool kResid = EPI == SDVAR_EPI_RESID_F32;
      const int ew = warp & 3;                               // TMEM lane quarter = rows ew*32 .. +32 of this CTA's 128
      uint8_t* ring = sX + (size_t)ew * kEpiSlots * kEpiSlotBytes;
      uint64_t* xf = xfull + ew * kEpiSlots;
      const int my_tiles = cluster_id < num_tiles ? (num_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
      const int total = my_tiles * kChunks;                  // 32-column chunks this warp walks, across all its tiles
      auto chunk_coords = [&](int gi, int& row0, int& col0) {
        int mb, nb;
        tile_coords(cluster_id + (gi / kChunks) * num_clusters, num_m, num_n, mb, nb);
        row0 = (mb * 2 + (int)rank) * BM + ew * 32;
        col0 = nb * BN + (gi % kChunks) * 32;
      };
      int issued = 0;
      auto issue_load = [&]() {                              // lane 0 only
        int r0, c0;
        chunk_coords(issued, r0, c0);
        const int slot = issued % kEpiSlots;
        ptx::mbar_expect_tx(&xf[slot], kEpiSlotBytes);
        ptx::tma_load_2d(ring + slot * kEpiSlotBytes, &tmX, &xf[slot], c0, r0);   // out-of-range parts are zero-filled
        ++issued;
      };
      if (kResid && lane == 0)
        while (issued < total && issued < kEpiAhead) issue_load();
      int as = 0;
      uint32_t aphase = 0;
      int row0 = 0, colb = 0;
      for (int gi = 0; gi < total; ++gi) {
        const int c = gi % kChunks, slot = gi % kEpiSlots;
        if (c == 0) {
          chunk_coords(gi, row0, colb);
          ptx::mbar_wait(&tfull[as], aphase);
          ptx::tc_fence_after();
        }
        const int col0 = colb + c * 32;
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN + c * 32), r);
        if constexpr (kResid) {
          ptx::mbar_wait(&xf[slot], (uint32_t)(gi / kEpiSlots) & 1u);
        } else {
          if (lane == 0) ptx::bulk_wait_group_read<kEpiSlots - 1>();   // the store that last used this slot has read it
          __syncwarp();
        }
        ptx::tmem_ld_wait();
        if (col0 < N) {
          const int row = row0 + lane;
          uint8_t* srow = ring + slot * kEpiSlotBytes + lane * 128;
          const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
          [[maybe_unused]] const float4* g4 = nullptr;
          if constexpr (kResid)   // one gate row per image: warp-uniform most of the time, L1-resident
          {
            const int gimg = row < M ? row / ep.tokens_per_img : 0;
            g4 = reinterpret_cast<const float4*>(ep.gate + (size_t)(ep.slot_map != nullptr ? __ldg(ep.slot_map + gimg) : gimg) * ep.ld_gate + col0);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4* p = reinterpret_cast<float4*>(srow + ((q ^ (lane & 7)) << 4));   // SWIZZLE_128B: 16-byte unit ^ (row & 7)
            const float4 bb = ep.bias != nullptr ? __ldg(b4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            if constexpr (kResid) {
              const float4 g = __ldg(g4 + q);
              o = *p;
              o.x += (__uint_as_float(r[4 * q]) + bb.x) * g.x;
              o.y += (__uint_as_float(r[4 * q + 1]) + bb.y) * g.y;
              o.z += (__uint_as_float(r[4 * q + 2]) + bb.z) * g.z;
              o.w += (__uint_as_float(r[4 * q + 3]) + bb.w) * g.w;
            } else {
              o = make_float4(__uint_as_float(r[4 * q]) + bb.x, __uint_as_float(r[4 * q + 1]) + bb.y,
                              __uint_as_float(r[4 * q + 2]) + bb.z, __uint_as_float(r[4 * q + 3]) + bb.w);
            }
            *p = o;
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (col0 < N && row0 < M) ptx::tma_store_2d(&tmX, ring + slot * kEpiSlotBytes, col0, row0);   // clipped at M, N
          ptx::bulk_commit_group();
          if (kResid && issued < total) {
            ptx::bulk_wait_group_read<kEpiSlots - kEpiAhead>();   // the store that last used the slot being refilled has read it
            issue_load();
          }
        }
        if (c == kChunks - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty[as]);
          if (++as == kAccStages) { as = 0; aphase ^= 1; }
        }
      }
      if (lane == 0) ptx::bulk_wait_group<0>();
    } else {
      const int ew = warp & 3;                 // TMEM lane quarter this warp may read
      const int ch = (warp - 4) >> 2;          // column half of the tile
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mb, nb;
        tile_coords(t, num_m, num_n, mb, nb);
        const int row = (mb * 2 + (int)rank) * BM + ew * 32 + lane;
        EpiPre pre;
        ptx::mbar_wait(&tfull[as], aphase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN + ch * (BN / 2));
        gemm::epilogue_tile<EPI, BN / 2>(taddr, row, nb * BN + ch * (BN / 2), M, N, ep, pre);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[as]);
        if (++as == kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  }
  ptx::tc_fence_before();
  cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) tmem_dealloc2(tmem_base, kTmemCols);
}

template <int EPI>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmX, int M, int N, int K, const Epi& ep, cudaStream_t st) {
  SDVAR_SET_SMEM_ONCE(gemm2_kernel<EPI>, Cfg<EPI>::smem_bytes);
  const int sms = sm_count();
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int clusters = tiles < sms / 2 ? tiles : sms / 2;
  gemm2_kernel<EPI><<<2 * clusters, Cfg<EPI>::threads, Cfg<EPI>::smem_bytes, st>>>(tmA, tmB, tmX, M, N, K, ep);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

// explicit entry used by gemm_tcgen05.cu's dispatcher
int launch_pair(int epilogue, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const Epi& ep, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  const uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[1] = {(uint64_t)lda * 2};
  const uint32_t boxA[2] = {(uint32_t)BK, (uint32_t)BM};
  if (int rc = make_tmap_bf16(&tmA, A, 2, dimsA, strA, boxA)) return rc;
  const uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[1] = {(uint64_t)ldw * 2};
  const uint32_t boxB[2] = {(uint32_t)BK, (uint32_t)(BN / 2)};
  if (int rc = make_tmap_bf16(&tmB, W, 2, dimsB, strB, boxB)) return rc;
  CUtensorMap tmX = tmA;   // unused unless the epilogue is staged
  if (epilogue == SDVAR_EPI_F32 || epilogue == SDVAR_EPI_RESID_F32) {
    SDVAR_REQUIRE(N % 32 == 0 && ep.ldo % 4 == 0, "fp32 epilogue needs N %% 32 == 0 and ldo %% 4 == 0 (N=%d ldo=%d)", N, ep.ldo);
    const uint64_t dimsX[2] = {(uint64_t)N, (uint64_t)M}, strX[1] = {(uint64_t)ep.ldo * 4};
    const uint32_t boxX[2] = {32u, 32u};
    if (int rc = make_tmap_f32(&tmX, ep.out_f32, 2, dimsX, strX, boxX)) return rc;
  }
  switch (epilogue) {
    case SDVAR_EPI_F32: return launch<SDVAR_EPI_F32>(tmA, tmB, tmX, M, N, K, ep, st);
    case SDVAR_EPI_BF16: return launch<SDVAR_EPI_BF16>(tmA, tmB, tmX, M, N, K, ep, st);
    case SDVAR_EPI_GELU_BF16: return launch<SDVAR_EPI_GELU_BF16>(tmA, tmB, tmX, M, N, K, ep, st);
    case SDVAR_EPI_RESID_F32: return launch<SDVAR_EPI_RESID_F32>(tmA, tmB, tmX, M, N, K, ep, st);
    default: return launch<SDVAR_EPI_QKV>(tmA, tmB, tmX, M, N, K, ep, st);
  }
}

}  // namespace gemm2
}  // namespace sdvar
