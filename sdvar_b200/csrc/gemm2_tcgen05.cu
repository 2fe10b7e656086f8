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
//       tempty[a] (count 8 = 4 warps x 2 CTAs)
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
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr int kThreads = 256;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2;   // 16 KiB + 16 KiB per CTA per stage
constexpr int kTmemCols = kAccStages * BN;                         // 512
constexpr int kGroupM = 8;                                         // in 256-row cluster tiles
constexpr size_t kSmemBytes = 1024 + (size_t)kStages * (A_BYTES + B_BYTES) + 256;
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
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)kStages * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * (A_BYTES + B_BYTES));
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kAccStages);

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
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&full[i], 2); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 8); }
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
    const int ew = warp - 4;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int mb, nb;
      tile_coords(t, num_m, num_n, mb, nb);
      const int row = (mb * 2 + (int)rank) * BM + ew * 32 + lane;
      EpiPre pre;
      gemm::epilogue_prefetch<EPI>(pre, row, nb * BN, M, N, ep);
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
      gemm::epilogue_tile<EPI, BN>(taddr, row, nb * BN, M, N, ep, pre);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[as]);
      if (++as == kAccStages) { as = 0; aphase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) tmem_dealloc2(tmem_base, kTmemCols);
}

template <int EPI>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epi& ep, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    SDVAR_CUDA(cudaFuncSetAttribute(gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  const int sms = sm_count();
  const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int clusters = tiles < sms / 2 ? tiles : sms / 2;
  gemm2_kernel<EPI><<<2 * clusters, kThreads, kSmemBytes, st>>>(tmA, tmB, M, N, K, ep);
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
  switch (epilogue) {
    case SDVAR_EPI_F32: return launch<SDVAR_EPI_F32>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_BF16: return launch<SDVAR_EPI_BF16>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_GELU_BF16: return launch<SDVAR_EPI_GELU_BF16>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_RESID_F32: return launch<SDVAR_EPI_RESID_F32>(tmA, tmB, M, N, K, ep, st);
    default: return launch<SDVAR_EPI_QKV>(tmA, tmB, M, N, K, ep, st);
  }
}

}  // namespace gemm2
}  // namespace sdvar
