// gemm_tcgen05.cu -- K1: D = epilogue(A[M,K] @ W[N,K]^T) on the 5th-gen tensor cores.
//
// Replaces every F.linear on the path: mat_qkv (models/basic_var.py:93), proj (:119), fc1/fc2 (:52),
// ada_lin (:156), head (models/var.py:125).
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   TMA producer : cp.async.bulk.tensor 128x64 bf16 boxes of A and W, SWIZZLE_128B, into a
//                           kStages-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1   MMA issuer   : one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) from
//                           shared-memory descriptors into one of two TMEM accumulator stages; tcgen05.commit
//                           releases ring slots and publishes finished accumulators
//   warp 2   TMEM allocator (256 columns = 2 x 128-column fp32 accumulators)
//   warps 4-7 epilogue    : tcgen05.ld (lane = output row, registers = columns) -> fused epilogue -> global.
// The epilogue of tile i overlaps the main loop of tile i+1 through the double-buffered accumulator.
// Tiles are visited in groups of 16 M-blocks x all N-blocks so that the A panel of a group stays in L2.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "gemm_epilogue.cuh"
#include <stdlib.h>

namespace sdvar {

namespace gemm {
constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16;
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr int kThreads = 256;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int kTmemCols = kAccStages * BN;  // 256
constexpr int kGroupM = 16;
constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kStages * (A_BYTES + B_BYTES) + 256 /*barriers*/;

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
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, Epi ep) {
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
  const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int kblocks = K / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int mb, nb;
        tile_coords(t, num_m, num_n, mb, nb);
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          ptx::mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
          ptx::tma_load_2d(sA + (size_t)stage * A_BYTES, &tmA, &full[stage], kb * BK, mb * BM);
          ptx::tma_load_2d(sB + (size_t)stage * B_BYTES, &tmB, &full[stage], kb * BK, nb * BN);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
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
            ptx::umma_f16(d, ptx::umma_desc_k_sw128(a_addr + k * UMMA_K * 2), ptx::umma_desc_k_sw128(b_addr + k * UMMA_K * 2),
                          idesc, (uint32_t)((kb | k) != 0));
          ptx::umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull[as]);
        if (++as == kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;  // == warp % 4: TMEM lane quarter this warp may read
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int mb, nb;
      tile_coords(t, num_m, num_n, mb, nb);
      EpiPre pre;
      epilogue_prefetch<EPI>(pre, mb * BM + ew * 32 + lane, nb * BN, M, N, ep);
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
      epilogue_tile<EPI, BN>(taddr, mb * BM + ew * 32 + lane, nb * BN, M, N, ep, pre);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[as]);
      if (++as == kAccStages) { as = 0; aphase ^= 1; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

template <int EPI>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, const Epi& ep, cudaStream_t st) {
  SDVAR_SET_SMEM_ONCE(gemm_kernel<EPI>, kSmemBytes);
  const int sms = sm_count();
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < sms ? tiles : sms;
  gemm_kernel<EPI><<<grid, kThreads, kSmemBytes, st>>>(tmA, tmB, M, N, K, ep);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

}  // namespace gemm
}  // namespace sdvar

namespace sdvar {
namespace gemm2 {
int launch_pair(int epilogue, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const gemm::Epi& ep, cudaStream_t st);
}
// 0 = automatic, 1 = force the 1-CTA kernel, 2 = force the CTA-pair kernel (SDVAR_GEMM env, for A/B measurements)
static int gemm_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SDVAR_GEMM");
    mode = e ? atoi(e) : 0;
  }
  return mode;
}
}  // namespace sdvar

using namespace sdvar;

extern "C" int sdvar_gemm_bf16(const sdvar_bf16* A, int lda, const sdvar_bf16* W, int ldw, int M, int N, int K,
                               const sdvar_gemm_epilogue* e, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(A && W && e, "NULL argument");
  SDVAR_REQUIRE(M > 0 && N > 0 && K > 0, "bad shape M=%d N=%d K=%d", M, N, K);
  SDVAR_REQUIRE(K % gemm::BK == 0, "K=%d must be a multiple of 64", K);
  SDVAR_REQUIRE(N % 32 == 0, "N=%d must be a multiple of 32", N);
  SDVAR_REQUIRE(lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0, "lda/ldw must be >= K and multiples of 8");
  gemm::Epi ep{};
  ep.bias = e->bias;
  ep.out_f32 = e->out_f32;
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(e->out_bf16);
  ep.ldo = e->ldo;
  ep.gate = e->gate;
  ep.ld_gate = e->ld_gate;
  ep.tokens_per_img = e->tokens_per_img;
  ep.slot_map = e->slot_map;
  ep.q_out = reinterpret_cast<__nv_bfloat16*>(e->q_out);
  ep.k_cache = reinterpret_cast<__nv_bfloat16*>(e->k_cache);
  ep.vT_cache = reinterpret_cast<__nv_bfloat16*>(e->vT_cache);
  ep.scale_mul = e->scale_mul;
  ep.H = e->H; ep.Lq = e->Lq; ep.Lmax = e->Lmax; ep.Lmax_pad = e->Lmax_pad; ep.kv_off = e->kv_off; ep.l2norm = e->l2norm;
  ep.C = e->H * 64;
  SDVAR_REQUIRE(((uintptr_t)e->bias & 15) == 0, "bias must be 16-byte aligned");
  switch (e->epilogue) {
    case SDVAR_EPI_F32:
      SDVAR_REQUIRE(e->out_f32 && e->ldo >= N && e->ldo % 4 == 0 && ((uintptr_t)e->out_f32 & 15) == 0, "bad F32 output");
      break;
    case SDVAR_EPI_BF16:
    case SDVAR_EPI_GELU_BF16:
      SDVAR_REQUIRE(e->out_bf16 && e->ldo >= N && e->ldo % 8 == 0 && ((uintptr_t)e->out_bf16 & 15) == 0, "bad BF16 output");
      break;
    case SDVAR_EPI_RESID_F32:
      SDVAR_REQUIRE(e->out_f32 && e->ldo >= N && e->ldo % 4 == 0 && ((uintptr_t)e->out_f32 & 15) == 0, "bad RESID output");
      SDVAR_REQUIRE(e->gate && e->ld_gate % 4 == 0 && e->tokens_per_img > 0 && ((uintptr_t)e->gate & 15) == 0, "bad gate");
      break;
    case SDVAR_EPI_QKV:
      SDVAR_REQUIRE(e->q_out && e->k_cache && e->vT_cache && e->bias, "QKV epilogue needs q_out/k_cache/vT_cache/bias");
      SDVAR_REQUIRE(e->H > 0 && N == 3 * e->H * 64, "QKV: N=%d must equal 3*H*64 (H=%d)", N, e->H);
      SDVAR_REQUIRE(e->Lq > 0 && M % e->Lq == 0 && e->kv_off >= 0 && e->kv_off + e->Lq <= e->Lmax && e->Lmax <= e->Lmax_pad,
                    "QKV: bad cache geometry Lq=%d kv_off=%d Lmax=%d", e->Lq, e->kv_off, e->Lmax);
      SDVAR_REQUIRE(!e->l2norm || e->scale_mul, "QKV: scale_mul is NULL");
      SDVAR_REQUIRE(((uintptr_t)e->q_out & 15) == 0 && ((uintptr_t)e->k_cache & 15) == 0, "QKV: alignment");
      break;
    default:
      SDVAR_REQUIRE(false, "unknown epilogue %d", e->epilogue);
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(st, FAM_GEMM, 2.0 * M * (double)N * K);
  // CTA pairs (256x256 tiles) once there are enough tiles to fill 74 clusters; small M stays on 128x128 tiles
  const int mode = gemm_mode();
  const long long pair_tiles = (long long)((M + 255) / 256) * ((N + 255) / 256);
  if (mode == 2 || (mode == 0 && pair_tiles >= 74)) return gemm2::launch_pair(e->epilogue, A, lda, W, ldw, M, N, K, ep, st);
  CUtensorMap tmA, tmB;
  {
    const uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[1] = {(uint64_t)lda * 2};
    const uint32_t boxA[2] = {(uint32_t)gemm::BK, (uint32_t)gemm::BM};
    if (int rc = make_tmap_bf16(&tmA, A, 2, dimsA, strA, boxA)) return rc;
    const uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[1] = {(uint64_t)ldw * 2};
    const uint32_t boxB[2] = {(uint32_t)gemm::BK, (uint32_t)gemm::BN};
    if (int rc = make_tmap_bf16(&tmB, W, 2, dimsB, strB, boxB)) return rc;
  }
  switch (e->epilogue) {
    case SDVAR_EPI_F32: return gemm::launch<SDVAR_EPI_F32>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_BF16: return gemm::launch<SDVAR_EPI_BF16>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_GELU_BF16: return gemm::launch<SDVAR_EPI_GELU_BF16>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_RESID_F32: return gemm::launch<SDVAR_EPI_RESID_F32>(tmA, tmB, M, N, K, ep, st);
    default: return gemm::launch<SDVAR_EPI_QKV>(tmA, tmB, M, N, K, ep, st);
  }
}
