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

struct Epi {
  const float* bias;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ldo;
  const float* gate;
  int ld_gate, tokens_per_img;
  __nv_bfloat16 *q_out, *k_cache, *vT_cache;
  const float* scale_mul;
  int H, Lq, Lmax, Lmax_pad, kv_off, l2norm, C;
};

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& mb, int& nb) {
  const int per_group = kGroupM * num_n;
  const int g = t / per_group;
  const int first_m = g * kGroupM;
  const int gsz = min(kGroupM, num_m - first_m);
  const int r = t - g * per_group;
  mb = first_m + r % gsz;
  nb = r / gsz;
}

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

template <int EPI>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, int row, int col_base, int M, int N, const Epi& ep) {
  if constexpr (EPI == SDVAR_EPI_QKV) {
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
        if (sect < 2) {
          __nv_bfloat16* dst = (sect == 0)
              ? ep.q_out + (((size_t)img * ep.H + h) * ep.Lq + t) * 64
              : ep.k_cache + (((size_t)img * ep.H + h) * ep.Lmax + ep.kv_off + t) * 64;
          uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            d[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                              pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
        } else {
          __nv_bfloat16* dst = ep.vT_cache + ((size_t)img * ep.H + h) * 64 * ep.Lmax_pad + ep.kv_off + t;
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
        } else if constexpr (EPI == SDVAR_EPI_RESID_F32) {
          const int img = row / ep.tokens_per_img;
          const float4* g4 = reinterpret_cast<const float4*>(ep.gate + (size_t)img * ep.ld_gate + col0);
          float4* d = reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * ep.ldo + col0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 g = __ldg(g4 + q);
            float4 o = d[q];
            o.x += v[4 * q] * g.x; o.y += v[4 * q + 1] * g.y; o.z += v[4 * q + 2] * g.z; o.w += v[4 * q + 3] * g.w;
            d[q] = o;
          }
        }
      }
    }
  }
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
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
      epilogue_tile<EPI>(taddr, mb * BM + ew * 32 + lane, nb * BN, M, N, ep);
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
  static bool attr_set = false;
  if (!attr_set) {
    SDVAR_CUDA(cudaFuncSetAttribute(gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < sms ? tiles : sms;
  gemm_kernel<EPI><<<grid, kThreads, kSmemBytes, st>>>(tmA, tmB, M, N, K, ep);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

}  // namespace gemm
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
  CUtensorMap tmA, tmB;
  {
    const uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[1] = {(uint64_t)lda * 2};
    const uint32_t boxA[2] = {(uint32_t)gemm::BK, (uint32_t)gemm::BM};
    if (int rc = make_tmap_bf16(&tmA, A, 2, dimsA, strA, boxA)) return rc;
    const uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[1] = {(uint64_t)ldw * 2};
    const uint32_t boxB[2] = {(uint32_t)gemm::BK, (uint32_t)gemm::BN};
    if (int rc = make_tmap_bf16(&tmB, W, 2, dimsB, strB, boxB)) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(st, FAM_GEMM, 2.0 * M * (double)N * K);
  switch (e->epilogue) {
    case SDVAR_EPI_F32: return gemm::launch<SDVAR_EPI_F32>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_BF16: return gemm::launch<SDVAR_EPI_BF16>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_GELU_BF16: return gemm::launch<SDVAR_EPI_GELU_BF16>(tmA, tmB, M, N, K, ep, st);
    case SDVAR_EPI_RESID_F32: return gemm::launch<SDVAR_EPI_RESID_F32>(tmA, tmB, M, N, K, ep, st);
    default: return gemm::launch<SDVAR_EPI_QKV>(tmA, tmB, M, N, K, ep, st);
  }
}
