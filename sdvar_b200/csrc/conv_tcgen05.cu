// conv_tcgen05.cu -- decoder convolutions as an implicit GEMM on tcgen05, channels-last bf16.
//
// Replaces nn.Conv2d(k=3, padding=1) / nn.Conv2d(k=1) of the VQVAE decoder (reference models/basic_vae.py:18-61, 163-226,
// models/vqvae.py:62-63).  out[p, co] = sum_{tap, ci} x[p + off(tap), ci] * w[tap, co, ci]  (+ bias[co]) (+ res[p, co]).
//
//   M = N*H*W output pixels, N = Cout, K = taps * Cin.  No im2col: the A operand of K block (tap, 32-channel chunk) is ONE
//   4-D TMA box over the activation (C, W, H, N) shifted by the tap offset -- out-of-image coordinates are zero-filled by the
//   TMA unit, which IS the convolution's zero padding.  Cin of the decoder is 32/160/320/640: multiples of 32, not of 64, so
//   the K blocks are 32 channels wide and the tiles use SWIZZLE_64B (64-byte rows, 8-row groups of 512 bytes).
//   A CTA pair owns 256 consecutive pixels x BN output channels (tcgen05.mma.cta_group::2, M=256): each CTA loads its own
//   128 pixels and HALF of the weight tile per K block, accumulators (2 x BN fp32 columns) live in tensor memory so the
//   epilogue of tile i overlaps the main loop of tile i+1.  Cout of the decoder is 160/320/640 = 1/2/4 tiles of BN = 160.
//   Epilogue (8 warps per CTA): tcgen05.ld -> + bias -> + skip connection (bf16) -> bf16 channels-last, or, for conv_out,
//   clamp -> fp32 NCHW image.  Same barrier protocol as gemm2_tcgen05.cu.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sdvar {
namespace conv {

constexpr int BM = 128;                 // pixels per CTA (256 per cluster tile)
constexpr int KC = 32, UMMA_K = 16;     // channels per K block
constexpr int kAccStages = 2;
constexpr int A_BYTES = BM * KC * 2;    // 8 KiB
constexpr int kThreads = 384;           // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int kEpiWarps = 8;
constexpr int kBarBytes = 512;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;

template <int BN>
struct Cfg {
  static constexpr int b_bytes = (BN / 2) * KC * 2;
  static constexpr int b_stride = (b_bytes + 1023) & ~1023;
  static constexpr int stage_bytes = A_BYTES + b_stride;
  static constexpr int stages = (200 * 1024 / stage_bytes) > 16 ? 16 : (200 * 1024 / stage_bytes);
  static constexpr int tmem_cols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  static constexpr size_t smem_bytes = 1024 + (size_t)stages * stage_bytes + kBarBytes;
};

struct Params {
  int Nimg, H, W, Cin, Cout, taps;      // taps: 9 (3x3, padding 1), 1, or 4 (one parity class of upsample2x + 3x3, see below)
  int tw, dy0, dx0, wtap0;              // tap t reads the input at (y + dy0 + t / tw, x + dx0 + t % tw) with weight matrix wtap0 + t
  int wtaps;                            // weight matrices in w_packed (9, 1, or 16 for the four parity classes)
  int up, pa, pb;                       // up: the output pixel of input-grid pixel (n, Y, X) is (n, 2Y + pa, 2X + pb) of a (2H, 2W) image
  int BW, BH;                           // pixel box of one CTA: BW x BH x (128 / (BW*BH)) images
  long long M;                          // Nimg * H * W
  const float* bias;                    // nullable
  const __nv_bfloat16* res;             // nullable, (M, Cout)
  __nv_bfloat16* out_bf16;              // (M, Cout) channels-last, or
  float* out_f32_nchw;                  // (Nimg, Cout, H, W) with clamp
  float lo, hi;
};

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
__device__ __forceinline__ void mbar_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(ptx::smem_u32(bar) & kPeerMask), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ptx::smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
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
// K-major operand tile written by TMA with SWIZZLE_64B: rows of 64 bytes (32 bf16), 8-row groups 512 bytes apart;
// descriptor version 1, layout type SWIZZLE_64B = 4 (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// rows of ROWB = 64 (SWIZZLE_64B) or 128 bytes (SWIZZLE_128B), 8-row groups 8*ROWB apart.  The start address may be ANY row of a
// TMA-written tile (base_offset stays 0): the tensor core swizzles on the absolute shared-memory address bits, exactly like the
// TMA unit that wrote the tile (checked on a B200 with tools/ubench/umma_shift.cu for every row shift 0..9, both swizzles).
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc_k_rows(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROWB) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;
  return d;
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Epilogue of both kernels (8 warps per CTA: TMEM lane quarter x column half): tcgen05.ld -> + bias -> + skip connection ->
// bf16 channels-last, or clamp -> fp32 NCHW.  A thread owns one output pixel; its 16-column chunks are 32-byte stores.
template <int BN>
__device__ __forceinline__ void epilogue_loop(const Params& p, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty, int warp, int lane,
                                              uint32_t rank, int cluster_id, int num_clusters, long long num_tiles, int num_n, int HW) {
  const int ew = warp & 3;                    // TMEM lane quarter this warp may read
  const int half = (warp - 4) >> 2;           // column half of the tile
  constexpr int nch = BN / 16, nch0 = (nch + 1) / 2;
  const int cbeg = half == 0 ? 0 : nch0, cend = half == 0 ? nch0 : nch;
  int as = 0;
  uint32_t aphase = 0;
  for (long long t = cluster_id; t < num_tiles; t += num_clusters) {
    const long long mb = t / num_n;
    const int nb = (int)(t - mb * num_n);
    const long long row = (mb * 2 + rank) * BM + ew * 32 + lane;
    const bool rv = row < p.M;
    long long orow = row;               // output pixel (channels-last row) this thread writes
    if (p.up && rv) {
      const long long n = row / HW;
      const int rem = (int)(row - n * HW), Y = rem / p.W, X = rem - Y * p.W;
      orow = (n * (2 * p.H) + 2 * Y + p.pa) * (2 * p.W) + 2 * X + p.pb;
    }
    ptx::mbar_wait(&tfull[as], aphase);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(as * BN);
    for (int c = cbeg; c < cend; ++c) {
      const int col = nb * BN + c * 16;
      uint32_t r[16];
      tmem_ld_32x16(taddr + c * 16, r);
      uint4 rs[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      const bool cv0 = rv && col < p.Cout, cv1 = rv && col + 8 < p.Cout;
      if (p.res != nullptr) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + row * p.Cout + col);
        if (cv0) rs[0] = __ldg(rp);
        if (cv1) rs[1] = __ldg(rp + 1);
      }
      ptx::tmem_ld_wait();
      if (cv0) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
          if (col + 16 <= p.Cout && (p.Cout & 3) == 0) {      // whole chunk in range: four 16-byte loads (warp-uniform, L1-resident)
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bb = __ldg(b4 + q);
              v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col + j < p.Cout) v[j] += __ldg(p.bias + col + j);
          }
        }
        if (p.out_bf16 != nullptr) {
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(rs);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[2 * j] += __uint_as_float(rw[j] << 16);
            v[2 * j + 1] += __uint_as_float(rw[j] & 0xFFFF0000u);
          }
          uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + orow * p.Cout + col);
          op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          if (cv1)
            op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        } else {
          const long long n = row / HW;
          const long long pix = row - n * HW;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (col + j < p.Cout) p.out_f32_nchw[(n * p.Cout + col + j) * HW + pix] = fminf(fmaxf(v[j], p.lo), p.hi);
        }
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_leader(&tempty[as]);
    if (++as == kAccStages) { as = 0; aphase ^= 1; }
  }
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
  using C = Cfg<BN>;
  constexpr int kStages = C::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * C::stage_bytes);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kAccStages);
  static_assert((2 * 16 + 2 * kAccStages) * 8 + 4 <= kBarBytes, "barrier region");

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_n = (p.Cout + BN - 1) / BN;
  const long long num_m = (p.M + 2 * BM - 1) / (2 * BM);
  const long long num_tiles = num_m * num_n;
  const int chunks = p.Cin / KC;
  const int kblocks = p.taps * chunks;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int HW = p.H * p.W;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&full[i], 2); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 2 * kEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, C::tmem_cols);
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = cluster_id; t < num_tiles; t += num_clusters) {
        const long long mb = t / num_n;
        const int nb = (int)(t - mb * num_n);
        const long long p0 = (mb * 2 + rank) * BM;             // this CTA's 128 pixels
        const int n0 = (int)(p0 / HW), rem = (int)(p0 - (long long)n0 * HW);
        const int y0 = rem / p.W, x0 = rem - y0 * p.W;
        const int col0 = nb * BN + (int)rank * (BN / 2);       // this CTA's half of the weight tile
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.dy0 + tap / p.tw, dx = p.dx0 + tap % p.tw;
          for (int cc = 0; cc < chunks; ++cc) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * C::stage_bytes;
            mbar_expect_tx_leader(&full[stage], A_BYTES + C::b_bytes);
            tma_load_4d_pair(sa, &tmA, &full[stage], cc * KC, x0 + dx, y0 + dy, n0);
            tma_load_3d_pair(sa + A_BYTES, &tmB, &full[stage], cc * KC, col0, p.wtap0 + tap);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (long long t = cluster_id; t < num_tiles; t += num_clusters) {
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + (size_t)stage * C::stage_bytes);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < KC / UMMA_K; ++k)
            umma2_f16(d, umma_desc_k_sw64(a_addr + k * UMMA_K * 2), umma_desc_k_sw64(b_addr + k * UMMA_K * 2), idesc,
                      (uint32_t)((kb | k) != 0));
          umma2_commit_mc(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(&tfull[as]);
        if (++as == kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    epilogue_loop<BN>(p, tmem_base, tfull, tempty, warp, lane, rank, cluster_id, num_clusters, num_tiles, num_n, HW);
  }
  ptx::tc_fence_before();
  cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) tmem_dealloc2(tmem_base, C::tmem_cols);
}


// ------------------------------------------------------------------------------------------------------------------------
// Halo-strip kernel for 3x3 convolutions on images at least 128 pixels wide (76 % of the decoder's convolution FLOPs).
// The generic kernel above re-loads the activation for each of the nine taps: 13 KiB of operands per 160 tensor-core cycles
// per SM, 9.6 TB/s of L2->SM traffic chip-wide -- it runs at the L2 bandwidth limit with the tensor pipe half idle (ncu:
// l1tex__m_xbar2l1tex_read_bytes 19.7 GB for a 160->160 layer at 256x256, lts throughput 61 %).  Here a CTA's 128 output pixels lie
// on ONE image row, and per channel chunk ONE 4-D TMA box brings the three input rows y-1..y+1 with a one-pixel halo on either
// side (3 x 130 pixels) into shared memory; the nine taps are nine VIEWS of that strip: the UMMA descriptor of tap (dy, dx)
// simply starts (dy+1)*130 + (dx+1) rows into it.  Activation traffic drops 3x; only the weight tiles stream per tap.
template <int BN, int KCH>
struct HaloCfg {
  static constexpr int rowb = KCH * 2;                              // 64 -> SWIZZLE_64B, 128 -> SWIZZLE_128B
  static constexpr int strip_px = BM + 2;
  static constexpr int a_bytes = 3 * strip_px * rowb;
  static constexpr int a_stride = (a_bytes + 1023) & ~1023;
  static constexpr int b_bytes = 3 * (BN / 2) * rowb;               // one kernel row: the three dx taps of one dy
  static constexpr int b_tap = (BN / 2) * rowb;
  static constexpr int b_stride = (b_bytes + 1023) & ~1023;
  static constexpr int a_stages = KCH == 32 ? 4 : 2;
  static constexpr int b_stages = (216 * 1024 - a_stages * a_stride) / b_stride > 16 ? 16 : (216 * 1024 - a_stages * a_stride) / b_stride;
  static constexpr int tmem_cols = 2 * BN <= 32 ? 32 : 2 * BN <= 256 ? 256 : 512;
  static constexpr size_t smem_bytes = 1024 + (size_t)a_stages * a_stride + (size_t)b_stages * b_stride + kBarBytes;
};

template <int BN, int KCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
  using C = HaloCfg<BN, KCH>;
  constexpr int kAS = C::a_stages, kBS = C::b_stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)kAS * C::a_stride;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + (size_t)kBS * C::b_stride);
  uint64_t* a_empty = a_full + kAS;
  uint64_t* b_full = a_empty + kAS;
  uint64_t* b_empty = b_full + kBS;
  uint64_t* tfull = b_empty + kBS;
  uint64_t* tempty = tfull + kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + kAccStages);
  static_assert((2 * 4 + 2 * 16 + 2 * kAccStages) * 8 + 4 <= kBarBytes, "barrier region");

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_n = (p.Cout + BN - 1) / BN;
  const long long num_m = (p.M + 2 * BM - 1) / (2 * BM);
  const long long num_tiles = num_m * num_n;
  const int chunks = (p.Cin + KCH - 1) / KCH;     // a ragged last chunk (Cin = 160 with 64-channel chunks) is zero-filled by the TMA unit
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int HW = p.H * p.W;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kAS; ++i) { ptx::mbar_init(&a_full[i], 2); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kBS; ++i) { ptx::mbar_init(&b_full[i], 2); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { ptx::mbar_init(&tfull[i], 1); ptx::mbar_init(&tempty[i], 2 * kEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, C::tmem_cols);
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (long long t = cluster_id; t < num_tiles; t += num_clusters) {
        const long long mb = t / num_n;
        const int nb = (int)(t - mb * num_n);
        const long long p0 = (mb * 2 + rank) * BM;             // this CTA's 128 pixels: one image row segment
        const int n0 = (int)(p0 / HW), rem = (int)(p0 - (long long)n0 * HW);
        const int y0 = rem / p.W, x0 = rem - y0 * p.W;
        const int col0 = nb * BN + (int)rank * (BN / 2);
        for (int cc = 0; cc < chunks; ++cc) {
          ptx::mbar_wait(&a_empty[as], aph ^ 1);
          mbar_expect_tx_leader(&a_full[as], C::a_bytes);
          tma_load_4d_pair(sA + (size_t)as * C::a_stride, &tmA, &a_full[as], cc * KCH, x0 - 1, y0 - 1, n0);
          if (++as == kAS) { as = 0; aph ^= 1; }
          for (int ky = 0; ky < p.tw; ++ky) {                  // weights of one kernel row: (KCH, BN/2, tw taps) in one box
            ptx::mbar_wait(&b_empty[bs], bph ^ 1);
            mbar_expect_tx_leader(&b_full[bs], (uint32_t)(p.tw * C::b_tap));
            tma_load_3d_pair(sB + (size_t)bs * C::b_stride, &tmB, &b_full[bs], cc * KCH, col0, p.wtap0 + p.tw * ky);
            if (++bs == kBS) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, BN);
      int as = 0, bs = 0, acs = 0;
      uint32_t aph = 0, bph = 0, acph = 0;
      for (long long t = cluster_id; t < num_tiles; t += num_clusters) {
        ptx::mbar_wait(&tempty[acs], acph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acs * BN);
        for (int cc = 0; cc < chunks; ++cc) {
          ptx::mbar_wait(&a_full[as], aph);
          const uint32_t strip = ptx::smem_u32(sA + (size_t)as * C::a_stride);
          for (int ky = 0; ky < p.tw; ++ky) {
            ptx::mbar_wait(&b_full[bs], bph);
            ptx::tc_fence_after();
            // strip row 0 is image row y - 1, strip pixel 0 is x - 1: tap (ky, kx) reads row y + dy0 + ky, pixel x + dx0 + kx
            const uint32_t a_row = strip + (uint32_t)(((p.dy0 + 1 + ky) * C::strip_px + p.dx0 + 1) * C::rowb);
            const uint32_t b_row = ptx::smem_u32(sB + (size_t)bs * C::b_stride);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              if (kx >= p.tw) break;
#pragma unroll
              for (int k = 0; k < KCH / UMMA_K; ++k)
                umma2_f16(d, umma_desc_k_rows<C::rowb>(a_row + kx * C::rowb + k * UMMA_K * 2),
                          umma_desc_k_rows<C::rowb>(b_row + kx * C::b_tap + k * UMMA_K * 2), idesc, (uint32_t)((cc | ky | kx | k) != 0));
            }
            umma2_commit_mc(&b_empty[bs]);
            if (++bs == kBS) { bs = 0; bph ^= 1; }
          }
          umma2_commit_mc(&a_empty[as]);
          if (++as == kAS) { as = 0; aph ^= 1; }
        }
        umma2_commit_mc(&tfull[acs]);
        if (++acs == kAccStages) { acs = 0; acph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    epilogue_loop<BN>(p, tmem_base, tfull, tempty, warp, lane, rank, cluster_id, num_clusters, num_tiles, num_n, HW);
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, C::tmem_cols);
}

template <int BN, int KCH>
static int launch_halo(const void* x, const void* w_packed, const Params& p, cudaStream_t st) {
  using C = HaloCfg<BN, KCH>;
  CUtensorMap tmA, tmB;
  const uint64_t dimsA[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.Nimg};
  const uint64_t strA[3] = {(uint64_t)p.Cin * 2, (uint64_t)p.W * p.Cin * 2, (uint64_t)p.H * p.W * p.Cin * 2};
  const uint32_t boxA[4] = {(uint32_t)KCH, (uint32_t)C::strip_px, 3u, 1u};
  const uint64_t dimsB[3] = {(uint64_t)p.Cin, (uint64_t)p.Cout, (uint64_t)p.wtaps};
  const uint64_t strB[2] = {(uint64_t)p.Cin * 2, (uint64_t)p.Cout * p.Cin * 2};
  const uint32_t boxB[3] = {(uint32_t)KCH, (uint32_t)(BN / 2), (uint32_t)p.tw};
  if (KCH == 32) {
    if (int rc = make_tmap_bf16_sw64(&tmA, x, 4, dimsA, strA, boxA)) return rc;
    if (int rc = make_tmap_bf16_sw64(&tmB, w_packed, 3, dimsB, strB, boxB)) return rc;
  } else {
    if (int rc = make_tmap_bf16_nd(&tmA, x, 4, dimsA, strA, boxA)) return rc;
    if (int rc = make_tmap_bf16_nd(&tmB, w_packed, 3, dimsB, strB, boxB)) return rc;
  }
  auto kern = conv_halo_kernel<BN, KCH>;
  SDVAR_SET_SMEM_ONCE(kern, C::smem_bytes);
  const int sms = sm_count();
  const long long tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.Cout + BN - 1) / BN);
  const int clusters = tiles < sms / 2 ? (int)tiles : sms / 2;
  kern<<<2 * clusters, kThreads, C::smem_bytes, st>>>(tmA, tmB, p);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

template <int BN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p, cudaStream_t st) {
  SDVAR_SET_SMEM_ONCE(conv_kernel<BN>, Cfg<BN>::smem_bytes);
  const int sms = sm_count();
  const long long tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * ((p.Cout + BN - 1) / BN);
  const int clusters = tiles < sms / 2 ? (int)tiles : sms / 2;
  conv_kernel<BN><<<2 * clusters, kThreads, Cfg<BN>::smem_bytes, st>>>(tmA, tmB, p);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

}  // namespace conv
}  // namespace sdvar

using namespace sdvar;

// one launch: `taps` = tw*tw taps at offsets (dy0 + t / tw, dx0 + t % tw) using weight matrices wtap0 + t of the `wtaps` packed ones;
// up = 1 writes input-grid pixel (n, Y, X) to pixel (n, 2Y + pa, 2X + pb) of a (2H, 2W) output
static int conv_dispatch(const sdvar_bf16* x, int N, int H, int W, int Cin, const sdvar_bf16* w_packed, int wtaps, int tw, int dy0, int dx0,
                         int wtap0, int up, int pa, int pb, int Cout, const float* bias, const sdvar_bf16* res, sdvar_bf16* y,
                         float* y_f32_nchw, float lo, float hi, cudaStream_t st) {
  // pixel box of one CTA: 128 consecutive pixels in (n, y, x) order must be a box BW x BH x BI
  const int BW = W < conv::BM ? W : conv::BM;
  SDVAR_REQUIRE(W % BW == 0 && conv::BM % BW == 0, "W=%d must divide or be a multiple of %d", W, conv::BM);
  const int BH = H < conv::BM / BW ? H : conv::BM / BW;
  SDVAR_REQUIRE(H % BH == 0 && (conv::BM / BW) % BH == 0, "H=%d does not tile into %d-pixel boxes of width %d", H, conv::BM, BW);
  const int BI = conv::BM / (BW * BH);
  const int taps = tw * tw;
  conv::Params p;
  p.Nimg = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = taps; p.BW = BW; p.BH = BH;
  p.tw = tw; p.dy0 = dy0; p.dx0 = dx0; p.wtap0 = wtap0; p.wtaps = wtaps; p.up = up; p.pa = pa; p.pb = pb;
  p.M = (long long)N * H * W;
  p.bias = bias;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res);
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(y);
  p.out_f32_nchw = y_f32_nchw;
  p.lo = lo; p.hi = hi;
  const int BN = Cout % 160 == 0 ? 160 : Cout >= 128 ? 128 : Cout > 16 ? 32 : 16;
  ProfileScope prof(st, FAM_CONV, 2.0 * (double)p.M * Cout * Cin * taps);
  static const bool no_halo = getenv("SDVAR_CONV_NO_HALO") != nullptr;   // A/B switch for profiling
  if (tw >= 2 && W % conv::BM == 0 && (BN == 160 || BN == 128 || BN == 16) && !no_halo) {
    // rows of at least 128 pixels: halo-strip kernel (BN = 16: conv_out, bound by the activation stream alone)
    if (BN == 160) return conv::launch_halo<160, 32>(x, w_packed, p, st);
    if (BN == 16) return conv::launch_halo<16, 32>(x, w_packed, p, st);
    return conv::launch_halo<128, 32>(x, w_packed, p, st);
  }
  CUtensorMap tmA, tmB;
  const uint64_t dimsA[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  const uint64_t strA[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
  const uint32_t boxA[4] = {(uint32_t)conv::KC, (uint32_t)BW, (uint32_t)BH, (uint32_t)BI};
  if (int rc = make_tmap_bf16_sw64(&tmA, x, 4, dimsA, strA, boxA)) return rc;
  const uint64_t dimsB[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)wtaps};
  const uint64_t strB[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
  const uint32_t boxB[3] = {(uint32_t)conv::KC, (uint32_t)(BN / 2), 1u};
  if (int rc = make_tmap_bf16_sw64(&tmB, w_packed, 3, dimsB, strB, boxB)) return rc;
  switch (BN) {
    case 160: return conv::launch<160>(tmA, tmB, p, st);
    case 128: return conv::launch<128>(tmA, tmB, p, st);
    case 32: return conv::launch<32>(tmA, tmB, p, st);
    default: return conv::launch<16>(tmA, tmB, p, st);
  }
}

extern "C" int sdvar_conv_nhwc(const sdvar_bf16* x, int N, int H, int W, int Cin, const sdvar_bf16* w_packed, int taps, int Cout,
                               const float* bias, const sdvar_bf16* res, sdvar_bf16* y, float* y_f32_nchw, float lo, float hi,
                               void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && w_packed && ((y != nullptr) != (y_f32_nchw != nullptr)), "NULL argument, or not exactly one output");
  SDVAR_REQUIRE(taps == 9 || taps == 1, "taps=%d (3x3 with padding 1, or 1x1)", taps);
  SDVAR_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cin % conv::KC == 0 && Cout > 0, "bad geometry N=%d H=%d W=%d Cin=%d Cout=%d", N, H, W,
                Cin, Cout);
  SDVAR_REQUIRE(y == nullptr || Cout % 8 == 0, "channels-last bf16 output needs Cout %% 8 == 0 (Cout=%d)", Cout);
  SDVAR_REQUIRE(y_f32_nchw == nullptr || res == nullptr, "the NCHW fp32 output takes no skip connection");
  SDVAR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)res & 15) == 0,
                "16-byte alignment");
  const int tw = taps == 9 ? 3 : 1;
  return conv_dispatch(x, N, H, W, Cin, w_packed, taps, tw, taps == 9 ? -1 : 0, taps == 9 ? -1 : 0, 0, 0, 0, 0, Cout, bias, res, y, y_f32_nchw,
                       lo, hi, (cudaStream_t)stream);
}

// conv3x3(padding 1) of the nearest-2x upsampled x WITHOUT materialising the upsampled tensor (reference Upsample2x,
// models/basic_vae.py:31-33).  Output pixel (2Y+a, 2X+b) sees only the 2x2 input pixels (Y+a-1.., X+b-1..): the nine taps
// collapse onto four whose weights are sums of the original ones, so each of the four parity classes (a, b) is a 2x2 convolution
// on the LOW-resolution input -- 16 multiply-adds per output channel pair instead of 36, and no 4x larger intermediate.
// w_par: (16, Cout, Cin) bf16, matrix (a*2 + b)*4 + (u*2 + v) = sum of the 3x3 weights that land on input pixel (Y+a-1+u, X+b-1+v).
extern "C" int sdvar_conv_up2x_nhwc(const sdvar_bf16* x, int N, int H, int W, int Cin, const sdvar_bf16* w_par, int Cout,
                                    const float* bias, sdvar_bf16* y, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && w_par && y, "NULL argument");
  SDVAR_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cin % conv::KC == 0 && Cout > 0 && Cout % 8 == 0, "bad geometry N=%d H=%d W=%d Cin=%d Cout=%d",
                N, H, W, Cin, Cout);
  SDVAR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_par & 15) == 0 && ((uintptr_t)y & 15) == 0, "16-byte alignment");
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      if (int rc = conv_dispatch(x, N, H, W, Cin, w_par, 16, 2, a - 1, b - 1, (a * 2 + b) * 4, 1, a, b, Cout, bias, nullptr, y, nullptr, 0.f,
                                 0.f, (cudaStream_t)stream))
        return rc;
  return SDVAR_OK;
}
