// sampling.cu -- K3 fused logits epilogue and K4 speculative verify (HBM-bound row kernels).
//
// Both kernels give one 256-thread CTA a row of V fp32 logits (V = 1024*NV): thread t owns the float4
// chunks f = i*256 + t, i.e. 128-bit coalesced streaming loads, 4*NV values in registers.  Every float
// operation that decides an index uses explicit round-to-nearest intrinsics (no FMA contraction) and the
// reduction order of oracle/spec_c/sdvar_spec.c ("256-lane order"), so the result is BIT-EXACT to that
// spec: top-k is a 32-step radix select on order-preserving keys (integer counts), top-p a 32-step
// bisection on the canonical masked mass, the sample an argmax with lowest-index tie break.
//
// Reference semantics: models/var.py:199-202, models/helpers.py:6-19; verify: SURVEY.md A7.
#include <stdlib.h>

#include "common.cuh"

namespace sdvar {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// exp(x) for x <= 0, bit-identical to sdvar_spec_expf.  Branch-free so the 16-32 independent exponentials of a thread
// interleave: no clamp (anything below -87.33, -inf and NaN included, is replaced by 0 with one select at the end) and the
// power of two is applied by ONE integer add to the exponent field -- every surviving result is a normal float, so the add is
// the exact product (round 1 clamped at -104 and multiplied twice to round denormal results like the C spec did).
constexpr float kExpMin = -87.33f;
__device__ __forceinline__ float spec_expf(float x) {
  // n = rint(x * log2 e) with one rounding and without the quarter-rate FRND / F2I conversions: the exact product is added to
  // 1.5*2^23 inside one fma (the spec is stated that way), the sum lands in the binade with ulp 1, so the fma itself rounds to
  // nearest-even; bits(tm) = 0x4B400000 + n
  const float tm = __fmaf_rn(x, 1.44269504088896340736f, 12582912.0f);
  const float n = __fsub_rn(tm, 12582912.0f);
  float r = __fmaf_rn(n, -0.693145751953125f, x);
  r = __fmaf_rn(n, -1.42860682030941723212e-6f, r);
  float p = 1.0f / 5040.0f;
  p = __fmaf_rn(p, r, 1.0f / 720.0f);
  p = __fmaf_rn(p, r, 1.0f / 120.0f);
  p = __fmaf_rn(p, r, 1.0f / 24.0f);
  p = __fmaf_rn(p, r, 1.0f / 6.0f);
  p = __fmaf_rn(p, r, 0.5f);
  p = __fmaf_rn(p, r, 1.0f);
  p = __fmaf_rn(p, r, 1.0f);
  const float e = u2f(f2u(p) + (f2u(tm) << 23));      // (0x4B400000 << 23) == 0 mod 2^32: the shift leaves n << 23
  return (x >= kExpMin) ? e : 0.0f;
}

// ---- two exponentials at once on the packed fp32x2 pipe (FFMA2 / FADD2 / FMUL2: one issue slot for two IEEE-RN operations, so
// the results are bit-identical to two spec_expf calls while the instruction count per exponential halves).
struct f32x2 {
  unsigned long long v;
};
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(f32x2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f32x2 splat2(float c) { return pk2(c, c); }
// (exp(x0), exp(x1)), both x <= 0 (or -inf); every lane operation is the one spec_expf performs, in the same order
__device__ __forceinline__ f32x2 spec_expf2(float x0, float x1) {
  const f32x2 xc = pk2(x0, x1);
  const f32x2 tm = fma2(xc, splat2(1.44269504088896340736f), splat2(12582912.0f));
  const f32x2 n = add2(tm, splat2(-12582912.0f));
  f32x2 r = fma2(n, splat2(-0.693145751953125f), xc);
  r = fma2(n, splat2(-1.42860682030941723212e-6f), r);
  f32x2 p = splat2(1.0f / 5040.0f);
  p = fma2(p, r, splat2(1.0f / 720.0f));
  p = fma2(p, r, splat2(1.0f / 120.0f));
  p = fma2(p, r, splat2(1.0f / 24.0f));
  p = fma2(p, r, splat2(1.0f / 6.0f));
  p = fma2(p, r, splat2(0.5f));
  p = fma2(p, r, splat2(1.0f));
  p = fma2(p, r, splat2(1.0f));
  float tm0, tm1, p0, p1;
  unpk2(tm, tm0, tm1);
  unpk2(p, p0, p1);
  const float e0 = u2f(f2u(p0) + (f2u(tm0) << 23)), e1 = u2f(f2u(p1) + (f2u(tm1) << 23));
  return pk2((x0 >= kExpMin) ? e0 : 0.0f, (x1 >= kExpMin) ? e1 : 0.0f);
}

// IEEE a / b for a >= 0 and b > 0.  ptxas' inline division has a fast path for operands with ordinary exponents and CALLs a
// ~60-instruction subroutine otherwise -- and a ZERO numerator counts as "otherwise".  Most numerators here are exact zeros
// (masked-out vocabulary entries), which sent every warp through the subroutine on every division (40 % of K3's
// instructions).  Dividing 1 instead and selecting 0 afterwards gives the same bits (0 / b == +0).
__device__ __forceinline__ float fdiv_nz(float a, float b) {
  const bool z = a == 0.0f;
  const float q = __fdiv_rn(z ? 1.0f : a, b);
  return z ? 0.0f : q;
}

// ---- block reductions.  `slot` alternates between two smem buffers so one barrier per reduction suffices.
struct RedSmem {
  float f[2][2][kWarps];
  int i[2][kWarps];
  unsigned long long u64[2][kWarps];
  uint32_t u[2][2][kWarps];
  int flag[2];
  float pq[2][2];
};

// canonical float sum of two independent quantities at once
__device__ __forceinline__ void block_sum2(float& a, float& b, RedSmem& s, int& slot) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, off));
    b = __fadd_rn(b, __shfl_xor_sync(0xffffffffu, b, off));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s.f[slot][0][w] = a; s.f[slot][1][w] = b; }
  __syncthreads();
  a = s.f[slot][0][0];
  b = s.f[slot][1][0];
#pragma unroll
  for (int k = 1; k < kWarps; ++k) { a = __fadd_rn(a, s.f[slot][0][k]); b = __fadd_rn(b, s.f[slot][1][k]); }
  slot ^= 1;
}
__device__ __forceinline__ float block_sum(float a, RedSmem& s, int& slot) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, off));
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s.f[slot][0][w] = a;
  __syncthreads();
  a = s.f[slot][0][0];
#pragma unroll
  for (int k = 1; k < kWarps; ++k) a = __fadd_rn(a, s.f[slot][0][k]);
  slot ^= 1;
  return a;
}
// canonical float sum and an integer count with a single barrier
__device__ __forceinline__ void block_sum_fi(float& a, int& c, RedSmem& s, int& slot) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, off));
  c = __reduce_add_sync(0xffffffffu, c);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s.f[slot][0][w] = a; s.i[slot][w] = c; }
  __syncthreads();
  a = s.f[slot][0][0];
  c = s.i[slot][0];
#pragma unroll
  for (int k = 1; k < kWarps; ++k) { a = __fadd_rn(a, s.f[slot][0][k]); c += s.i[slot][k]; }
  slot ^= 1;
}
__device__ __forceinline__ int block_sum_int(int c, RedSmem& s, int& slot) {
  c = __reduce_add_sync(0xffffffffu, c);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s.i[slot][w] = c;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int k = 0; k < kWarps; ++k) t += s.i[slot][k];
  slot ^= 1;
  return t;
}
__device__ __forceinline__ void block_max_u32x2(uint32_t& a, uint32_t& b, RedSmem& s, int& slot) {
  a = __reduce_max_sync(0xffffffffu, a);
  b = __reduce_max_sync(0xffffffffu, b);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s.u[slot][0][w] = a; s.u[slot][1][w] = b; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kWarps; ++k) { a = max(a, s.u[slot][0][k]); b = max(b, s.u[slot][1][k]); }
  slot ^= 1;
}
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, RedSmem& s, int& slot) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
    v = o > v ? o : v;
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s.u64[slot][w] = v;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kWarps; ++k) v = s.u64[slot][k] > v ? s.u64[slot][k] : v;
  slot ^= 1;
  return v;
}
// (value, index) -> u64 whose max is "largest value, then lowest index"
__device__ __forceinline__ unsigned long long pack_best(float r, int idx) {
  return ((unsigned long long)fkey(r) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}

// test hook: y[i] = (packed, scalar) spec exponentials of x[i], so the exp spec can be pinned element by element
__global__ void spec_expf_kernel(const float* __restrict__ x, long long n, float* __restrict__ y_packed, float* __restrict__ y_scalar) {
  for (long long i = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 2LL * gridDim.x * blockDim.x) {
    const float a = x[i], b = (i + 1 < n) ? x[i + 1] : 0.0f;
    float e0, e1;
    unpk2(spec_expf2(a, b), e0, e1);
    y_packed[i] = e0;
    y_scalar[i] = spec_expf(a);
    if (i + 1 < n) { y_packed[i + 1] = e1; y_scalar[i + 1] = spec_expf(b); }
  }
}

// ================================================================================================
// K3
// ================================================================================================
// Two kernels, selected by the launch-uniform filter setting (the spec has the same two regimes):
//   k3_plain_kernel     no top-k / top-p: CFG mix, softmax in the canonical 256-lane float order, exponential race.
//   k3_filtered_kernel  top-k and/or top-p.  Round 1 searched the k-th largest key and the top-p threshold by bisection,
//                       ~26 block-wide barrier rounds of 16 compares per thread each, and ran every exponential and division
//                       on all V entries although only ~k survive (0.15-0.23 of HBM).  Now:
//     1. the exact k-th largest value comes from ONE shared-memory histogram pass (512 value-space bins, integer atomics),
//        a suffix scan over the bins, and an exact rank among the handful of candidates in the crossing bin;
//     2. the survivors (~top_k of V) are compacted into shared memory together with their noise and vocabulary index, so
//        exponentials, probabilities, divisions and the race touch ~k/256 entries per thread instead of V/256;
//     3. every sum of the filtered regime is an integer sum of fixed-point terms (spec: E = rint(e 2^40), mass =
//        rint(p 2^30)), so it does not depend on the order in which the atomics / the compaction happen to run: the kernel
//        is free to reorder and still bit-exact to oracle/spec_c; the top-p cut is a second histogram (of masses) + scan +
//        exact resolution inside the crossing bin.
//     Degenerate rows (more than 256 candidates in a crossing bin, non-finite value range) take bisection fall-backs.
constexpr int kBins = 512;
constexpr float kFixE = 1099511627776.0f;   // 2^40
constexpr float kFixM = 1073741824.0f;      // 2^30

template <int NV>
__global__ void __launch_bounds__(kThreads, 4)
k3_plain_kernel(const float* __restrict__ logits, int B, int L, int in_ld, int in_off, int out_ld, int out_off, SegTable seg,
                const float* __restrict__ noise, long long* __restrict__ idx_out, float* __restrict__ mixed_out,
                float* __restrict__ prob_out) {
  constexpr int V = NV * 1024;
  constexpr int E = NV * 4;
  __shared__ RedSmem sm;
  int slot = 0;
  const int tid = threadIdx.x;
  const long long rows = (long long)B * L;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / L), pos = (int)(row - (long long)b * L);
    const int j = seg_of(seg, pos);
    const float t1 = seg.t1[j], t2 = seg.t2[j];
    const long long orow = (long long)b * out_ld + out_off + pos;
    const float4* pc = reinterpret_cast<const float4*>(logits + ((long long)b * in_ld + in_off + pos) * V);
    const float4* pu = reinterpret_cast<const float4*>(logits + ((long long)(B + b) * in_ld + in_off + pos) * V);
    float x[E];
    {
      float4 a[NV], c[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) { a[i] = ldg_stream(pc + i * kThreads + tid); c[i] = ldg_stream(pu + i * kThreads + tid); }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        x[4 * i + 0] = __fsub_rn(__fmul_rn(a[i].x, t1), __fmul_rn(c[i].x, t2));
        x[4 * i + 1] = __fsub_rn(__fmul_rn(a[i].y, t1), __fmul_rn(c[i].y, t2));
        x[4 * i + 2] = __fsub_rn(__fmul_rn(a[i].z, t1), __fmul_rn(c[i].z, t2));
        x[4 * i + 3] = __fsub_rn(__fmul_rn(a[i].w, t1), __fmul_rn(c[i].w, t2));
      }
    }
    if (mixed_out != nullptr) {
      float4* po = reinterpret_cast<float4*>(mixed_out + orow * V);
#pragma unroll
      for (int i = 0; i < NV; ++i) stg_stream(po + i * kThreads + tid, make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]));
    }
    if (noise == nullptr) continue;
    const float4* pn = reinterpret_cast<const float4*>(noise + row * V);
    float4 nz[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) nz[i] = ldg_stream(pn + i * kThreads + tid);
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < E; ++e) mx = fmaxf(mx, x[e]);
    uint32_t kmax = fkey(mx), z0 = 0;
    block_max_u32x2(kmax, z0, sm, slot);
    const float m = fkey_inv(kmax);
    float z = 0.0f;
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      unpk2(spec_expf2(__fsub_rn(x[e], m), __fsub_rn(x[e + 1], m)), x[e], x[e + 1]);   // x now holds the exponentials
      z = __fadd_rn(__fadd_rn(z, x[e]), x[e + 1]);
    }
    const float Z2 = block_sum(z, sm, slot);
    float best = -1.0f, bestp = 0.0f;
    int bi = 0x7FFFFFFF;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float nn[4] = {nz[i].x, nz[i].y, nz[i].z, nz[i].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float pv = __fdiv_rn(x[4 * i + c], Z2);
        const float r = __fdiv_rn(pv, nn[c]);
        if (r > best) { best = r; bi = 4 * (i * kThreads + tid) + c; bestp = pv; }
      }
    }
    const unsigned long long w = block_max_u64(pack_best(best, bi), sm, slot);
    int win = (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
    if (win == 0x7FFFFFFF) win = 0;
    if (bi == win || (tid == 0 && (uint32_t)(w >> 32) == fkey(-1.0f))) {
      if (idx_out) idx_out[orow] = win;
      if (prob_out) prob_out[orow] = (bi == win) ? bestp : 0.0f;
    }
  }
}

struct K3Smem {
  uint32_t hist[kBins];
  float cand[256];
  uint32_t candm[256];
  uint32_t wtot[kWarps];
  __align__(16) unsigned long long zpart[2][kWarps];   // per-warp fixed-point partial sums (two buffers: consecutive sums may have no barrier in between)
  int bstar;
  uint32_t above;
  int ncand, n_list;
  float xK;
  uint32_t tkey, minkey;
  RedSmem red;
};

// inclusive warp scan of a u32
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, v, off);
    if (lane >= off) v += o;
  }
  return v;
}
// block-wide inclusive scan over the 256 threads (thread order); `tot` receives the grand total.  One barrier.
__device__ __forceinline__ uint32_t block_incl_scan(uint32_t v, K3Smem& s, uint32_t& tot) {
  const int w = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v);
  if ((threadIdx.x & 31) == 31) s.wtot[w] = inc;
  __syncthreads();
  uint32_t base = 0, t = 0;
#pragma unroll
  for (int k = 0; k < kWarps; ++k) {
    const uint32_t x = s.wtot[k];
    base += (k < w) ? x : 0u;
    t += x;
  }
  tot = t;
  return inc + base;
}
// fixed-point exponential E = rint(e * 2^40) as two 20-bit limbs (hi * 2^20 + lo); e in [0, 1]
__device__ __forceinline__ void fix_e(float e, uint32_t& hi, uint32_t& lo) {
  const float a = __fmul_rn(e, 1048576.0f);                 // e * 2^20, exact
  const float fh = floorf(a);
  hi = (uint32_t)fh;
  lo = __float2uint_rn(__fmul_rn(__fsub_rn(a, fh), 1048576.0f));   // both operations exact; one rounding in the conversion
}
// block sum of per-thread limb sums -> u64 total: two warp reductions, one 64-bit partial per warp, one barrier, then every
// thread adds the 8 partials (4 LDS.128).  Integer sums: any order gives the same total.  (per-thread sums < 2^25, warp < 2^30)
__device__ __forceinline__ unsigned long long block_sum_fix(uint32_t hi, uint32_t lo, K3Smem& s, int buf) {
  hi = __reduce_add_sync(0xffffffffu, hi);
  lo = __reduce_add_sync(0xffffffffu, lo);
  if ((threadIdx.x & 31) == 0) s.zpart[buf][threadIdx.x >> 5] = ((unsigned long long)hi << 20) + (unsigned long long)lo;
  __syncthreads();
  const ulonglong2* z2 = reinterpret_cast<const ulonglong2*>(s.zpart[buf]);
  unsigned long long t = 0;
#pragma unroll
  for (int k = 0; k < kWarps / 2; ++k) { const ulonglong2 v = z2[k]; t += v.x + v.y; }
  return t;
}

// value-space bin of x: floor(x * scale + off) with off = 0.5 - lo * scale; for lo <= x <= hi the fused multiply-add stays
// inside (0, 512) as long as |lo| * scale < 2^18 (checked by the caller), so no clamp is needed.  Monotone in x.
__device__ __forceinline__ int vbin(float x, float scale, float off) { return __float2int_rd(__fmaf_rn(x, scale, off)); }

// ---- the part of the filtered path that works on the compacted list; QR = list entries per thread (n <= 256 * QR).
// The list holds every entry whose top-k bin is >= the crossing bin: all survivors plus the few candidates below the exact
// cut, which are dropped here (they contribute nothing once xK is known).
struct K3Tail {      // returned BY VALUE: reference parameters of the out-of-line instance would pin these in local memory
  float xlow;
  int incl, slot;
};
template <int QR>
__device__ __forceinline__ K3Tail k3_tail(K3Smem& s, int slot, const float2* lxn, const uint16_t* li, int n, float m, float xmin,
                                          bool use_k, int top_k, bool hist_k, float kscale, float koff, float thr, bool sample,
                                          long long orow, long long* idx_out, float* prob_out) {
  const int tid = threadIdx.x;
  float xlow;
  bool incl;
  float xv[QR], nzv[QR];
#pragma unroll
  for (int q = 0; q < QR; ++q) {
    const int i = q * kThreads + tid;
    const float2 t = i < n ? lxn[i] : make_float2(-INFINITY, 1.0f);
    xv[q] = t.x; nzv[q] = t.y;
  }
  // ---- exact k-th largest value among the candidates of the crossing bin ----
  float xK = xmin;
  if (use_k && hist_k) {
    const int bst = s.bstar;
    const uint32_t above = s.above;
#pragma unroll
    for (int q = 0; q < QR; ++q)
      if (q * kThreads + tid < n && vbin(xv[q], kscale, koff) == bst) {
        const int p = atomicAdd(&s.ncand, 1);
        if (p < 256) s.cand[p] = xv[q];
      }
    __syncthreads();
    const int nc = s.ncand;
    if (nc <= 256) {
      if (tid < nc) {
        const float xi = s.cand[tid];
        uint32_t g = 0, ge = 0;
        for (int jj = 0; jj < nc; ++jj) { const float xj = s.cand[jj]; g += xj > xi; ge += xj >= xi; }
        const uint32_t kk = (uint32_t)top_k - above;
        if (g < kk && kk <= ge) s.xK = xi;
      }
      __syncthreads();
      xK = s.xK;
    } else {
      // fall-back: bisection over the candidates' keys (all other list entries are above the bin and count as `above`)
      uint32_t lo = 0u, hi = 0xFFFFFFFFu;      // invariant: #{cand key >= lo} >= kk, #{cand key >= hi} < kk  (hi exclusive top)
      const int kk = top_k - (int)above;
      uint32_t key[QR];
#pragma unroll
      for (int q = 0; q < QR; ++q)
        key[q] = (q * kThreads + tid < n && vbin(xv[q], kscale, koff) == bst) ? fkey(__fadd_rn(xv[q], 0.0f)) : 0u;   // 0: below every real key
#pragma unroll 1
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t tr = lo | (1u << bit);
        int c = 0;
#pragma unroll
        for (int q = 0; q < QR; ++q) c += (key[q] >= tr) ? 1 : 0;
        c = block_sum_int(c, s.red, slot);
        if (c >= kk) lo = tr;
      }
      (void)hi;
      xK = fkey_inv(lo);
    }
    if (tid == 0) s.ncand = 0;      // next use is behind at least one barrier
  } else if (use_k) {
    xK = s.xK;                        // found by the caller's fall-back
  }
  // ---- exponentials of the survivors; dropped candidates get exp(-inf) = 0 and vanish from every sum ----
  float ev[QR];
  uint32_t hi = 0, lo = 0;
#pragma unroll
  for (int q = 0; q < QR; q += 2) {
    const float x0 = (xv[q] >= xK) ? __fsub_rn(xv[q], m) : -INFINITY;
    const float x1 = (q + 1 < QR && xv[q + 1] >= xK) ? __fsub_rn(xv[q + 1], m) : -INFINITY;
    float e0, e1;
    unpk2(spec_expf2(x0, x1), e0, e1);
    ev[q] = e0;
    if (q + 1 < QR) ev[q + 1] = e1;
    uint32_t h, l;
    fix_e(e0, h, l); hi += h; lo += l;
    fix_e(e1, h, l); hi += h; lo += l;
  }
  const unsigned long long Zi = block_sum_fix(hi, lo, s, 0);
  unsigned long long Z2i = Zi;
  xlow = xK; incl = true;            // keep <=> x >= xK
  if (thr >= 0.0f) {
    const float Z = __fmul_rn(__ull2float_rn(Zi), 1.0f / kFixE);
    const uint32_t thr_i = (uint32_t)__fmul_rn(thr, kFixM);
    const float range = __fsub_rn(m, xK);
    uint32_t mass[QR];
#pragma unroll
    for (int q = 0; q < QR; ++q) mass[q] = ev[q] > 0.0f ? __float2uint_rn(__fmul_rn(__fdiv_rn(ev[q], Z), kFixM)) : 0u;
    bool all_but_max = false, have = false;
    float xB = -INFINITY;
    bool strictB = true;
    if (range > 0.0f && range < INFINITY && __fmul_rn(fabsf(xK), __fdiv_rn(511.0f, range)) < 262144.0f) {
      const float scale = __fdiv_rn(511.0f, range), off = __fmaf_rn(-xK, scale, 0.5f);
#pragma unroll
      for (int q = 0; q < QR; ++q)
        if (xv[q] >= xK) atomicAdd(&s.hist[vbin(xv[q], scale, off)], mass[q]);
      __syncthreads();
      // ascending scan of the bin masses: the crossing bin is the first whose inclusive cumulation exceeds thr_i
      const uint32_t m0 = s.hist[2 * tid], m1 = s.hist[2 * tid + 1];
      uint32_t tot;
      const uint32_t inc = block_incl_scan(m0 + m1, s, tot);
      const uint32_t before = inc - (m0 + m1);
      s.hist[2 * tid] = 0; s.hist[2 * tid + 1] = 0;
      if (before <= thr_i && thr_i < before + m0) { s.bstar = 2 * tid; s.above = before; }
      else if (before + m0 <= thr_i && thr_i < inc) { s.bstar = 2 * tid + 1; s.above = before + m0; }
      __syncthreads();
      if (tot <= thr_i) {
        all_but_max = true;                  // the whole row is removable: only the maximum survives
      } else {
        const int bst = s.bstar;
        const uint32_t below = s.above;
#pragma unroll
        for (int q = 0; q < QR; ++q)
          if (xv[q] >= xK && vbin(xv[q], scale, off) == bst) {
            const int p = atomicAdd(&s.ncand, 1);
            if (p < 256) { s.cand[p] = xv[q]; s.candm[p] = mass[q]; }
          }
        __syncthreads();
        const int nc = s.ncand;
        if (nc <= 256) {
          if (tid < nc) {
            const float xi = s.cand[tid];
            uint32_t le = below;
            for (int jj = 0; jj < nc; ++jj) le += (s.cand[jj] <= xi) ? s.candm[jj] : 0u;
            const uint32_t k = fkey(__fadd_rn(xi, 0.0f));
            if (le <= thr_i) atomicMax(&s.tkey, k);
            atomicMin(&s.minkey, k);
          }
          __syncthreads();
          const uint32_t tk = s.tkey;
          strictB = tk == 0u;
          xB = fkey_inv(strictB ? s.minkey : tk);
          have = true;
        }
      }
    } else if (range > 0.0f || !(range == 0.0f)) {
      have = false;                            // degenerate value range: bit-serial search below
    } else {
      have = true;                             // every survivor equals the maximum: one tie group holding the max, nothing removed
    }
    if (!all_but_max && !have) {
      // fall-back (crossing bin with > 256 entries, or a value range the histogram cannot bin): bit-serial search of the
      // largest key T with mass{key <= T} <= thr_i over the survivors
      uint32_t T = 0;
#pragma unroll 1
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t tr = T | (1u << bit);
        int a = 0;
#pragma unroll
        for (int q = 0; q < QR; ++q)
          if (xv[q] >= xK && fkey(__fadd_rn(xv[q], 0.0f)) <= tr) a += (int)mass[q];
        a = block_sum_int(a, s.red, slot);      // masses sum to ~2^30: fits an int
        if ((uint32_t)a <= thr_i) T = tr;
      }
      strictB = T == 0u;
      xB = T ? fkey_inv(T) : -INFINITY;
    }
    // one threshold for everything below: keep <=> incl ? x >= xlow : x > xlow  (the row maximum always satisfies it)
    if (all_but_max) { xlow = m; incl = true; }
    else if (strictB) { if (xB > xK) { xlow = xB; incl = true; } }        // removed <=> x < xB
    else if (xB >= xK) { xlow = xB; incl = false; }                         // removed <=> x <= xB
    hi = 0; lo = 0;
#pragma unroll
    for (int q = 0; q < QR; ++q) {
      const bool keep = incl ? (xv[q] >= xlow) : (xv[q] > xlow);
      if (!keep) ev[q] = 0.0f;
      uint32_t h, l;
      fix_e(ev[q], h, l); hi += h; lo += l;
    }
    if (sample) Z2i = block_sum_fix(hi, lo, s, 1);
  }
  if (!sample) return K3Tail{xlow, incl ? 1 : 0, slot};
  const float Z2 = __fmul_rn(__ull2float_rn(Z2i), 1.0f / kFixE);
  float best = -1.0f, bestp = 0.0f;
  int bi = 0x7FFFFFFF;
#pragma unroll
  for (int q = 0; q < QR; ++q) {
    if (ev[q] > 0.0f) {
      const float pv = __fdiv_rn(ev[q], Z2);
      const float r = __fdiv_rn(pv, nzv[q]);
      const int v = (int)li[q * kThreads + tid];
      if (r > best || (r == best && v < bi)) { best = r; bi = v; bestp = pv; }
    }
  }
  const unsigned long long w = block_max_u64(pack_best(best, bi), s.red, slot);
  int win = (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
  if (win == 0x7FFFFFFF) win = 0;
  if (bi == win || (tid == 0 && (uint32_t)(w >> 32) == fkey(-1.0f))) {
    if (idx_out) idx_out[orow] = win;
    if (prob_out) prob_out[orow] = (bi == win) ? bestp : 0.0f;
  }
  return K3Tail{xlow, incl ? 1 : 0, slot};
}

// long lists (top-p without top-k, huge tie groups): kept out of line so that its 4*E-register working set does not
// inflate the register allocation of the common path
template <int E>
__device__ __noinline__ K3Tail k3_tail_long(K3Smem& s, int slot, const float2* lxn, const uint16_t* li, int n, float m, float xmin,
                                            bool use_k, int top_k, bool hist_k, float kscale, float koff, float thr, bool sample,
                                            long long orow, long long* idx_out, float* prob_out) {
  return k3_tail<E>(s, slot, lxn, li, n, m, xmin, use_k, top_k, hist_k, kscale, koff, thr, sample, orow, idx_out, prob_out);
}

template <int NV, int OCC>
__global__ void __launch_bounds__(kThreads, OCC)
k3_filtered_kernel(const float* __restrict__ logits, int B, int L, int in_ld, int in_off, int out_ld, int out_off, SegTable seg,
                   int top_k, float thr, const float* __restrict__ noise, long long* __restrict__ idx_out,
                   float* __restrict__ mixed_out, float* __restrict__ prob_out) {
  constexpr int V = NV * 1024;
  constexpr int E = NV * 4;
  constexpr int QF = E < 6 ? E : 6;     // fast tail: lists of up to 1536 entries
  extern __shared__ __align__(16) unsigned char k3_dyn[];
  float2* lxn = reinterpret_cast<float2*>(k3_dyn);           // [V + 2] (logit, noise) of the listed entries (+ the dummy slot)
  uint16_t* li = reinterpret_cast<uint16_t*>(lxn + V + 2);   // [V + 2] their vocabulary index
  __shared__ K3Smem s;
  int slot = 0;
  const int tid = threadIdx.x, lane = tid & 31;
  const long long rows = (long long)B * L;
  for (int i = tid; i < kBins; i += kThreads) s.hist[i] = 0;
  if (tid == 0) { s.ncand = 0; s.n_list = 0; s.tkey = 0; s.minkey = 0xFFFFFFFFu; }
  __syncthreads();
  const bool sample = noise != nullptr;
  const bool use_k = top_k > 0 && top_k < V;
  const bool small = rows < 0x7FFFFFFFLL;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    int b, pos;
    if (small) { b = (int)((uint32_t)row / (uint32_t)L); pos = (int)((uint32_t)row - (uint32_t)b * (uint32_t)L); }
    else { b = (int)(row / L); pos = (int)(row - (long long)b * L); }
    const int j = seg_of(seg, pos);
    const float t1 = seg.t1[j], t2 = seg.t2[j];
    const long long orow = (long long)b * out_ld + out_off + pos;
    const float4* pc = reinterpret_cast<const float4*>(logits + ((long long)b * in_ld + in_off + pos) * V);
    const float4* pu = reinterpret_cast<const float4*>(logits + ((long long)(B + b) * in_ld + in_off + pos) * V);
    float x[E];
    {
      float4 a[NV], c[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) { a[i] = ldg_stream(pc + i * kThreads + tid); c[i] = ldg_stream(pu + i * kThreads + tid); }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        x[4 * i + 0] = __fsub_rn(__fmul_rn(a[i].x, t1), __fmul_rn(c[i].x, t2));
        x[4 * i + 1] = __fsub_rn(__fmul_rn(a[i].y, t1), __fmul_rn(c[i].y, t2));
        x[4 * i + 2] = __fsub_rn(__fmul_rn(a[i].z, t1), __fmul_rn(c[i].z, t2));
        x[4 * i + 3] = __fsub_rn(__fmul_rn(a[i].w, t1), __fmul_rn(c[i].w, t2));
      }
    }
    // ---- row max / min ----
    float mx = -INFINITY, mn = INFINITY;
#pragma unroll
    for (int e = 0; e < E; ++e) { mx = fmaxf(mx, x[e]); mn = fminf(mn, x[e]); }
    uint32_t kmx = fkey(__fadd_rn(mx, 0.0f)), kmnc = ~fkey(__fadd_rn(mn, 0.0f));
    block_max_u32x2(kmx, kmnc, s.red, slot);
    const float m = fkey_inv(kmx), xmin = fkey_inv(~kmnc);

    // ---- top-k, coarse: histogram of the value-space bins, suffix scan -> crossing bin b* (exact cut: in the tail) ----
    bool hist_k = false;
    float kscale = 0.0f, koff = 0.0f, xcut = xmin;      // list <=> vbin(x) >= b*  (hist) or x >= xcut (fall-back / no top-k)
    int bst = 0;
    if (use_k) {
      const float range = __fsub_rn(m, xmin);
      if (range > 0.0f && range < INFINITY && __fmul_rn(fabsf(xmin), __fdiv_rn(511.0f, range)) < 262144.0f) {
        hist_k = true;
        kscale = __fdiv_rn(511.0f, range);
        koff = __fmaf_rn(-xmin, kscale, 0.5f);
#pragma unroll
        for (int e = 0; e < E; ++e) atomicAdd(&s.hist[vbin(x[e], kscale, koff)], 1u);
      }
    }
    float4 nz[NV];
    if (sample) {      // issued here so that the latency overlaps the scan
      const float4* pn = reinterpret_cast<const float4*>(noise + row * V);
#pragma unroll
      for (int i = 0; i < NV; ++i) nz[i] = ldg_stream(pn + i * kThreads + tid);
    }
    if (use_k) {
      if (hist_k) {
        __syncthreads();
        // suffix scan: thread t owns bins 511-2t (first) and 510-2t
        const uint32_t c0 = s.hist[kBins - 1 - 2 * tid], c1 = s.hist[kBins - 2 - 2 * tid];
        uint32_t tot;
        const uint32_t inc = block_incl_scan(c0 + c1, s, tot);
        const uint32_t before = inc - (c0 + c1);
        s.hist[kBins - 1 - 2 * tid] = 0; s.hist[kBins - 2 - 2 * tid] = 0;
        const uint32_t k = (uint32_t)top_k;
        if (before < k && k <= before + c0) { s.bstar = kBins - 1 - 2 * tid; s.above = before; }
        else if (before + c0 < k && k <= inc) { s.bstar = kBins - 2 - 2 * tid; s.above = before + c0; }
        __syncthreads();
        bst = s.bstar;
      } else {
        // fall-back: bisection on the order-preserving keys (round-1 algorithm), any input
        uint32_t key[E];
#pragma unroll
        for (int e = 0; e < E; ++e) key[e] = fkey(__fadd_rn(x[e], 0.0f));
        uint32_t lo = ~kmnc, hi = kmx + 1u;
        int cL = V, cH = 0;
#pragma unroll 1
        while (cL - cH > 1 && hi - lo > 1u) {
          const uint32_t mid = lo + ((hi - lo) >> 1);
          int c = 0;
#pragma unroll
          for (int e = 0; e < E; ++e) c += (key[e] >= mid) ? 1 : 0;
          c = block_sum_int(c, s.red, slot);
          if (c >= top_k) { lo = mid; cL = c; } else { hi = mid; cH = c; }
        }
        uint32_t K = lo;
        if (cL - cH == 1 && hi - lo > 1u) {
          uint32_t mnk = 0, z = 0;
#pragma unroll
          for (int e = 0; e < E; ++e) mnk = max(mnk, (key[e] >= lo) ? ~key[e] : 0u);
          block_max_u32x2(mnk, z, s.red, slot);
          K = ~mnk;
        }
        xcut = fkey_inv(K);
        if (tid == 0) s.xK = xcut;
      }
    }

    // ---- compact the listed entries (any order: every later sum is an integer sum).  Branch-free: entries that are not
    // listed are stored to a dummy slot (index V) instead of being jumped over -- a profile of the branchy version showed
    // ~290 of ~2060 instructions per thread-row in BRA / BSSY / BSYNC and ~120 in spill traffic of the per-element flags.
    {
      uint32_t cnt = 0;
      if (hist_k) {
#pragma unroll
        for (int e = 0; e < E; ++e) cnt += (vbin(x[e], kscale, koff) >= bst) ? 1u : 0u;
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) cnt += (x[e] >= xcut) ? 1u : 0u;
      }
      const uint32_t inc = warp_incl_scan(cnt);
      int base = 0;
      if (lane == 31) base = atomicAdd(&s.n_list, (int)inc);
      base = __shfl_sync(0xffffffffu, base, 31);
      int p = base + (int)(inc - cnt);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float nn[4] = {nz[i].x, nz[i].y, nz[i].z, nz[i].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float xe = x[4 * i + c];
          const bool a = hist_k ? (vbin(xe, kscale, koff) >= bst) : (xe >= xcut);
          const int dst = a ? p : V;
          lxn[dst] = make_float2(xe, sample ? nn[c] : 1.0f);
          li[dst] = (uint16_t)(4 * (i * kThreads + tid) + c);
          p += a ? 1 : 0;
        }
      }
    }
    __syncthreads();
    const int n = s.n_list;
    const K3Tail tl = (n <= QF * kThreads)
        ? k3_tail<QF>(s, slot, lxn, li, n, m, xmin, use_k, top_k, hist_k, kscale, koff, thr, sample, orow, idx_out, prob_out)
        : k3_tail_long<E>(s, slot, lxn, li, n, m, xmin, use_k, top_k, hist_k, kscale, koff, thr, sample, orow, idx_out, prob_out);
    slot = tl.slot;
    const float xlow = tl.xlow;
    const bool incl = tl.incl != 0;

    if (mixed_out != nullptr) {
      float4* po = reinterpret_cast<float4*>(mixed_out + orow * V);
      if (incl) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          stg_stream(po + i * kThreads + tid, make_float4(x[4 * i] >= xlow ? x[4 * i] : -INFINITY, x[4 * i + 1] >= xlow ? x[4 * i + 1] : -INFINITY,
                                                           x[4 * i + 2] >= xlow ? x[4 * i + 2] : -INFINITY, x[4 * i + 3] >= xlow ? x[4 * i + 3] : -INFINITY));
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          stg_stream(po + i * kThreads + tid, make_float4(x[4 * i] > xlow ? x[4 * i] : -INFINITY, x[4 * i + 1] > xlow ? x[4 * i + 1] : -INFINITY,
                                                           x[4 * i + 2] > xlow ? x[4 * i + 2] : -INFINITY, x[4 * i + 3] > xlow ? x[4 * i + 3] : -INFINITY));
      }
    }
    __syncthreads();      // every thread is past its last read of the row's shared state
    if (tid == 0) { s.ncand = 0; s.n_list = 0; s.tkey = 0; s.minkey = 0xFFFFFFFFu; }
  }
}

// ================================================================================================
// K4
// ================================================================================================
// K4 v4.  One 256-thread CTA per token row, four CTAs per SM.  Each thread streams its 2 x 4NV values of the two logit
// rows straight into REGISTERS with 128-bit no-allocate loads (the layout K3 uses, which runs at 0.8-0.99 of HBM): no
// shared-memory staging, so occupancy is bounded by registers only and four independent rows per SM hide each other's
// load and reduction latency.  (v2/v3 staged the rows through a bulk-TMA shared-memory ring -- 64 KiB per CTA, three CTAs
// per SM -- and v2 additionally wrote the exponentials back to shared memory; v3's profile showed the kernel waiting on the
// ring and on barriers with the issue slots 55 % busy.)  Exponentials replace the logits in the registers; the owner of
// element d picks its two exponentials with a select chain (one warp pays); a rejected row resamples straight from the
// registers.  Per-(image, stage) counters are integer atomics into the caller's zeroed workspace, copied out and re-zeroed
// by the last CTA, so the launch needs no init kernel.  Arithmetic and reduction order are those of oracle/spec_c:
// sdvar_spec_verify, bit for bit.

// u / noise row of token (b, pos): dense (b*L + pos), or "stage-major" = the concatenation over stages j of (B*l_j) blocks,
// i.e. exactly the tensors a caller draws stage by stage ((B*l_j, V) each) laid end to end
__device__ __forceinline__ long long aux_row(int stage_major, int B, int L, const SegTable& seg, int b, int pos, int j) {
  if (!stage_major) return (long long)b * L + pos;
  const int lj = seg.begin[j + 1] - seg.begin[j];
  return (long long)B * seg.begin[j] + (long long)b * lj + (pos - seg.begin[j]);
}

template <int NV>
__global__ void __launch_bounds__(kThreads, NV <= 4 ? 4 : 2)
k4_verify_kernel(const float* __restrict__ xt, const float* __restrict__ xd, const long long* __restrict__ draft_idx,
                 const float* __restrict__ u, const float* __restrict__ noise, int stage_major, int B, int L, SegTable seg,
                 long long* __restrict__ out_idx, unsigned char* __restrict__ accept, float* __restrict__ p_d_out,
                 float* __restrict__ q_d_out, int* __restrict__ first_reject, int* __restrict__ n_accept,
                 int* __restrict__ accepted_stages, int* __restrict__ summary, int* ws) {
  constexpr int V = NV * 1024;
  __shared__ RedSmem sm;
  __shared__ int s_last;
  int slot = 0;
  const int tid = threadIdx.x;
  const long long rows = (long long)B * L;
  int* ws_acc = ws + 4;               // [B*S] accepted tokens per (image, stage)
  int* ws_fr = ws + 4 + B * seg.S;    // [B*S] max over rejected tokens of (l_j - position): 0 = no reject
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float4* pt = reinterpret_cast<const float4*>(xt + row * V);
    const float4* pd = reinterpret_cast<const float4*>(xd + row * V);
    // pass 1: values to registers, row maxima (order-independent)
    float4 a[NV], c[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { a[i] = ldg_stream(pt + i * kThreads + tid); c[i] = ldg_stream(pd + i * kThreads + tid); }
    const int d = (int)draft_idx[row];
    float mt = -INFINITY, md = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      mt = fmaxf(fmaxf(mt, fmaxf(a[i].x, a[i].y)), fmaxf(a[i].z, a[i].w));
      md = fmaxf(fmaxf(md, fmaxf(c[i].x, c[i].y)), fmaxf(c[i].z, c[i].w));
    }
    uint32_t kt = fkey(mt), kd = fkey(md);
    block_max_u32x2(kt, kd, sm, slot);
    mt = fkey_inv(kt);
    md = fkey_inv(kd);
    // pass 2: exponentials in place (registers), canonical sums: scalar adds in element order, packed exponentials on the
    // register pairs the 128-bit loads delivered
    float zt = 0.0f, zd = 0.0f;
    const f32x2 nmt = pk2(-mt, -mt), nmd = pk2(-md, -md);    // a - m == a + (-m) exactly
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float x0, x1;
      unpk2(add2(pk2(a[i].x, a[i].y), nmt), x0, x1);
      unpk2(spec_expf2(x0, x1), a[i].x, a[i].y);
      unpk2(add2(pk2(a[i].z, a[i].w), nmt), x0, x1);
      unpk2(spec_expf2(x0, x1), a[i].z, a[i].w);
      unpk2(add2(pk2(c[i].x, c[i].y), nmd), x0, x1);
      unpk2(spec_expf2(x0, x1), c[i].x, c[i].y);
      unpk2(add2(pk2(c[i].z, c[i].w), nmd), x0, x1);
      unpk2(spec_expf2(x0, x1), c[i].z, c[i].w);
      zt = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(zt, a[i].x), a[i].y), a[i].z), a[i].w);
      zd = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(zd, c[i].x), c[i].y), c[i].z), c[i].w);
    }
    block_sum2(zt, zd, sm, slot);  // zt, zd now hold Zt, Zd
    // the owner of element d evaluates the accept test and publishes the per-token outputs itself
    const bool owner = ((d >> 2) & (kThreads - 1)) == tid;
    int rej = 0;
    if (owner) {
      float etd = 0.0f, edd = 0.0f;
      const int di = d >> 10, dc = d & 3;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (i == di) {
          etd = dc == 0 ? a[i].x : dc == 1 ? a[i].y : dc == 2 ? a[i].z : a[i].w;
          edd = dc == 0 ? c[i].x : dc == 1 ? c[i].y : dc == 2 ? c[i].z : c[i].w;
        }
      }
      const float izt = __fdiv_rn(1.0f, zt), izd = __fdiv_rn(1.0f, zd);
      const int b = (int)(row / L), pos = (int)(row - (long long)b * L);
      const int j = seg_of(seg, pos);
      const float uu = u[aux_row(stage_major, B, L, seg, b, pos, j)];
      const float pdv = __fmul_rn(etd, izt), qdv = __fmul_rn(edd, izd);
      rej = (__fmul_rn(uu, qdv) < pdv) ? 0 : 1;
      accept[row] = (unsigned char)(rej ^ 1);
      if (p_d_out) p_d_out[row] = pdv;
      if (q_d_out) q_d_out[row] = qdv;
      if (!rej) {
        out_idx[row] = d;
        atomicAdd(&ws_acc[b * seg.S + j], 1);
      } else {
        atomicMax(&ws_fr[b * seg.S + j], seg.begin[j + 1] - pos);
      }
    }
    const int acc = __syncthreads_or(rej) ? 0 : 1;
    if (!acc) {  // block-uniform: residual resample straight from the registers
      const float izt = __fdiv_rn(1.0f, zt), izd = __fdiv_rn(1.0f, zd);
      const int b = (int)(row / L), pos = (int)(row - (long long)b * L);
      const float4* pn = reinterpret_cast<const float4*>(noise + aux_row(stage_major, B, L, seg, b, pos, seg_of(seg, pos)) * V);
      float4 nz[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) nz[i] = ldg_stream(pn + i * kThreads + tid);
      float best = -1.0f;
      int bi = 0x7FFFFFFF;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w}, cv[4] = {c[i].x, c[i].y, c[i].z, c[i].w};
        const float nn[4] = {nz[i].x, nz[i].y, nz[i].z, nz[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float r1 = __fmaf_rn(av[q], izt, -__fmul_rn(cv[q], izd));
          r1 = r1 > 0.0f ? r1 : 0.0f;
          const float r = __fmul_rn(r1, __frcp_rn(nn[q]));
          if (r > best) { best = r; bi = 4 * (i * kThreads + tid) + q; }
        }
      }
      unsigned long long w = block_max_u64(pack_best(best, bi), sm, slot);
      if ((uint32_t)(w >> 32) == fkey(0.0f)) {
        // the residual is identically zero (p == q element-wise, a reject can then only come from rounding in the accept
        // test): the spec resamples from p itself.  Detected from the winner (best ratio 0 <=> every residual is 0), so the
        // common path needs no extra block-wide vote.
        best = -1.0f;
        bi = 0x7FFFFFFF;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
          const float nn[4] = {nz[i].x, nz[i].y, nz[i].z, nz[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float r = __fmul_rn(__fmul_rn(av[q], izt), __frcp_rn(nn[q]));
            if (r > best) { best = r; bi = 4 * (i * kThreads + tid) + q; }
          }
        }
        w = block_max_u64(pack_best(best, bi), sm, slot);
      }
      int out = (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
      if (out == 0x7FFFFFFF) out = 0;
      if (tid == 0) out_idx[row] = out;
    }
  }
  // ---- last CTA finalises the per-image / batch scan (integer atomics => deterministic) and re-zeroes the workspace ----
  __threadfence();
  if (tid == 0) s_last = (atomicAdd(&ws[0], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    int mn = seg.S, na = 0;
    for (int b = tid; b < B; b += kThreads) {
      int a2 = 0;
      bool open = true;
      for (int j2 = 0; j2 < seg.S; ++j2) {
        const int lj = seg.begin[j2 + 1] - seg.begin[j2];
        const int n = __ldcg(&ws_acc[b * seg.S + j2]);
        const int f = __ldcg(&ws_fr[b * seg.S + j2]);
        ws_acc[b * seg.S + j2] = 0;
        ws_fr[b * seg.S + j2] = 0;
        n_accept[b * seg.S + j2] = n;
        first_reject[b * seg.S + j2] = lj - f;
        na += n;
        if (open && n == lj) ++a2; else open = false;
      }
      accepted_stages[b] = a2;
      mn = min(mn, a2);
    }
    // min over images via max of (S - accepted)
    uint32_t neg = (uint32_t)(seg.S - mn), z = 0;
    block_max_u32x2(neg, z, sm, slot);
    const int total_acc = block_sum_int(na, sm, slot);
    if (tid == 0) {
      summary[0] = seg.S - (int)neg;
      summary[1] = total_acc;
      summary[2] = (int)rows - total_acc;
      summary[3] = 0;
      ws[0] = 0;
    }
  }
}

// reference rule: top-1 match (models/var.py:1199-1206)
template <int NV>
__global__ void __launch_bounds__(kThreads, 4)
top1_match_kernel(const float* __restrict__ xt, const long long* __restrict__ draft_idx, int B, int L, SegTable seg,
                  unsigned char* __restrict__ match, int* n_match) {
  constexpr int V = NV * 1024;
  __shared__ RedSmem sm;
  int slot = 0;
  const int tid = threadIdx.x;
  const long long rows = (long long)B * L;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / L), pos = (int)(row - (long long)b * L);
    const float4* pt = reinterpret_cast<const float4*>(xt + row * V);
    float best = -INFINITY;
    int bi = 0x7FFFFFFF;
    bool first = true;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 a = ldg_stream(pt + i * kThreads + tid);
      const float vv[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (first || vv[c] > best) { best = vv[c]; bi = 4 * (i * kThreads + tid) + c; first = false; }
    }
    const unsigned long long w = block_max_u64(pack_best(best, bi), sm, slot);
    if (tid == 0) {
      const int am = (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
      const int mt = (am == (int)draft_idx[row]) ? 1 : 0;
      match[row] = (unsigned char)mt;
      if (mt) atomicAdd(&n_match[b * seg.S + seg_of(seg, pos)], 1);
    }
  }
}

static int fill_seg(SegTable& seg, const int* seg_begin_host, int S, int L, const float* t1, const float* t2) {
  SDVAR_REQUIRE(S >= 1 && S <= SDVAR_MAX_SEG, "S=%d out of range [1,%d]", S, SDVAR_MAX_SEG);
  SDVAR_REQUIRE(seg_begin_host != nullptr, "seg_begin_host is NULL");
  SDVAR_REQUIRE(seg_begin_host[0] == 0 && seg_begin_host[S] == L, "segment table must cover [0,L)");
  seg.S = S;
  for (int j = 0; j <= S; ++j) seg.begin[j] = seg_begin_host[j];
  for (int j = 0; j < S; ++j) {
    SDVAR_REQUIRE(seg.begin[j + 1] > seg.begin[j], "empty segment %d", j);
    seg.t1[j] = t1 ? t1[j] : 1.0f;
    seg.t2[j] = t2 ? t2[j] : 0.0f;
  }
  return SDVAR_OK;
}

static int row_grid(long long rows, int blocks_per_sm) {
  const int sms = sm_count();
  const long long g = (long long)sms * blocks_per_sm;
  return (int)(rows < g ? rows : g);
}

}  // namespace sdvar

using namespace sdvar;

extern "C" int sdvar_sample_cfg_topk_topp(const float* logits_2BLV, int B, int L, int in_ld, int in_off, int out_ld, int out_off,
                                          int V, const int* seg_begin_host, int S,
                                          const float* t1_host, const float* t2_host, int top_k, float one_minus_top_p,
                                          const float* noise, long long* idx_out, float* mixed_out, float* prob_out,
                                          void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(logits_2BLV && B > 0 && L > 0, "bad logits/B/L");
  SDVAR_REQUIRE(in_off >= 0 && in_ld >= in_off + L, "bad row mapping in_ld=%d in_off=%d L=%d", in_ld, in_off, L);
  SDVAR_REQUIRE(out_off >= 0 && out_ld >= out_off + L, "bad output row mapping out_ld=%d out_off=%d L=%d", out_ld, out_off, L);
  SDVAR_REQUIRE(V % 1024 == 0 && V >= 1024 && V <= 8192, "V=%d must be a multiple of 1024 in [1024,8192]", V);
  SDVAR_REQUIRE(((uintptr_t)logits_2BLV & 15) == 0 && ((uintptr_t)noise & 15) == 0 && ((uintptr_t)mixed_out & 15) == 0,
                "row pointers must be 16-byte aligned");
  SDVAR_REQUIRE(t1_host && t2_host, "t1/t2 are NULL");
  SegTable seg;
  if (int rc = fill_seg(seg, seg_begin_host, S, L, t1_host, t2_host)) return rc;
  const long long rows = (long long)B * L;
  const int grid = row_grid(rows, 4);
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(st, FAM_SAMPLE, (double)rows * (8.0 * V + (noise ? 4.0 * V + 8.0 : 0.0) + (mixed_out ? 4.0 * V : 0.0)));
  const bool filtered = (top_k > 0 && top_k < V) || one_minus_top_p >= 0.0f;
  static const int occ = [] { const char* e = getenv("SDVAR_K3_OCC"); return e && atoi(e) == 4 ? 4 : 3; }();   // CTAs per SM (A/B switch)
  const size_t dyn = (size_t)(V + 2) * 10;      // list: (logit, noise) fp32 pairs + vocabulary index (u16), + a dummy slot
#define SDVAR_K3(NV)                                                                                                        \
  case NV:                                                                                                                  \
    if (filtered && occ == 3) {                                                                                             \
      SDVAR_SET_SMEM_ONCE((k3_filtered_kernel<NV, 3>), dyn);                                                                \
      k3_filtered_kernel<NV, 3><<<row_grid(rows, 3), kThreads, dyn, st>>>(logits_2BLV, B, L, in_ld, in_off, out_ld, out_off, seg, top_k, \
                                                          one_minus_top_p, noise, idx_out, mixed_out, prob_out);           \
    } else if (filtered) {                                                                                                  \
      SDVAR_SET_SMEM_ONCE((k3_filtered_kernel<NV, 4>), dyn);                                                                \
      k3_filtered_kernel<NV, 4><<<grid, kThreads, dyn, st>>>(logits_2BLV, B, L, in_ld, in_off, out_ld, out_off, seg, top_k, \
                                                          one_minus_top_p, noise, idx_out, mixed_out, prob_out);           \
    } else {                                                                                                                \
      k3_plain_kernel<NV><<<grid, kThreads, 0, st>>>(logits_2BLV, B, L, in_ld, in_off, out_ld, out_off, seg, noise, idx_out, \
                                                     mixed_out, prob_out);                                                 \
    }                                                                                                                       \
    break;
  switch (V / 1024) {
    SDVAR_K3(1) SDVAR_K3(2) SDVAR_K3(4) SDVAR_K3(8)
    default:
      SDVAR_REQUIRE(false, "V=%d unsupported (1024,2048,4096,8192)", V);
  }
#undef SDVAR_K3
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

extern "C" long long sdvar_verify_workspace_bytes(int B, int S) {
  if (B <= 0 || S <= 0 || S > SDVAR_MAX_SEG) return SDVAR_ERR_ARG;
  return (long long)sizeof(int) * (4 + 2LL * B * S);
}

extern "C" int sdvar_verify_accept_resample(const float* xt, const float* xd, const long long* draft_idx, const float* u,
                                            const float* noise, int stage_major_aux, int B, int L, int V,
                                            const int* seg_begin_host, int S,
                                            long long* out_idx, unsigned char* accept, float* p_d_out, float* q_d_out,
                                            int* first_reject, int* n_accept, int* accepted_stages, int* summary,
                                            int* workspace, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(xt && xd && draft_idx && u && noise && out_idx && accept && first_reject && n_accept && accepted_stages &&
                    summary && workspace,
                "NULL argument");
  SDVAR_REQUIRE(B > 0 && L > 0, "bad B/L");
  SDVAR_REQUIRE(V % 1024 == 0 && V >= 1024 && V <= 8192, "V=%d must be a multiple of 1024 in [1024,8192]", V);
  SDVAR_REQUIRE(((uintptr_t)xt & 15) == 0 && ((uintptr_t)xd & 15) == 0 && ((uintptr_t)noise & 15) == 0, "16-byte alignment");
  SegTable seg;
  if (int rc = fill_seg(seg, seg_begin_host, S, L, nullptr, nullptr)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(st, FAM_VERIFY, (double)B * L * (8.0 * V + 17.0));
  const int grid = row_grid((long long)B * L, V <= 4096 ? 4 : 2);
#define SDVAR_K4(NV)                                                                                                 \
  case NV: {                                                                                                         \
    k4_verify_kernel<NV><<<grid, kThreads, 0, st>>>(xt, xd, draft_idx, u, noise, stage_major_aux, B, L, seg, out_idx, \
                                                          accept, p_d_out, q_d_out, first_reject, n_accept,         \
                                                          accepted_stages, summary, workspace);                     \
  } break;
  switch (V / 1024) {
    SDVAR_K4(1) SDVAR_K4(2) SDVAR_K4(4) SDVAR_K4(8)
    default:
      SDVAR_REQUIRE(false, "V=%d unsupported", V);
  }
#undef SDVAR_K4
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

extern "C" int sdvar_verify_top1(const float* xt, const long long* draft_idx, int B, int L, int V,
                                 const int* seg_begin_host, int S, unsigned char* match, int* n_match, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(xt && draft_idx && match && n_match, "NULL argument");
  SDVAR_REQUIRE(V % 1024 == 0 && V >= 1024 && V <= 8192, "V=%d must be a multiple of 1024 in [1024,8192]", V);
  SegTable seg;
  if (int rc = fill_seg(seg, seg_begin_host, S, L, nullptr, nullptr)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  SDVAR_CUDA(cudaMemsetAsync(n_match, 0, sizeof(int) * B * S, st));
  const int grid = row_grid((long long)B * L, 4);
  switch (V / 1024) {
    case 1: top1_match_kernel<1><<<grid, kThreads, 0, st>>>(xt, draft_idx, B, L, seg, match, n_match); break;
    case 2: top1_match_kernel<2><<<grid, kThreads, 0, st>>>(xt, draft_idx, B, L, seg, match, n_match); break;
    case 4: top1_match_kernel<4><<<grid, kThreads, 0, st>>>(xt, draft_idx, B, L, seg, match, n_match); break;
    case 8: top1_match_kernel<8><<<grid, kThreads, 0, st>>>(xt, draft_idx, B, L, seg, match, n_match); break;
    default: SDVAR_REQUIRE(false, "V=%d unsupported", V);
  }
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

// y_packed / y_scalar [n] = the kernels' exponential (fp32x2 and scalar code paths) of x[n] <= 0: bit-exactness test hook
extern "C" int sdvar_debug_spec_expf(const float* x, long long n, float* y_packed, float* y_scalar, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && y_packed && y_scalar && n > 0, "bad argument");
  spec_expf_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(x, n, y_packed, y_scalar);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}
