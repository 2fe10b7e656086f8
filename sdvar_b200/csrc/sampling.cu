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
//   k3_filtered_kernel  top-k and/or top-p.  Round 1 searched both cuts by bisection (~26 block-wide barrier rounds, 0.15-0.23
//                       of HBM); round 2 (first half) compacted the survivors of a histogram cut into shared memory and worked
//                       on the list (16 barriers, ~1950 instructions per warp-row, 0.28-0.37).  v3, this one (0.37-0.47):
//     1. the row never leaves the registers: 128 threads x 32 entries for V <= 4096 (the cut-location phases cost every thread
//        the same however many entries it owns, so fewer, fatter threads halve that overhead per row; 128 registers, no spills).
//        ONE pass computes every exponential (packed fp32x2) and adds each entry to three shared-memory histograms over 1024
//        value bins: count, and the two 20-bit limbs of E = rint(e * 2^40).  Bin offset and limbs come from magic-addend
//        roundings on the packed pipe -- no float<->int conversion, no division;
//     2. one block scan over the bins (from the top) gives, for every bin, the entries and the mass above it, hence the bin
//        the top-k cut falls into AND the few bins the top-p cut can fall into (the spec states the top-p test on the
//        fixed-point numerators: E{x_w > x_v} >= Zi - floor(Zi * thr), no per-entry probability);
//     3. the entries of those bins (~20-40 of the row) are the only ones looked at individually: all-pairs ranks and sums
//        among them, NT / P threads per candidate, settle both cuts exactly (tie groups whole) and yield Zi and the final
//        sum Z2i without another pass (tests/test_k3_model.py restates this logic on the CPU against the definition);
//     4. the race runs on the numerators, argmax e * R(noise), R the spec's division-free reciprocal (packed Newton steps);
//        the noise row is fetched into shared memory by cp.async at row start.
//     Every sum that decides something is an integer sum of fixed-point terms, so the atomics may run in any order and the
//     kernel is bit-exact to oracle/spec_c.  9 barriers per row.  Rows the histograms cannot take (non-finite or zero value
//     range, more candidates than threads, top_p == 0) go through k3_row_slow (bit-serial searches, any input).
//     What bounds it now (profiles/ncu_k3_r02b.md): ~2400 instructions per warp-row at 16 resident warps and 12 288
//     shared-memory atomics per row (~2 lanes per clock per SM measured) -- instruction issue and the atomics each account for
//     about the whole row time; removing either alone (pruned histograms / branch-free code) gave 0-5 %.
constexpr int kBins = 1024;
constexpr int kCandMax = 256;
constexpr int kRcpSteps = 4;                // SDVAR_RCP_STEPS of the spec
constexpr float kFixE = 1099511627776.0f;   // 2^40
constexpr float kFixM = 1073741824.0f;      // 2^30
constexpr float kMagic = 12582912.0f;       // 1.5 * 2^23: x + kMagic rounds x to an integer that sits in the low mantissa bits
constexpr uint32_t kMagicBits = 0x4B400000u;
constexpr float kMagic4 = 3145728.0f;        // 1.5 * 2^21: one ulp is 1/4 there
constexpr uint32_t kMagic4Bits = 0x4A400000u;

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

// ---- K3 filtered, v3: the row stays in registers; ONE pass builds three shared-memory histograms ----
struct K3Smem {
  __align__(16) uint32_t hcnt[kBins];          // entries per value bin
  __align__(16) uint32_t hhi[kBins];           // sum of the raw high limbs (each carries the magic exponent bits, removed in the scan)
  __align__(16) uint32_t hlo[kBins];           // sum of the raw low limbs
  unsigned long long pdex[kBins + 1];          // pdex[b + 1] = E of all bins above b; pdex[0] = E of the whole row
  uint32_t pcab[kBins];                        // entries in the bins above b
  unsigned long long cE[kCandMax];             // candidates: fixed-point exponential,
  float cx[kCandMax];                          //   logit,
  unsigned char cg[kCandMax];                  //   group (0: the top-k crossing bin on its own, 1: the top-p bin range)
  unsigned long long wE[kWarps];
  uint32_t wc[kWarps];
  __align__(16) unsigned long long zpart[2][kWarps];
  unsigned long long Zi, Z2i;
  uint32_t tkey, minkey;
  int bK, bPlo, bPhi, nc, slow;
  float xK;
  RedSmem red;
};

// fixed-point exponential E = rint(e * 2^40) as two 20-bit limbs (hi * 2^20 + lo); e in [0, 1]
__device__ __forceinline__ void fix_e(float e, uint32_t& hi, uint32_t& lo) {
  const float a = __fmul_rn(e, 1048576.0f);                 // e * 2^20, exact
  const float fh = floorf(a);
  hi = (uint32_t)fh;
  lo = __float2uint_rn(__fmul_rn(__fsub_rn(a, fh), 1048576.0f));   // both operations exact; one rounding in the conversion
}
// removable mass of the top-p cut: floor(Zi * thr_fix / 2^30), Zi < 2^53, thr_fix <= 2^30
__device__ __forceinline__ unsigned long long thr_mass(unsigned long long Zi, uint32_t thr_fix) {
  const unsigned long long lo = Zi * (unsigned long long)thr_fix, hi = __umul64hi(Zi, (unsigned long long)thr_fix);
  return (hi << 34) | (lo >> 30);
}
// (R(n0), R(n1)): the spec's division-free reciprocal (sdvar_spec_rcp), two at a time on the packed pipe
__device__ __forceinline__ f32x2 spec_rcp2(float n0, float n1) {
  f32x2 y = pk2(u2f(0x7EF311C7u - f2u(n0)), u2f(0x7EF311C7u - f2u(n1)));
  const f32x2 nn = pk2(u2f(f2u(n0) ^ 0x80000000u), u2f(f2u(n1) ^ 0x80000000u)), one = splat2(1.0f);
#pragma unroll
  for (int i = 0; i < kRcpSteps; ++i) {
    const f32x2 t = fma2(nn, y, one);
    y = fma2(y, t, y);
  }
  return y;
}

// ---- block reductions of the filtered kernel, NW warps per CTA (128 threads x 32 entries for V <= 4096: the cut-location
// phases cost every THREAD the same whatever the row length per thread, so half the threads halve that overhead per row)
template <int NW>
__device__ __forceinline__ void k3_block_max2(uint32_t& a, uint32_t& b, RedSmem& s, int& slot) {
  a = __reduce_max_sync(0xffffffffu, a);
  b = __reduce_max_sync(0xffffffffu, b);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s.u[slot][0][w] = a; s.u[slot][1][w] = b; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NW; ++k) { a = max(a, s.u[slot][0][k]); b = max(b, s.u[slot][1][k]); }
  slot ^= 1;
}
template <int NW>
__device__ __forceinline__ int k3_block_sum_int(int c, RedSmem& s, int& slot) {
  c = __reduce_add_sync(0xffffffffu, c);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s.i[slot][w] = c;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) t += s.i[slot][k];
  slot ^= 1;
  return t;
}
template <int NW>
__device__ __forceinline__ unsigned long long k3_block_max_u64(unsigned long long v, RedSmem& s, int& slot) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
    v = o > v ? o : v;
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) s.u64[slot][w] = v;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NW; ++k) v = s.u64[slot][k] > v ? s.u64[slot][k] : v;
  slot ^= 1;
  return v;
}
// block sum of per-thread limb sums -> u64 total (integer sums: any order gives the same total; per-thread sums < 2^26, warp < 2^31)
template <int NW>
__device__ __forceinline__ unsigned long long k3_block_sum_fix(uint32_t hi, uint32_t lo, K3Smem& s, int buf) {
  hi = __reduce_add_sync(0xffffffffu, hi);
  lo = __reduce_add_sync(0xffffffffu, lo);
  if ((threadIdx.x & 31) == 0) s.zpart[buf][threadIdx.x >> 5] = ((unsigned long long)hi << 20) + (unsigned long long)lo;
  __syncthreads();
  unsigned long long t = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) t += s.zpart[buf][k];
  return t;
}

// ---- generic path of one row (block-wide, any input): bit-serial searches on the order-preserving keys with one block
// reduction per bit.  ~70 barriers: only for rows the histograms cannot take (non-finite or zero value range, more than
// threads candidates around a cut, a bin with more than 4095 entries, top_p == 0).  Re-reads the row so that the caller's
// registers stay out of local memory.  Returns the cut as a key (keep <=> key(x) >= klow) and the final fixed-point sum.
struct K3Cut {
  uint32_t klow;
  unsigned long long Z2i;
  int slot;
};
template <int NV, int NT>
__device__ __noinline__ K3Cut k3_row_slow(K3Smem& s, int slot, const float4* __restrict__ pc, const float4* __restrict__ pu, float t1,
                                          float t2, bool use_k, int top_k, bool use_p, uint32_t thr_fix, uint32_t kmx) {
  constexpr int CH = NV * 256 / NT, E = CH * 4, NW = NT / 32;
  const int tid = threadIdx.x;
  uint32_t key[E], hi[E], lo[E];
  float x[E];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const float4 a = ldg_stream(pc + i * NT + tid), c = ldg_stream(pu + i * NT + tid);
    x[4 * i + 0] = __fsub_rn(__fmul_rn(a.x, t1), __fmul_rn(c.x, t2));
    x[4 * i + 1] = __fsub_rn(__fmul_rn(a.y, t1), __fmul_rn(c.y, t2));
    x[4 * i + 2] = __fsub_rn(__fmul_rn(a.z, t1), __fmul_rn(c.z, t2));
    x[4 * i + 3] = __fsub_rn(__fmul_rn(a.w, t1), __fmul_rn(c.w, t2));
  }
#pragma unroll
  for (int e = 0; e < E; ++e) key[e] = fkey(__fadd_rn(x[e], 0.0f));
  uint32_t K = 0;
  if (use_k) {
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t tr = K | (1u << bit);
      int c = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) c += (key[e] >= tr) ? 1 : 0;
      c = k3_block_sum_int<NW>(c, s.red, slot);
      if (c >= top_k) K = tr;
    }
  }
  const float m = fkey_inv(kmx);
  uint32_t sh = 0, sl = 0;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const float ee = (key[e] >= K) ? spec_expf(__fsub_rn(x[e], m)) : 0.0f;
    fix_e(ee, hi[e], lo[e]);
    sh += hi[e];
    sl += lo[e];
  }
  int buf = 0;
  const unsigned long long Zi = k3_block_sum_fix<NW>(sh, sl, s, buf);
  buf ^= 1;
  uint32_t klow = K;
  unsigned long long Z2i = Zi;
  if (use_p) {
    const unsigned long long thrE = thr_mass(Zi, thr_fix);
    uint32_t T = 0;      // largest key value with E{alive, key <= T} <= thrE
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t tr = T | (1u << bit);
      sh = 0; sl = 0;
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (key[e] <= tr) { sh += hi[e]; sl += lo[e]; }      // entries below the top-k cut carry zero limbs
      const unsigned long long a = k3_block_sum_fix<NW>(sh, sl, s, buf);
      buf ^= 1;
      if (a <= thrE) T = tr;
    }
    const uint32_t cut = T >= kmx ? kmx : T + 1u;            // removed <=> key <= T and key != key(max)
    klow = max(K, cut);
    sh = 0; sl = 0;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (key[e] >= klow) { sh += hi[e]; sl += lo[e]; }
    Z2i = k3_block_sum_fix<NW>(sh, sl, s, buf);
  }
  return K3Cut{klow, Z2i, slot};
}

template <int NV, int NT, int OCC>
__global__ void __launch_bounds__(NT, OCC)
k3_filtered_kernel(const float* __restrict__ logits, int B, int L, int in_ld, int in_off, int out_ld, int out_off, SegTable seg,
                   int top_k, float thr, const float* __restrict__ noise, long long* __restrict__ idx_out,
                   float* __restrict__ mixed_out, float* __restrict__ prob_out) {
  constexpr int V = NV * 1024;
  constexpr int CH = NV * 256 / NT;          // float4 chunks per thread: thread t owns chunks f = i * NT + t
  constexpr int E = CH * 4;                  // entries per thread (<= 32: one flag bit each)
  constexpr int NW = NT / 32;
  constexpr int BPT = kBins / NT;            // bins per thread in the scan
  constexpr int kCand = NT < kCandMax ? NT : kCandMax;
  static_assert(E <= 32 && BPT % 4 == 0, "k3_filtered_kernel shape");
  __shared__ K3Smem s;
  extern __shared__ __align__(16) unsigned char k3_noise[];      // [V] floats: the row's noise, fetched by cp.async at row start
  int slot = 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const long long rows = (long long)B * L;
  for (int i = tid; i < kBins; i += NT) { s.hcnt[i] = 0; s.hhi[i] = 0; s.hlo[i] = 0; }
  if (tid == 0) { s.nc = 0; s.slow = 0; s.tkey = 0; s.minkey = 0xFFFFFFFFu; s.bPlo = 0x7FFFFFFF; s.bPhi = -1; }
  __syncthreads();
  const bool sample = noise != nullptr;
  const bool use_k = top_k > 0 && top_k < V;
  const bool use_p = thr >= 0.0f;
  const uint32_t thr_fix = use_p ? (uint32_t)__fmul_rn(thr, kFixM) : 0u;      // floor
  const bool small = rows < 0x7FFFFFFFLL;
  const int q0 = kBins - BPT - BPT * tid;   // this thread scans bins q0+BPT-1 .. q0 (descending)
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
      float4 a[CH], c[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) { a[i] = ldg_stream(pc + i * NT + tid); c[i] = ldg_stream(pu + i * NT + tid); }
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        x[4 * i + 0] = __fsub_rn(__fmul_rn(a[i].x, t1), __fmul_rn(c[i].x, t2));
        x[4 * i + 1] = __fsub_rn(__fmul_rn(a[i].y, t1), __fmul_rn(c[i].y, t2));
        x[4 * i + 2] = __fsub_rn(__fmul_rn(a[i].z, t1), __fmul_rn(c[i].z, t2));
        x[4 * i + 3] = __fsub_rn(__fmul_rn(a[i].w, t1), __fmul_rn(c[i].w, t2));
      }
    }
    if (sample) {      // every thread fetches (and later reads) its own chunks: no barrier needed, only the wait
      const float4* pn = reinterpret_cast<const float4*>(noise + row * V);
#pragma unroll
      for (int i = 0; i < CH; ++i)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(k3_noise) + (uint32_t)(i * NT + tid) * 16u),
                     "l"(pn + i * NT + tid)
                     : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- row max / min ----
    float mx = -INFINITY, mn = INFINITY;
#pragma unroll
    for (int e = 0; e < E; ++e) { mx = fmaxf(mx, x[e]); mn = fminf(mn, x[e]); }
    uint32_t kmx = fkey(__fadd_rn(mx, 0.0f)), kmnc = ~fkey(__fadd_rn(mn, 0.0f));
    k3_block_max2<NW>(kmx, kmnc, s.red, slot);
    const float m = fkey_inv(kmx), xmin = fkey_inv(~kmnc);

    // ---- value bins.  tb(x) = fma(x, scale, offm) lands in [2^21, 2^22), where one ulp is 1/4: bits(tb) - bits(kMagic4) counts
    // quarter bins, so (bits(tb) & 0xffc) IS the byte offset of bin(x) in a histogram -- one LOP per entry, no float->int
    // conversion, monotone in x.  Bins only LOCATE the cuts; every sum that decides something is an exact integer, so the
    // binning is not part of the spec.
    // (Measured and dropped: leaving entries below mean + z sigma out of the histograms, verified by the scan's total.  It cut
    // the shared-memory atomics 3.3x but ptxas wraps every predicated ATOMS in a branch of its own -- +12 instructions per
    // entry -- and the atomics were not the limiter: 4 % faster with the branches, slower than this version without.) ----
    bool fast = false;
    float scale = 0.0f, offm = 0.0f;
    if (__fsub_rn(m, xmin) > 0.0f && __fsub_rn(m, xmin) < INFINITY && thr_fix < (1u << 30)) {
      scale = __fdividef((float)(kBins - 3), __fsub_rn(m, xmin));          // any value will do as long as the checks below hold
      offm = __fmaf_rn(-xmin, scale, kMagic4 + 1.0f);
      const uint32_t blo = f2u(__fmaf_rn(xmin, scale, offm)) - kMagic4Bits, bhi = f2u(__fmaf_rn(m, scale, offm)) - kMagic4Bits;
      fast = __fmul_rn(fmaxf(fabsf(xmin), fabsf(m)), scale) < 262144.0f && blo < 4u * kBins && bhi < 4u * kBins;
    }

    // ---- pass 1: exponentials (kept in registers), fixed-point limbs, three histogram atomics per entry ----
    float ev[E];
    {
      const f32x2 nm2 = splat2(-m);
      if (fast) {
        const f32x2 sc2 = splat2(scale), of2 = splat2(offm), p20 = splat2(1048576.0f), mg = splat2(kMagic), nmg = splat2(-kMagic);
#pragma unroll
        for (int q = 0; q < E; q += 2) {
          const f32x2 xp = pk2(x[q], x[q + 1]);
          float d0, d1;
          unpk2(add2(xp, nm2), d0, d1);                       // x + (-m) == x - m exactly
          const f32x2 e2 = spec_expf2(d0, d1);
          unpk2(e2, ev[q], ev[q + 1]);
          // E = rint(e * 2^40) = hi * 2^20 + lo with hi = rint(e * 2^20), lo = rint((e * 2^20 - hi) * 2^20) in [-2^19, 2^19]:
          // both roundings by the magic addend; bits(th) = kMagicBits + hi, bits(tl) = kMagicBits + lo
          const f32x2 a2 = mul2(e2, p20);
          const f32x2 th = add2(a2, mg);
          const f32x2 hf = add2(th, nmg);
          const f32x2 dd = fma2(hf, splat2(-1.0f), a2);
          const f32x2 tl = fma2(dd, p20, mg);
          const f32x2 tb = fma2(xp, sc2, of2);
          float tb0, tb1, th0, th1, tl0, tl1;
          unpk2(tb, tb0, tb1);
          unpk2(th, th0, th1);
          unpk2(tl, tl0, tl1);
          // (the mask only drops the quarter-bin bits of a finite row; it also keeps a NaN from addressing outside the arrays)
          const uint32_t o0 = (f2u(tb0) & (4u * kBins - 4u)) >> 2, o1 = (f2u(tb1) & (4u * kBins - 4u)) >> 2;
          atomicAdd(&s.hcnt[o0], 1u);
          atomicAdd(&s.hhi[o0], f2u(th0));
          atomicAdd(&s.hlo[o0], f2u(tl0));
          atomicAdd(&s.hcnt[o1], 1u);
          atomicAdd(&s.hhi[o1], f2u(th1));
          atomicAdd(&s.hlo[o1], f2u(tl1));
        }
      } else {
#pragma unroll
        for (int q = 0; q < E; q += 2) {
          float d0, d1;
          unpk2(add2(pk2(x[q], x[q + 1]), nm2), d0, d1);
          unpk2(spec_expf2(d0, d1), ev[q], ev[q + 1]);
        }
      }
    }

    // ---- scan of the bins from the top: entries / E above every bin; the top-k crossing bin ----
    unsigned long long Etot = 0;
    if (fast) {
      __syncthreads();
      uint32_t cc[BPT], hh[BPT], ll[BPT];      // index k <-> bin q0 + BPT - 1 - k (descending)
#pragma unroll
      for (int g = 0; g < BPT / 4; ++g) {
        const int qq = q0 + BPT - 4 - 4 * g;
        const uint4 c4 = *reinterpret_cast<const uint4*>(&s.hcnt[qq]);
        const uint4 h4 = *reinterpret_cast<const uint4*>(&s.hhi[qq]);
        const uint4 l4 = *reinterpret_cast<const uint4*>(&s.hlo[qq]);
        *reinterpret_cast<uint4*>(&s.hcnt[qq]) = make_uint4(0u, 0u, 0u, 0u);      // ready for the next row
        *reinterpret_cast<uint4*>(&s.hhi[qq]) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(&s.hlo[qq]) = make_uint4(0u, 0u, 0u, 0u);
        cc[4 * g] = c4.w; cc[4 * g + 1] = c4.z; cc[4 * g + 2] = c4.y; cc[4 * g + 3] = c4.x;
        hh[4 * g] = h4.w; hh[4 * g + 1] = h4.z; hh[4 * g + 2] = h4.y; hh[4 * g + 3] = h4.x;
        ll[4 * g] = l4.w; ll[4 * g + 1] = l4.z; ll[4 * g + 2] = l4.y; ll[4 * g + 3] = l4.x;
      }
      unsigned long long Eb[BPT];
      uint32_t ct = 0;
      unsigned long long Et = 0;
#pragma unroll
      for (int k = 0; k < BPT; ++k) {
        const uint32_t H = hh[k] - cc[k] * kMagicBits;                          // exact mod 2^32: the true sums fit (cc <= 4095)
        const int Lo = (int)(ll[k] - cc[k] * kMagicBits);
        Eb[k] = ((unsigned long long)H << 20) + (unsigned long long)(long long)Lo;
        ct += cc[k];
        Et += Eb[k];
      }
      uint32_t ci = ct;
      unsigned long long Ei = Et;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t oc = __shfl_up_sync(0xffffffffu, ci, off);
        const unsigned long long oE = __shfl_up_sync(0xffffffffu, Ei, off);
        if (lane >= off) { ci += oc; Ei += oE; }
      }
      if (lane == 31) { s.wc[wid] = ci; s.wE[wid] = Ei; }
      __syncthreads();
      uint32_t cb = ci - ct;
      unsigned long long eb = Ei - Et;
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const uint32_t wc = s.wc[k];
        const unsigned long long wE = s.wE[k];
        if (k < wid) { cb += wc; eb += wE; }
        Etot += wE;
      }
#pragma unroll
      for (int k = 0; k < BPT; ++k) {
        const int bin = q0 + BPT - 1 - k;
        s.pcab[bin] = cb;
        s.pdex[bin + 1] = eb;
        if (cc[k] > 4095u) s.slow = 1;
        if (use_k && cb < (uint32_t)top_k && (uint32_t)top_k <= cb + cc[k]) s.bK = bin;
        cb += cc[k];
        eb += Eb[k];
      }
      if (tid == NT - 1) s.pdex[0] = eb;      // == Etot
      __syncthreads();
      fast = s.slow == 0;
    }

    // ---- the bins a top-p crossing can fall into, whatever the exact top-k cut inside bin bK turns out to be ----
    int bK = -2, plo = kBins, phi = kBins;
    if (fast) {
      if (use_k) bK = s.bK;
      if (use_p) {
        const unsigned long long zmin = use_k ? s.pdex[bK + 1] : Etot, zmax = use_k ? s.pdex[bK] : Etot;
        const unsigned long long tmin = zmin - thr_mass(zmin, thr_fix), tmax = zmax - thr_mass(zmax, thr_fix);
        unsigned long long din = s.pdex[q0 + BPT];      // E above this thread's top bin
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
          const int bin = q0 + BPT - 1 - k;
          const unsigned long long dex = din;
          din = s.pdex[bin];                          // E above, this bin included
          if (din != dex && dex < tmax && din >= tmin) { atomicMin(&s.bPlo, bin); atomicMax(&s.bPhi, bin); }
        }
        __syncthreads();
        plo = s.bPlo;
        phi = s.bPhi;
      }
    }

    // ---- candidates: the entries of bin bK (group 0, only when it lies below the top-p range) and of bins plo..phi (group 1) ----
    int nc = 0;
    bool kgroup = false;
    if (fast) {
      kgroup = use_k && bK < plo;
      // in quarter-bin units relative to bits(tb): bin == bK <=> bits - cK < 4; plo <= bin <= phi <=> bits - cP < 4 * (phi - plo + 1);
      const uint32_t cK = kMagic4Bits + 4u * (uint32_t)bK, wK = kgroup ? 4u : 0u;
      const uint32_t cP = kMagic4Bits + 4u * (uint32_t)plo, wP = 4u * (uint32_t)(phi - plo + 1), span = (uint32_t)(phi - plo);
      uint32_t mask = 0;
      {
        const f32x2 sc2 = splat2(scale), of2 = splat2(offm);
#pragma unroll
        for (int q = 0; q < E; q += 2) {
          float tb0, tb1;
          unpk2(fma2(pk2(x[q], x[q + 1]), sc2, of2), tb0, tb1);
          if (f2u(tb0) - cK < wK || f2u(tb0) - cP < wP) mask |= 1u << q;
          if (f2u(tb1) - cK < wK || f2u(tb1) - cP < wP) mask |= 2u << q;
        }
      }
      while (mask) {                        // rare: ~20-40 entries of the row; the entry is re-read to keep x[] / ev[] in registers
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        const int f = (i >> 2) * NT + tid;
        const float xc = __ldg(reinterpret_cast<const float*>(pc + f) + (i & 3)), xu = __ldg(reinterpret_cast<const float*>(pu + f) + (i & 3));
        const float xe = __fsub_rn(__fmul_rn(xc, t1), __fmul_rn(xu, t2));
        uint32_t h, l;
        fix_e(spec_expf(__fsub_rn(xe, m)), h, l);
        const uint32_t o = (f2u(__fmaf_rn(xe, scale, offm)) - kMagic4Bits) >> 2;
        const int p = atomicAdd(&s.nc, 1);
        if (p < kCand) {
          s.cx[p] = xe;
          s.cE[p] = ((unsigned long long)h << 20) + (unsigned long long)l;
          s.cg[p] = (o - (uint32_t)plo <= span) ? 1 : 0;
        }
      }
      __syncthreads();
      nc = s.nc;
      fast = nc <= kCand;
    }

    // ---- exact cuts among the candidates: all-pairs ranks / sums, NT / P threads per candidate ----
    uint32_t klow = 0;
    unsigned long long Z2i = 0;
    bool z2_shared = false;
    if (fast) {
      int P = 8, lg = NT == 128 ? 4 : 5;     // lg = log2(threads per candidate) = log2(NT / P)
      while (P < nc) { P <<= 1; --lg; }
      const int G = 1 << lg, i = tid >> lg, sub = tid & (G - 1);
      float xi = 0.0f;
      int gi = 0;
      uint32_t cnt = 0;                      // (entries >= x_i) + 65536 * (entries > x_i), same group
      unsigned long long Ege = 0;            // E of the entries >= x_i, same group
      if (i < nc) {
        xi = s.cx[i];
        gi = s.cg[i];
        for (int jj = sub; jj < nc; jj += G) {
          const float xj = s.cx[jj];
          const bool same = s.cg[jj] == gi;
          const bool ge = same && xj >= xi, gt = same && xj > xi;
          cnt += (ge ? 1u : 0u) + (gt ? 65536u : 0u);
          Ege += ge ? s.cE[jj] : 0ull;
        }
      }
      for (int off = G >> 1; off >= 1; off >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        Ege += __shfl_xor_sync(0xffffffffu, Ege, off);
      }
      const bool lead = i < nc && sub == 0;
      const uint32_t cge = cnt & 0xFFFFu, cgt = cnt >> 16;
      unsigned long long Egt = 0, baseE = 0;
      if (lead) {
        Egt = Ege - (unsigned long long)(cge - cgt) * s.cE[i];      // ties share x, hence e, hence E
        baseE = gi ? s.pdex[phi + 1] : s.pdex[bK + 1];
        if (use_k && (gi == 0 || !kgroup)) {
          const uint32_t base = gi ? s.pcab[phi] : s.pcab[bK];
          if (base + cgt < (uint32_t)top_k && (uint32_t)top_k <= base + cge) { s.xK = xi; s.Zi = baseE + Ege; }
        }
      }
      float xK = xmin;
      unsigned long long Zi = Etot;
      if (use_k) {
        __syncthreads();
        xK = s.xK;
        Zi = s.Zi;
      }
      klow = fkey(__fadd_rn(xK, 0.0f));
      Z2i = Zi;
      if (use_p) {
        const unsigned long long target = Zi - thr_mass(Zi, thr_fix);      // removed <=> E{alive, x_w > x_v} >= target
        const uint32_t ki = fkey(__fadd_rn(xi, 0.0f));
        const bool removed = lead && gi == 1 && xi >= xK && xi < m && baseE + Egt >= target;
        if (removed) atomicMax(&s.tkey, ki);
        if (lead && gi == 1) atomicMin(&s.minkey, ki);
        __syncthreads();
        const uint32_t tk = s.tkey;
        if (tk != 0u) {
          klow = tk + 1u;                    // tk < key(max): no overflow; tk >= key(xK): removed entries are alive
          if (removed && ki == tk) s.Z2i = baseE + Egt;      // read by the winner of the race, behind its barrier
          z2_shared = true;
        } else if (plo < kBins && (use_k ? kgroup : (plo > 0 && s.pcab[plo - 1] < (uint32_t)V))) {      // (plo >= kBins: NaN row, no crossing bin)
          // no candidate is removed, but alive entries lie below bin plo, and every one of those is (E above them >=
          // E{bins >= plo} >= target): the cut sits right below the lowest entry of bin plo
          klow = s.minkey;
          Z2i = s.pdex[plo];
        }
      }
    } else {
      const K3Cut c = k3_row_slow<NV, NT>(s, slot, pc, pu, t1, t2, use_k, top_k, use_p, thr_fix, kmx);
      klow = c.klow;
      Z2i = c.Z2i;
      slot = c.slot;
      // re-materialise the row behind the call instead of keeping 2 E registers alive across it (which would put spill
      // stores on the fast path)
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const float4 a = ldg_stream(pc + i * NT + tid), cu = ldg_stream(pu + i * NT + tid);
        x[4 * i + 0] = __fsub_rn(__fmul_rn(a.x, t1), __fmul_rn(cu.x, t2));
        x[4 * i + 1] = __fsub_rn(__fmul_rn(a.y, t1), __fmul_rn(cu.y, t2));
        x[4 * i + 2] = __fsub_rn(__fmul_rn(a.z, t1), __fmul_rn(cu.z, t2));
        x[4 * i + 3] = __fsub_rn(__fmul_rn(a.w, t1), __fmul_rn(cu.w, t2));
      }
      const f32x2 nm2 = splat2(-m);
#pragma unroll
      for (int q = 0; q < E; q += 2) {
        float d0, d1;
        unpk2(add2(pk2(x[q], x[q + 1]), nm2), d0, d1);
        unpk2(spec_expf2(d0, d1), ev[q], ev[q + 1]);
      }
    }
    // one threshold for everything below: keep <=> x >= xthr (finite: a -inf logit is never in the race)
    const float xthr = fkey_inv(max(klow, 0x00800000u));

    if (sample) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      float4 nz[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) nz[i] = reinterpret_cast<const float4*>(k3_noise)[i * NT + tid];
      float best = -1.0f;
      int bi = 0x7FFFFFFF;
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        float r0, r1, r2, r3;
        unpk2(mul2(pk2(ev[4 * i], ev[4 * i + 1]), spec_rcp2(nz[i].x, nz[i].y)), r0, r1);
        unpk2(mul2(pk2(ev[4 * i + 2], ev[4 * i + 3]), spec_rcp2(nz[i].z, nz[i].w)), r2, r3);
        const float rr[4] = {r0, r1, r2, r3};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float r = (x[4 * i + c] >= xthr) ? rr[c] : -1.0f;
          if (r > best) { best = r; bi = 4 * (i * NT + tid) + c; }
        }
      }
      const unsigned long long w = k3_block_max_u64<NW>(pack_best(best, bi), s.red, slot);
      int win = (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
      const bool none = (uint32_t)(w >> 32) == fkey(-1.0f);      // nothing alive (a row of -inf / NaN): index 0, probability 0
      if (none) win = 0;
      if (none ? tid == 0 : bi == win) {
        if (idx_out) idx_out[orow] = win;
        if (prob_out) {
          float pw = 0.0f;
          if (!none) {
            const int f = win >> 2, c = win & 3;
            const float xc = __ldg(reinterpret_cast<const float*>(pc + f) + c), xu = __ldg(reinterpret_cast<const float*>(pu + f) + c);
            const float ew = spec_expf(__fsub_rn(__fsub_rn(__fmul_rn(xc, t1), __fmul_rn(xu, t2)), m));
            const unsigned long long z = z2_shared ? s.Z2i : Z2i;
            pw = __fdiv_rn(ew, __fmul_rn(__ull2float_rn(z), 1.0f / kFixE));
          }
          prob_out[orow] = pw;
        }
      }
    }
    if (mixed_out != nullptr) {
      float4* po = reinterpret_cast<float4*>(mixed_out + orow * V);
#pragma unroll
      for (int i = 0; i < CH; ++i)
        stg_stream(po + i * NT + tid, make_float4(x[4 * i] >= xthr ? x[4 * i] : -INFINITY, x[4 * i + 1] >= xthr ? x[4 * i + 1] : -INFINITY,
                                                         x[4 * i + 2] >= xthr ? x[4 * i + 2] : -INFINITY, x[4 * i + 3] >= xthr ? x[4 * i + 3] : -INFINITY));
    }
    if (!sample) __syncthreads();      // (the race's reduction is the barrier otherwise) every thread is past its reads of the row's shared state
    if (tid == 0) { s.nc = 0; s.slow = 0; s.tkey = 0; s.minkey = 0xFFFFFFFFu; s.bPlo = 0x7FFFFFFF; s.bPhi = -1; }
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
  static const int nt = [] { const char* e = getenv("SDVAR_K3_THREADS"); return e && atoi(e) == 256 ? 256 : 128; }();   // threads per row (A/B switch)
  const size_t dyn = (size_t)V * 4;      // the filtered kernel's noise row
#define SDVAR_K3(NV)                                                                                                        \
  case NV:                                                                                                                  \
    if (filtered && nt == 128 && NV <= 4) {                                                                                 \
      SDVAR_SET_SMEM_ONCE((k3_filtered_kernel<NV, (NV <= 4 ? 128 : 256), 4>), dyn);                                         \
      k3_filtered_kernel<NV, (NV <= 4 ? 128 : 256), 4><<<grid, (NV <= 4 ? 128 : 256), dyn, st>>>(                          \
          logits_2BLV, B, L, in_ld, in_off, out_ld, out_off, seg, top_k, one_minus_top_p, noise, idx_out, mixed_out, prob_out); \
    } else if (filtered) {                                                                                                  \
      SDVAR_SET_SMEM_ONCE((k3_filtered_kernel<NV, 256, 4>), dyn);                                                           \
      k3_filtered_kernel<NV, 256, 4><<<grid, kThreads, dyn, st>>>(logits_2BLV, B, L, in_ld, in_off, out_ld, out_off, seg, top_k, \
                                                          one_minus_top_p, noise, idx_out, mixed_out, prob_out);   \
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
