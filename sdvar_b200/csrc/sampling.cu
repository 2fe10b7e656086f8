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
#include "common.cuh"

namespace sdvar {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// exp(x) for x <= 0, bit-identical to sdvar_spec_expf.  Branch-free so the 16-32 independent exponentials of a thread
// interleave: the input is clamped at -104 (the spec returns 0 below it) and the power-of-two scaling is always done in two
// exact-or-once-rounded steps (2^(n+100) then 2^-100), which rounds identically to a single multiplication by 2^n.
__device__ __forceinline__ float spec_expf(float x) {
  const float xc = fmaxf(x, -104.0f);
  // n = rint(xc * log2 e) with one rounding and without the quarter-rate FRND / F2I conversions: the exact product is added to
  // 1.5*2^23 inside one fma (the spec is stated that way), the sum lands in the binade with ulp 1, so the fma itself rounds to
  // nearest-even; the integer is the mantissa difference
  const float tm = __fmaf_rn(xc, 1.44269504088896340736f, 12582912.0f);
  const float n = __fsub_rn(tm, 12582912.0f);
  float r = __fmaf_rn(n, -0.693145751953125f, xc);
  r = __fmaf_rn(n, -1.42860682030941723212e-6f, r);
  float p = 1.0f / 5040.0f;
  p = __fmaf_rn(p, r, 1.0f / 720.0f);
  p = __fmaf_rn(p, r, 1.0f / 120.0f);
  p = __fmaf_rn(p, r, 1.0f / 24.0f);
  p = __fmaf_rn(p, r, 1.0f / 6.0f);
  p = __fmaf_rn(p, r, 0.5f);
  p = __fmaf_rn(p, r, 1.0f);
  p = __fmaf_rn(p, r, 1.0f);
  const int ni = __float_as_int(tm) - 0x4B400000;   // in [-151, 0]
  // no select for x < -104: the clamped value scales to p * 2^-150 < 2^-150 * 1 = half of the smallest subnormal and the second
  // (once-rounded) multiplication returns exactly 0, which is what the spec returns below -104
  return __fmul_rn(__fmul_rn(p, u2f((uint32_t)(ni + 100 + 127) << 23)), u2f((uint32_t)(-100 + 127) << 23));
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
  const f32x2 xc = pk2(fmaxf(x0, -104.0f), fmaxf(x1, -104.0f));
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
  float tm0, tm1;
  unpk2(tm, tm0, tm1);
  const uint32_t s0 = (uint32_t)(__float_as_int(tm0) - 0x4B400000 + 100 + 127) << 23;
  const uint32_t s1 = (uint32_t)(__float_as_int(tm1) - 0x4B400000 + 100 + 127) << 23;
  return mul2(mul2(p, pk2(u2f(s0), u2f(s1))), splat2(u2f((uint32_t)(-100 + 127) << 23)));
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
// Register diet: a thread keeps only the order-preserving KEYS of its 4*NV mixed logits (the logit is the exact inverse of
// its key) plus, during the top-p search, their probabilities; exponentials are recomputed (bit-identically) for the final
// draw.  ~60 registers -> four 256-thread CTAs per SM, which is what hides the serial latency of the canonical reductions.
template <int NV>
__global__ void __launch_bounds__(kThreads, 4)
k3_sample_kernel(const float* __restrict__ logits, int B, int L, int in_ld, int in_off, SegTable seg, int top_k, float thr,
                 const float* __restrict__ noise, long long* __restrict__ idx_out, float* __restrict__ mixed_out,
                 float* __restrict__ prob_out) {
  constexpr int V = NV * 1024;
  constexpr int E = NV * 4;
  __shared__ RedSmem sm;
  int slot = 0;
  const int tid = threadIdx.x;
  const long long rows = (long long)B * L;
  const uint32_t kneg = fkey(-INFINITY);
  for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / L), pos = (int)(row - (long long)b * L);
    const int j = seg_of(seg, pos);
    const float t1 = seg.t1[j], t2 = seg.t2[j];
    const float4* pc = reinterpret_cast<const float4*>(logits + ((long long)b * in_ld + in_off + pos) * V);
    const float4* pu = reinterpret_cast<const float4*>(logits + ((long long)(B + b) * in_ld + in_off + pos) * V);
    uint32_t key[E];
    {
      float4 a[NV], c[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) { a[i] = ldg_stream(pc + i * kThreads + tid); c[i] = ldg_stream(pu + i * kThreads + tid); }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        key[4 * i + 0] = fkey(__fsub_rn(__fmul_rn(a[i].x, t1), __fmul_rn(c[i].x, t2)));
        key[4 * i + 1] = fkey(__fsub_rn(__fmul_rn(a[i].y, t1), __fmul_rn(c[i].y, t2)));
        key[4 * i + 2] = fkey(__fsub_rn(__fmul_rn(a[i].z, t1), __fmul_rn(c[i].z, t2)));
        key[4 * i + 3] = fkey(__fsub_rn(__fmul_rn(a[i].w, t1), __fmul_rn(c[i].w, t2)));
      }
    }

    // ---- row max / min (exact, on keys) ----
    uint32_t kmax = 0, kminc = 0;   // kminc = ~min
#pragma unroll
    for (int e = 0; e < E; ++e) { kmax = max(kmax, key[e]); kminc = max(kminc, ~key[e]); }
    block_max_u32x2(kmax, kminc, sm, slot);
    const float m = fkey_inv(kmax);

    // ---- top-k: K = key of the k-th largest.  Midpoint bisection on the key VALUE range [min, max] with early exit:
    // cL = #{key >= lo} (>= k) and cH = #{key >= hi} (< k) bracket the answer; once the bracket [lo,hi) holds exactly one key
    // that key IS the k-th largest (one min-reduction fetches it); ~log2(V)+2 rounds instead of 32.  The result is the
    // unique largest K with #{key >= K} >= k, i.e. exactly what the bit-serial search of the spec returns.
    uint32_t klow = ~kminc;   // smallest key that can still carry probability mass
    int alive = V;            // #{key >= klow}
    if (top_k > 0 && top_k < V) {
      uint32_t lo = ~kminc, hi = kmax + 1u;
      int cL = V, cH = 0;
#pragma unroll 1
      while (cL - cH > 1 && hi - lo > 1u) {
        // midpoint of the bracket in VALUE space when it is usable (logits straddle 0, and in key space the floats around 0
        // occupy half of the range: a key-space midpoint spends ~7 rounds walking up the exponents), else in key space
        uint32_t mid = lo + ((hi - lo) >> 1);
        const float flo = fkey_inv(lo), fhi = fkey_inv(hi - 1u);
        if (flo > -INFINITY) {
          const uint32_t mv = fkey(__fadd_rn(__fmul_rn(0.5f, flo), __fmul_rn(0.5f, fhi)));
          if (mv > lo && mv < hi) mid = mv;
        }
        int c = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) c += (key[e] >= mid) ? 1 : 0;
        c = block_sum_int(c, sm, slot);
        if (c >= top_k) { lo = mid; cL = c; } else { hi = mid; cH = c; }
      }
      uint32_t K = lo;
      if (cL - cH == 1 && hi - lo > 1u) {   // the smallest key that is still >= lo (min via max of the complement)
        uint32_t mn = 0, z = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) mn = max(mn, (key[e] >= lo) ? ~key[e] : 0u);
        block_max_u32x2(mn, z, sm, slot);
        K = ~mn;
      }
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (key[e] < K) key[e] = kneg;
      klow = K;
      alive = cL;             // #{key >= K}: the bracket [lo, hi) held no key below K
    }

    // ---- top-p: remove v iff mass{key <= key_v} <= thr, never the max ----
    if (thr >= 0.0f) {
      float p[E];
      float z = 0.0f;
#pragma unroll
      for (int e = 0; e < E; e += 2) {
        unpk2(spec_expf2(__fsub_rn(fkey_inv(key[e]), m), __fsub_rn(fkey_inv(key[e + 1]), m)), p[e], p[e + 1]);
        z = __fadd_rn(__fadd_rn(z, p[e]), p[e + 1]);
      }
      const float Z = block_sum(z, sm, slot);
      float tot = 0.0f;
#pragma unroll
      for (int e = 0; e < E; ++e) { p[e] = fdiv_nz(p[e], Z); tot = __fadd_rn(tot, p[e]); }
      tot = block_sum(tot, sm, slot);      // canonical mass of the whole row = mass{key <= kmax}
      // Midpoint bisection on the canonical masked mass with early exit: lo is good (mass{key<=lo} <= thr, nL = #{key<=lo}),
      // hi is bad.  The mass only changes at key values, so once at most one key separates lo from hi no threshold in
      // between can change the removed set {key <= lo}: the answer equals the spec's largest good threshold.
      // start from the largest threshold that is certainly good: below the smallest surviving key the masked mass is an exact 0.
      // (Starting from 0 spent ~11 of ~19 rounds bisecting the empty key range under the top-k cut.)
      uint32_t lo = klow > 0u ? klow - 1u : 0u, hi = kmax;
      int nL = V - alive, nH = V;
      if (tot <= thr) lo = kmax;           // everything is removable (the max itself is always kept)
#pragma unroll 1
      while (lo != kmax && nH - nL > 1 && hi - lo > 1u) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        float a = 0.0f;
        int c = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const bool in = key[e] <= mid;
          a = __fadd_rn(a, in ? p[e] : 0.0f);
          c += in ? 1 : 0;
        }
        block_sum_fi(a, c, sm, slot);
        if (a <= thr) { lo = mid; nL = c; } else { hi = mid; nH = c; }
      }
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (key[e] <= lo && key[e] != kmax) key[e] = kneg;
    }

    if (mixed_out != nullptr) {
      float4* po = reinterpret_cast<float4*>(mixed_out + row * V);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        stg_stream(po + i * kThreads + tid, make_float4(fkey_inv(key[4 * i]), fkey_inv(key[4 * i + 1]), fkey_inv(key[4 * i + 2]), fkey_inv(key[4 * i + 3])));
    }

    if (noise != nullptr) {
      const float4* pn = reinterpret_cast<const float4*>(noise + row * V);
      float4 nz[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) nz[i] = ldg_stream(pn + i * kThreads + tid);
      float ex[E];
      float z = 0.0f;
#pragma unroll
      for (int e = 0; e < E; e += 2) {
        unpk2(spec_expf2(__fsub_rn(fkey_inv(key[e]), m), __fsub_rn(fkey_inv(key[e + 1]), m)), ex[e], ex[e + 1]);
        z = __fadd_rn(__fadd_rn(z, ex[e]), ex[e + 1]);
      }
      const float Z2 = block_sum(z, sm, slot);
      float best = -1.0f, bestp = 0.0f;
      int bi = 0x7FFFFFFF;
      const bool filtered = (top_k > 0 && top_k < V) || thr >= 0.0f;   // block-uniform: masked entries (exact zeros) exist
      if (filtered) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float nn[4] = {nz[i].x, nz[i].y, nz[i].z, nz[i].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float pv = fdiv_nz(ex[4 * i + c], Z2);
            const float r = fdiv_nz(pv, nn[c]);
            if (r > best) { best = r; bi = 4 * (i * kThreads + tid) + c; bestp = pv; }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float nn[4] = {nz[i].x, nz[i].y, nz[i].z, nz[i].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float pv = __fdiv_rn(ex[4 * i + c], Z2);
            const float r = __fdiv_rn(pv, nn[c]);
            if (r > best) { best = r; bi = 4 * (i * kThreads + tid) + c; bestp = pv; }
          }
        }
      }
      const unsigned long long w = block_max_u64(pack_best(best, bi), sm, slot);
      int win = (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
      if (win == 0x7FFFFFFF) win = 0;
      if (bi == win || (tid == 0 && (uint32_t)(w >> 32) == fkey(-1.0f))) {
        if (idx_out) idx_out[row] = win;
        if (prob_out) prob_out[row] = (bi == win) ? bestp : 0.0f;
      }
    }
  }
}

// ================================================================================================
// K4
// ================================================================================================
__global__ void k4_init_kernel(int* first_reject, int* n_accept, int B, SegTable seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * seg.S) {
    const int j = i % seg.S;
    first_reject[i] = seg.begin[j + 1] - seg.begin[j];
    n_accept[i] = 0;
  }
}

// K4 v2: the two logit rows of a token (2 x V fp32 = 32 KiB) are staged in shared memory by 1-D bulk TMA copies
// (cp.async.bulk + mbarrier), double-buffered so row i+1 streams in while row i is processed; several CTAs per SM keep
// >= 96 KiB in flight per SM.  The accept path touches each value once (max, exp, sum: nothing is kept in registers);
// only a rejected row re-reads its staged logits to rebuild p and q for the residual resample.  Same arithmetic and
// reduction order as v1, i.e. bit-exact to oracle/spec_c/sdvar_spec.c.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init1(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_par(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "K4_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra K4_DONE;\n"
      "bra K4_WAIT;\n"
      "K4_DONE:\n"
      "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

template <int NV>
__global__ void __launch_bounds__(kThreads, 3)
k4_verify_kernel(const float* __restrict__ xt, const float* __restrict__ xd, const long long* __restrict__ draft_idx,
                 const float* __restrict__ u, const float* __restrict__ noise, int B, int L, SegTable seg,
                 long long* __restrict__ out_idx, unsigned char* __restrict__ accept, float* __restrict__ p_d_out,
                 float* __restrict__ q_d_out, int* first_reject, int* n_accept, int* accepted_stages, int* summary,
                 int* counter) {
  constexpr int V = NV * 1024;
  extern __shared__ __align__(128) unsigned char k4_smem[];
  float* stage_buf = reinterpret_cast<float*>(k4_smem);                    // [2][2][V]
  uint64_t* full = reinterpret_cast<uint64_t*>(k4_smem + 2 * 2 * V * 4);   // [2]
  __shared__ RedSmem sm;
  __shared__ int s_last;
  int slot = 0;
  const int tid = threadIdx.x;
  const long long rows = (long long)B * L;
  if (tid == 0) {
    mbar_init1(&full[0]);
    mbar_init1(&full[1]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((long long)blockIdx.x < rows) {
      mbar_expect(&full[0], 2 * V * 4);
      bulk_g2s(stage_buf, xt + (long long)blockIdx.x * V, V * 4, &full[0]);
      bulk_g2s(stage_buf + V, xd + (long long)blockIdx.x * V, V * 4, &full[0]);
    }
  }
  __syncthreads();
  uint32_t k = 0;
  for (long long row = blockIdx.x; row < rows; row += gridDim.x, ++k) {
    const uint32_t st = k & 1;
    // prefetch the next row into the other stage (all reads of that stage finished before the barrier closing iteration k-1)
    const long long nrow = row + gridDim.x;
    if (tid == 0 && nrow < rows) {
      mbar_expect(&full[st ^ 1], 2 * V * 4);
      bulk_g2s(stage_buf + (st ^ 1) * 2 * V, xt + nrow * V, V * 4, &full[st ^ 1]);
      bulk_g2s(stage_buf + (st ^ 1) * 2 * V + V, xd + nrow * V, V * 4, &full[st ^ 1]);
    }
    const int d = (int)draft_idx[row];
    mbar_wait_par(&full[st], (k >> 1) & 1);
    float4* st4 = reinterpret_cast<float4*>(stage_buf + st * 2 * V);
    float4* sd4 = st4 + V / 4;
    // pass 1: row maxima (order-independent)
    float mt = -INFINITY, md = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 a = st4[i * kThreads + tid], c = sd4[i * kThreads + tid];
      mt = fmaxf(fmaxf(mt, fmaxf(a.x, a.y)), fmaxf(a.z, a.w));
      md = fmaxf(fmaxf(md, fmaxf(c.x, c.y)), fmaxf(c.z, c.w));
    }
    uint32_t kt = fkey(mt), kd = fkey(md);
    block_max_u32x2(kt, kd, sm, slot);
    mt = fkey_inv(kt);
    md = fkey_inv(kd);
    // pass 2: canonical exp sums; the owner of element d keeps its two exponentials
    // two adjacent elements of the SAME row per packed instruction: the pairs are the register pairs LDS.128 delivers and
    // STS.128 takes back, so no moves are spent on re-pairing; the row sums stay scalar, in element order (the canonical order)
    float zt = 0.0f, zd = 0.0f;
    const f32x2 nmt = pk2(-mt, -mt), nmd = pk2(-md, -md);    // a - m == a + (-m) exactly
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 a = st4[i * kThreads + tid], c = sd4[i * kThreads + tid];
      float x0, x1;
      unpk2(add2(pk2(a.x, a.y), nmt), x0, x1);
      unpk2(spec_expf2(x0, x1), a.x, a.y);
      unpk2(add2(pk2(a.z, a.w), nmt), x0, x1);
      unpk2(spec_expf2(x0, x1), a.z, a.w);
      unpk2(add2(pk2(c.x, c.y), nmd), x0, x1);
      unpk2(spec_expf2(x0, x1), c.x, c.y);
      unpk2(add2(pk2(c.z, c.w), nmd), x0, x1);
      unpk2(spec_expf2(x0, x1), c.z, c.w);
      zt = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(zt, a.x), a.y), a.z), a.w);
      zd = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(zd, c.x), c.y), c.z), c.w);
      // the exponentials replace the staged logits (each thread rewrites only the chunks it owns), so neither the accept
      // test nor a residual resample has to exponentiate again
      st4[i * kThreads + tid] = a;
      sd4[i * kThreads + tid] = c;
    }
    block_sum2(zt, zd, sm, slot);  // zt, zd now hold Zt, Zd
    // the owner of element d evaluates the accept test and publishes the per-token outputs itself; everything only it needs
    // (1/Z, image / position / stage of the row, u) is computed there and not by the other 255 threads
    const bool owner = ((d >> 2) & (kThreads - 1)) == tid;
    int rej = 0;
    if (owner) {
      const float izt = __fdiv_rn(1.0f, zt), izd = __fdiv_rn(1.0f, zd);
      const int b = (int)(row / L), pos = (int)(row - (long long)b * L);
      const int j = seg_of(seg, pos);
      const float uu = u[row];
      const float* rt = stage_buf + st * 2 * V;
      const float pdv = __fmul_rn(rt[d], izt), qdv = __fmul_rn(rt[V + d], izd);
      rej = (__fmul_rn(uu, qdv) < pdv) ? 0 : 1;
      accept[row] = (unsigned char)(rej ^ 1);
      if (p_d_out) p_d_out[row] = pdv;
      if (q_d_out) q_d_out[row] = qdv;
      if (!rej) {
        out_idx[row] = d;
        atomicAdd(&n_accept[b * seg.S + j], 1);
      } else {
        atomicMin(&first_reject[b * seg.S + j], pos - seg.begin[j]);
      }
    }
    const int acc = __syncthreads_or(rej) ? 0 : 1;
    int out = d;
    if (!acc) {  // block-uniform: residual resample from the staged logits
      const float izt = __fdiv_rn(1.0f, zt), izd = __fdiv_rn(1.0f, zd);
      const float4* pn = reinterpret_cast<const float4*>(noise + row * V);
      float4 nz[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) nz[i] = ldg_stream(pn + i * kThreads + tid);
      float pv[NV * 4], rv[NV * 4];
      int anyp = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 a = st4[i * kThreads + tid], c = sd4[i * kThreads + tid];
        const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float p1 = __fmul_rn(av[q], izt);
          float r1 = __fsub_rn(p1, __fmul_rn(cv[q], izd));
          r1 = r1 > 0.0f ? r1 : 0.0f;
          anyp |= (r1 > 0.0f) ? 1 : 0;
          pv[4 * i + q] = p1;
          rv[4 * i + q] = r1;
        }
      }
      anyp = __syncthreads_or(anyp);
      float best = -1.0f;
      int bi = 0x7FFFFFFF;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float nn[4] = {nz[i].x, nz[i].y, nz[i].z, nz[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float r = fdiv_nz(anyp ? rv[4 * i + q] : pv[4 * i + q], nn[q]);
          if (r > best) { best = r; bi = 4 * (i * kThreads + tid) + q; }
        }
      }
      const unsigned long long w = block_max_u64(pack_best(best, bi), sm, slot);
      out = (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
      if (out == 0x7FFFFFFF) out = 0;
      if (tid == 0) out_idx[row] = out;
    }
    // (no trailing barrier: the last shared-memory reads of stage st are always followed by a block reduction barrier
    //  before the next iteration's prefetch can target that stage)
  }
  // ---- last CTA finalises the per-image / batch scan (integer atomics => deterministic) ----
  __threadfence();
  if (tid == 0) s_last = (atomicAdd(counter, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    int mn = seg.S, na = 0;
    for (int b = tid; b < B; b += kThreads) {
      int a = 0;
      bool open = true;
      for (int j2 = 0; j2 < seg.S; ++j2) {
        const int n = __ldcg(&n_accept[b * seg.S + j2]);
        na += n;
        if (open && n == seg.begin[j2 + 1] - seg.begin[j2]) ++a; else open = false;
      }
      accepted_stages[b] = a;
      mn = min(mn, a);
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
      *counter = 0;
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

extern "C" int sdvar_sample_cfg_topk_topp(const float* logits_2BLV, int B, int L, int in_ld, int in_off, int V,
                                          const int* seg_begin_host, int S,
                                          const float* t1_host, const float* t2_host, int top_k, float one_minus_top_p,
                                          const float* noise, long long* idx_out, float* mixed_out, float* prob_out,
                                          void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(logits_2BLV && B > 0 && L > 0, "bad logits/B/L");
  SDVAR_REQUIRE(in_off >= 0 && in_ld >= in_off + L, "bad row mapping in_ld=%d in_off=%d L=%d", in_ld, in_off, L);
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
#define SDVAR_K3(NV)                                                                                          \
  case NV:                                                                                                    \
    k3_sample_kernel<NV><<<grid, kThreads, 0, st>>>(logits_2BLV, B, L, in_ld, in_off, seg, top_k, one_minus_top_p, noise,   \
                                                    idx_out, mixed_out, prob_out);                           \
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

extern "C" int sdvar_verify_accept_resample(const float* xt, const float* xd, const long long* draft_idx, const float* u,
                                            const float* noise, int B, int L, int V, const int* seg_begin_host, int S,
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
  k4_init_kernel<<<(B * S + 255) / 256, 256, 0, st>>>(first_reject, n_accept, B, seg);
  SDVAR_LAUNCH_CHECK();
  const size_t k4_smem = (size_t)2 * 2 * V * 4 + 64;          // two stages x (target row + draft row) + mbarriers
  const int per_sm = (int)((220 * 1024) / (k4_smem + 1024)) < 3 ? (int)((220 * 1024) / (k4_smem + 1024)) : 3;
  SDVAR_REQUIRE(per_sm >= 1, "V=%d rows do not fit the shared-memory ring", V);
  const int grid = row_grid((long long)B * L, per_sm);
#define SDVAR_K4(NV)                                                                                               \
  case NV: {                                                                                                       \
    static bool attr = false;                                                                                      \
    if (!attr) {                                                                                                   \
      SDVAR_CUDA(cudaFuncSetAttribute(k4_verify_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4_smem)); \
      attr = true;                                                                                                 \
    }                                                                                                              \
    k4_verify_kernel<NV><<<grid, kThreads, k4_smem, st>>>(xt, xd, draft_idx, u, noise, B, L, seg, out_idx, accept, \
                                                          p_d_out, q_d_out, first_reject, n_accept, accepted_stages, \
                                                          summary, workspace);                                     \
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
