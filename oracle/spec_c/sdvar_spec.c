/*
 * sdvar_spec.c -- bit-exact arithmetic SPEC of the two index-producing kernels
 * (TEST INFRASTRUCTURE ONLY; never linked into the product library).
 *
 *   K3  sdvar_spec_sample : CFG mix -> top-k -> top-p -> softmax -> exponential-race sample
 *        follows reference models/var.py:199-202 and models/helpers.py:6-19
 *        (torch.multinomial(n=1) == argmax(p / Exp(1)), SURVEY.md A5 / pin P4)
 *   K4  sdvar_spec_verify : target-vs-draft softmax, accept u*q[d] < p[d], residual resample,
 *        first-reject scan per (image, stage)   -- north_star item 3, SURVEY.md A7.
 *        PARITY UNPINNED: nothing in the reference computes this rule.
 *
 * The CUDA kernels in sdvar_b200/csrc/sampling.cu must reproduce these functions BIT FOR BIT
 * (same exp polynomial, same reduction order, IEEE mul/sub/div without contraction).  Where the
 * reference leaves the floating-point evaluation order to ATen (softmax sum, ascending cumsum)
 * this spec fixes one order; the deviation from ATen is a few ulp and changes a token only on
 * near-ties (probability ~1e-7 per token) -- the golden tests pin that on recorded vectors.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/spec_c/Makefile).
 *
 * Canonical reduction order over a row of V floats (V % 4 == 0), "256-lane order":
 *   lane t in [0,256) owns float4 chunks f = i*256 + t (i = 0,1,..), i.e. elements 4f..4f+3;
 *   lane partial = sequential sum over i then over the 4 components, starting from +0;
 *   the 32 lanes of a warp are combined with an xor butterfly (offsets 16,8,4,2,1);
 *   the 8 warp results are added left to right.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LANES 256

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* order-preserving map float -> uint32 (x<y  <=>  key(x)<key(y) for non-NaN, -0 < +0) */
static inline uint32_t fkey(float f) {
  uint32_t b = f2u(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

/* exp(x) for x <= 0 (softmax numerators): Cody-Waite reduction, degree-7 Taylor/Horner in fmaf, 2^n applied by an integer add
 * to the exponent field.  Values below 2^-126 are flushed: exp(x) = 0 for x < -87.33 (and for -inf), which keeps every result a
 * normal float, so the exponent add is exact and the kernels need neither a clamp nor denormal handling. */
#define SDVAR_EXP_MIN (-87.33f)
float sdvar_spec_expf(float x) {
  if (!(x >= SDVAR_EXP_MIN)) return 0.0f; /* also -inf and NaN */
  /* n = rint(x * log2 e) with ONE rounding: the product enters the 1.5 * 2^23 addition unrounded (a fused multiply-add), whose
   * result has ulp 1, so the addition itself rounds to the nearest integer (ties to even).  Stated as an fma on purpose: ptxas
   * contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2, so a two-rounding definition could not be kept on the GPU. */
  const float tm = fmaf(x, 1.44269504088896340736f, 12582912.0f);
  const float n = tm - 12582912.0f;
  float r = fmaf(n, -0.693145751953125f, x);
  r = fmaf(n, -1.42860682030941723212e-6f, r);
  float p = 1.0f / 5040.0f;
  p = fmaf(p, r, 1.0f / 720.0f);
  p = fmaf(p, r, 1.0f / 120.0f);
  p = fmaf(p, r, 1.0f / 24.0f);
  p = fmaf(p, r, 1.0f / 6.0f);
  p = fmaf(p, r, 0.5f);
  p = fmaf(p, r, 1.0f);
  p = fmaf(p, r, 1.0f);
  /* p in [0.70, 1.42], n in [-126, 0]: p * 2^n is normal, so adding n to the exponent field is the exact product.
   * bits(tm) = 0x4B400000 + n and 0x4B400000 << 23 == 0 (mod 2^32), hence bits(tm) << 23 == n << 23. */
  return u2f(f2u(p) + (f2u(tm) << 23));
}

/* canonical sum of term[v] * (pred ? 1 : 0); `keys`==NULL means no predicate */
static float canon_sum(const float* term, int V, const uint32_t* keys, uint32_t kmax_incl) {
  float lane[LANES];
  const int nvec = V / 4;
  for (int t = 0; t < LANES; ++t) {
    float acc = 0.0f;
    for (int f = t; f < nvec; f += LANES)
      for (int c = 0; c < 4; ++c) {
        const int v = 4 * f + c;
        const float a = (keys == NULL || keys[v] <= kmax_incl) ? term[v] : 0.0f;
        acc = acc + a;
      }
    lane[t] = acc;
  }
  float total = 0.0f;
  for (int w = 0; w < LANES / 32; ++w) {
    float* l = lane + 32 * w;
    for (int off = 16; off >= 1; off >>= 1) {
      float nl[32];
      for (int i = 0; i < 32; ++i) nl[i] = l[i] + l[i ^ off];
      memcpy(l, nl, sizeof(nl));
    }
    total = (w == 0) ? l[0] : total + l[0];
  }
  return total;
}

static int find_seg(const int* seg_begin, int S, int pos) {
  int j = 0;
  while (j + 1 < S && pos >= seg_begin[j + 1]) ++j;
  return j;
}

/* ---- K3 ------------------------------------------------------------------------------- */
/* Two arithmetic regimes, chosen by the (launch-uniform) filter setting:
 *   no filter  : softmax sum in the canonical 256-lane float order over the whole row (canon_sum);
 *   top-k/top-p: every sum is an INTEGER sum of fixed-point terms, hence independent of the summation order -- the kernel
 *                may histogram the row in any order (shared-memory atomics) and still agree bit for bit:
 *                  E_v    = (uint64) rint(e_v * 2^40)             e_v = sdvar_spec_expf(x_v - max) in [0,1]
 *                  Zi     = sum_v E_v  over the survivors of the top-k cut
 *                  thrE   = floor(Zi * floor(thr * 2^30) / 2^30)  the removable mass, in units of 2^-40 (exact integer arithmetic)
 *                  removed by top-p  <=>  sum{E_w : w alive, x_w <= x_v} <= thrE  and x_v is not the row maximum
 *                  Z2     = (float)(sum of the remaining E_v) * 2^-40
 *                (the reference accumulates float probabilities in ascending order, helpers.py:12-16; tie groups are kept whole.
 *                 Comparing fixed-point numerators with thrE is the same test as comparing probabilities with thr, without a
 *                 division per entry.)
 *                The exponential race argmax p_v / noise_v (helpers.py:19 via torch.multinomial) is run on the numerators,
 *                argmax e_v * R(noise_v), with R a division-free reciprocal (magic-constant seed + 4 Newton steps in fma; relative
 *                error < 2^-22): the common factor 1/Z2 cannot change the winner beyond rounding.  prob_out = e_win / Z2.
 * Ordering is by VALUE: -0 and +0 are the same logit (keys are taken of x + 0). */
#define FIX_E 1099511627776.0f /* 2^40 */
#define FIX_M 1073741824.0f    /* 2^30 */

/* reciprocal of the race: y0 = bits(0x7EF311C7 - bits(n)), then y <- y + y * (1 - n y), SDVAR_RCP_STEPS times, every step two fmaf */
#define SDVAR_RCP_STEPS 4
float sdvar_spec_rcp(float n) {
  float y = u2f(0x7EF311C7u - f2u(n));
  for (int i = 0; i < SDVAR_RCP_STEPS; ++i) {
    const float t = fmaf(-n, y, 1.0f);
    y = fmaf(y, t, y);
  }
  return y;
}

static int cmp_u32(const void* a, const void* b) {
  const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* one row; x is scratch of V floats that receives the mixed+masked logits */
static long long sample_row(const float* xc, const float* xu, float t1, float t2, int V, int top_k, float thr,
                            const float* noise, float* x, float* e, uint32_t* keys, float* prob_out) {
  for (int v = 0; v < V; ++v) x[v] = xc[v] * t1 - xu[v] * t2; /* two roundings of the products, then the difference */
  for (int v = 0; v < V; ++v) keys[v] = fkey(x[v] + 0.0f);
  const int filtered = (top_k > 0 && top_k < V) || thr >= 0.0f;
  if (top_k > 0 && top_k < V) {
    /* K = key of the k-th largest = largest K with #{key >= K} >= k (ties with the k-th value survive) */
    uint32_t K = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t tr = K | (1u << bit);
      int cnt = 0;
      for (int v = 0; v < V; ++v) cnt += keys[v] >= tr;
      if (cnt >= top_k) K = tr;
    }
    for (int v = 0; v < V; ++v)
      if (keys[v] < K) { x[v] = -INFINITY; keys[v] = fkey(-INFINITY); }
  }
  float m = -INFINITY;
  uint32_t kmax = 0;
  for (int v = 0; v < V; ++v) { if (x[v] > m) m = x[v]; if (keys[v] > kmax) kmax = keys[v]; }
  for (int v = 0; v < V; ++v) e[v] = sdvar_spec_expf(x[v] - m);
  if (!filtered) {
    if (noise == NULL) return -1;
    const float Z2 = canon_sum(e, V, NULL, 0);
    float best = -1.0f, bp = 0.0f;
    long long bi = 0;
    for (int v = 0; v < V; ++v) {
      const float p = e[v] / Z2;
      const float r = p / noise[v];
      if (r > best) { best = r; bi = v; bp = p; }
    }
    if (prob_out) *prob_out = bp;
    return bi;
  }
  static __thread uint64_t Ei[65536];
  uint64_t Zi = 0;
  for (int v = 0; v < V; ++v) { Ei[v] = (uint64_t)llrintf(e[v] * FIX_E); Zi += Ei[v]; }
  if (thr >= 0.0f) {
    const uint32_t thr_fix = (uint32_t)(thr * FIX_M); /* floor */
    const uint64_t thrE = (uint64_t)(((unsigned __int128)Zi * thr_fix) >> 30);
    static __thread uint32_t srt[65536];
    int n = 0;
    for (int v = 0; v < V; ++v)
      if (x[v] > -INFINITY) srt[n++] = keys[v];
    qsort(srt, (size_t)n, sizeof(uint32_t), cmp_u32);
    /* T = largest alive key with E{alive, key <= T} <= thrE (0 = nothing removable) */
    uint32_t T = 0;
    for (int i = 0; i < n; ++i) {
      if (i + 1 < n && srt[i + 1] == srt[i]) continue; /* evaluate a tie group at its last member */
      uint64_t acc = 0;
      for (int v = 0; v < V; ++v)
        if (x[v] > -INFINITY && keys[v] <= srt[i]) acc += Ei[v];
      if (acc <= thrE) T = srt[i]; else break;
    }
    for (int v = 0; v < V; ++v)
      if (x[v] > -INFINITY && keys[v] <= T && keys[v] != kmax) { x[v] = -INFINITY; e[v] = 0.0f; Ei[v] = 0; }
  }
  if (noise == NULL) return -1;
  uint64_t Z2i = 0;
  for (int v = 0; v < V; ++v) Z2i += Ei[v];
  const float Z2 = (float)Z2i * (1.0f / FIX_E);
  float best = -1.0f;
  long long bi = 0;
  float bp = 0.0f;
  for (int v = 0; v < V; ++v) {
    if (!(x[v] > -INFINITY)) continue; /* masked: not in the race (an underflowed survivor runs with r = 0) */
    const float r = e[v] * sdvar_spec_rcp(noise[v]);
    if (r > best) { best = r; bi = v; bp = e[v]; }
  }
  if (prob_out) *prob_out = best >= 0.0f ? bp / Z2 : 0.0f;
  return bi;
}

/* rows are (b,pos), b<B, pos<L; cond logits at row b*L+pos, uncond at (B+b)*L+pos of logits_2BLV;
 * seg_begin[S+1] partitions [0,L) into stages, stage j uses t1[j]=fl32(1+t), t2[j]=fl32(t);
 * thr = fl32(1-top_p) or <0 when top_p is disabled; noise/idx_out/mixed_out/prob_out may be NULL. */
int sdvar_spec_sample(const float* logits_2BLV, int B, int L, int V, const int* seg_begin, int S, const float* t1,
                      const float* t2, int top_k, float thr, const float* noise, long long* idx_out,
                      float* mixed_out, float* prob_out) {
  if (V % 4 != 0 || V > 65536 || S < 1) return -1;
  static __thread float x[65536], e[65536];
  static __thread uint32_t keys[65536];
  for (int b = 0; b < B; ++b)
    for (int pos = 0; pos < L; ++pos) {
      const long long row = (long long)b * L + pos;
      const int j = find_seg(seg_begin, S, pos);
      float pr = 0.0f;
      const long long id = sample_row(logits_2BLV + row * V, logits_2BLV + ((long long)(B + b) * L + pos) * V, t1[j],
                                      t2[j], V, top_k, thr, noise ? noise + row * V : NULL, x, e, keys, &pr);
      if (idx_out && noise) idx_out[row] = id;
      if (prob_out && noise) prob_out[row] = pr;
      if (mixed_out) memcpy(mixed_out + row * V, x, sizeof(float) * V);
    }
  return 0;
}

/* ---- K4 ------------------------------------------------------------------------------- */
int sdvar_spec_verify(const float* xt, const float* xd, const long long* draft_idx, const float* u, const float* noise,
                      int B, int L, int V, const int* seg_begin, int S, long long* out_idx, unsigned char* accept,
                      float* p_d_out, float* q_d_out, int* first_reject /*(B,S)*/, int* n_accept /*(B,S)*/,
                      int* accepted_stages /*(B)*/, int* summary /*[4]: min accepted stages, #accept, #reject, 0*/) {
  if (V % 4 != 0 || V > 65536 || S < 1) return -1;
  static __thread float et[65536], ed[65536], r[65536];
  int tot_acc = 0, tot_rej = 0, min_stages = S;
  for (int b = 0; b < B; ++b) {
    for (int j = 0; j < S; ++j) { first_reject[b * S + j] = seg_begin[j + 1] - seg_begin[j]; n_accept[b * S + j] = 0; }
    for (int pos = 0; pos < L; ++pos) {
      const long long row = (long long)b * L + pos;
      const int j = find_seg(seg_begin, S, pos);
      const float *t = xt + row * V, *d = xd + row * V;
      float mt = -INFINITY, md = -INFINITY;
      for (int v = 0; v < V; ++v) { if (t[v] > mt) mt = t[v]; if (d[v] > md) md = d[v]; }
      for (int v = 0; v < V; ++v) { et[v] = sdvar_spec_expf(t[v] - mt); ed[v] = sdvar_spec_expf(d[v] - md); }
      const float Zt = canon_sum(et, V, NULL, 0), Zd = canon_sum(ed, V, NULL, 0);
      const long long di = draft_idx[row];
      /* probabilities are exp * (1/Z): one IEEE division per row, one multiplication per element */
      const float iZt = 1.0f / Zt, iZd = 1.0f / Zd;
      const float pd = et[di] * iZt, qd = ed[di] * iZd;
      const int acc = (u[row] * qd) < pd;
      long long o = di;
      if (!acc) {
        /* residual r = max(0, p - q) with p - q = fma(e_t, 1/Z_t, -(e_d * 1/Z_d)); the exponential race argmax r / noise is
         * run as argmax r * (1 / noise): one IEEE reciprocal and one product per entry (no division chain), lowest index on
         * ties.  Identically zero residual: the race is run on p itself. */
        int anypos = 0;
        for (int v = 0; v < V; ++v) {
          float rv = fmaf(et[v], iZt, -(ed[v] * iZd));
          rv = rv > 0.0f ? rv : 0.0f;
          r[v] = rv;
          anypos |= rv > 0.0f;
        }
        float best = -1.0f;
        o = 0;
        for (int v = 0; v < V; ++v) {
          const float num = anypos ? r[v] : et[v] * iZt;
          const float q = num * (1.0f / noise[row * V + v]);
          if (q > best) { best = q; o = v; }
        }
      }
      out_idx[row] = o;
      accept[row] = (unsigned char)acc;
      if (p_d_out) p_d_out[row] = pd;
      if (q_d_out) q_d_out[row] = qd;
      if (acc) { n_accept[b * S + j]++; tot_acc++; }
      else {
        tot_rej++;
        const int within = pos - seg_begin[j];
        if (within < first_reject[b * S + j]) first_reject[b * S + j] = within;
      }
    }
    int a = 0;
    while (a < S && n_accept[b * S + a] == seg_begin[a + 1] - seg_begin[a]) ++a;
    accepted_stages[b] = a;
    if (a < min_stages) min_stages = a;
  }
  summary[0] = min_stages; summary[1] = tot_acc; summary[2] = tot_rej; summary[3] = 0;
  return 0;
}

/* nearest codebook entry (reference models/quant.py:155-157), the arithmetic of sdvar_vq_nearest_code:
 * d = fma(-2, <z,e>, |z|^2 + |e|^2), every dot product a sequential fma chain over c, argmin with the lowest index on ties. */
int sdvar_spec_nearest_code(const float* z, const float* E, long long N, int C, int V, long long* idx_out) {
  for (long long r = 0; r < N; ++r) {
    const float* zr = z + r * C;
    float zsq = 0.0f;
    for (int c = 0; c < C; ++c) zsq = fmaf(zr[c], zr[c], zsq);
    float best = INFINITY;
    long long bi = 0;
    for (int v = 0; v < V; ++v) {
      const float* e = E + (long long)v * C;
      float esq = 0.0f, dot = 0.0f;
      for (int c = 0; c < C; ++c) esq = fmaf(e[c], e[c], esq);
      for (int c = 0; c < C; ++c) dot = fmaf(zr[c], e[c], dot);
      const float d = fmaf(-2.0f, dot, zsq + esq);
      if (d < best) { best = d; bi = v; }
    }
    idx_out[r] = bi;
  }
  return 0;
}
