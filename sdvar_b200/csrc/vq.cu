// vq.cu -- K5: VectorQuantizer2 next-input step and the stage input maps.
//
// Replaces models/var.py:205-211 + models/quant.py:187-196 (get_next_autoregressive_input),
// :199-206 (Phi = (1-r)*h + r*conv3x3(h), r = |quant_resi|), and models/var.py:179-188 (stage input maps).
//
// f_hat is tiny (32 KiB per image at 256 px), so this step is latency-bound, not HBM-bound: the design
// goal is few launches and on-chip staging.  Kernel A gives each CTA four output rows of one image: it
// gathers only the codebook rows the bicubic taps of rows y0-1..y0+4 touch, interpolates separably in
// shared memory, runs the 3x3 Phi conv with the weights staged in shared memory ([ci][tap][co], bank =
// co) and adds into f_hat through a transposing shared tile so the global update is row-contiguous.
// Kernel B is the overlapping-window area pooling (adaptive average) to the next stage's resolution.
#include "common.cuh"

namespace sdvar {

constexpr int kC = 32;       // Cvae
constexpr int kMaxHW = 32;   // 512 px pyramid

// Keys cubic convolution coefficients, A = -0.75 (ATen upsample_bicubic2d, align_corners=False)
__device__ __forceinline__ void cubic_coeffs(float t, float w[4]) {
  const float A = -0.75f;
  float x = t + 1.0f;
  w[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
  x = t;
  w[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 1.0f - t;
  w[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 2.0f - t;
  w[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}
// source index / fraction of output coordinate o (half-pixel centres, no clamping of the real coordinate)
__device__ __forceinline__ void cubic_src(int o, int pn, int HW, int& ix, float& t) {
  const float scale = (float)pn / (float)HW;
  const float s = scale * ((float)o + 0.5f) - 0.5f;
  const float fl = floorf(s);
  ix = (int)fl;
  t = s - fl;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// One CTA computes R consecutive output rows of one image, so the 36 KiB of Phi weights are staged once per R rows (round 1:
// one row per CTA, 7 waves of CTAs that each re-staged the weights with 32-way bank-conflicting transposing stores; ~130 us per
// launch at B=64).  HWMAX sizes the staging buffers for the pyramid in use (16: 256 px, 32: 512 px).
template <int HWMAX, int R>
struct VqSmemT {
  static constexpr int kSrcRows = R + 5;         // output rows y0-1 .. y0+R need at most R+2+3 source rows when upsampling
  float w[kC * 9 * (kC + 1)];                    // [ci][tap][co], rows padded to 33 floats: the transposing stores are conflict-free
  float bias[kC];
  float src[kSrcRows][HWMAX][kC];                // gathered codebook rows  [r][q][c]
  float tmp[kSrcRows][HWMAX][kC];                // after horizontal interpolation [r][x][c]
  float hup[R + 2][HWMAX + 2][kC];               // rows y0-1 .. y0+R with one zero column on each side [dy][x+1][c]
  float outT[R][kC][HWMAX + 1];                  // conv result transposed [row][co][x]
};

template <int HWMAX, int R>
__global__ void __launch_bounds__(256)
vq_accumulate_kernel(const long long* __restrict__ idx, int pn, int HW, const float* __restrict__ codebook,
                     const float* __restrict__ phi_w, const float* __restrict__ phi_b, float resi, float* __restrict__ f_hat,
                     float* __restrict__ f_rest) {
  using Smem = VqSmemT<HWMAX, R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int b = blockIdx.y, y0 = blockIdx.x * R, tid = threadIdx.x;
  const int lane = tid & 31, grp = tid >> 5;
  const bool up = (pn != HW);
  const int nrows = min(R, HW - y0);

  // stage Phi weights as [ci][tap][co] (global layout is [co][ci][3][3])
  for (int i = tid; i < kC * kC * 9; i += 256) {
    const int co = i / (kC * 9), rem = i - co * kC * 9, ci = rem / 9, tap = rem - ci * 9;
    s.w[(ci * 9 + tap) * (kC + 1) + co] = phi_w[i];
  }
  if (tid < kC) s.bias[tid] = phi_b[tid];

  // source rows needed by output rows y0-1 .. y0+nrows
  int r0, r1;
  if (up) {
    int ixa, ixb; float t;
    cubic_src(max(y0 - 1, 0), pn, HW, ixa, t);
    cubic_src(min(y0 + nrows, HW - 1), pn, HW, ixb, t);
    r0 = clampi(ixa - 1, 0, pn - 1);
    r1 = clampi(ixb + 2, 0, pn - 1);
  } else {
    r0 = max(y0 - 1, 0);
    r1 = min(y0 + nrows, HW - 1);
  }
  const int nr = r1 - r0 + 1;  // <= kSrcRows
  // gather: one warp per token, lane = channel (128-byte coalesced codebook rows)
  for (int tk = grp; tk < nr * pn; tk += 8) {
    const int r = tk / pn, q = tk - r * pn;
    const long long id = idx[(long long)b * pn * pn + (r0 + r) * pn + q];
    s.src[r][q][lane] = codebook[id * kC + lane];
  }
  __syncthreads();
  if (up) {
    // horizontal: tmp[r][x][c]
    for (int i = grp; i < nr * HW; i += 8) {
      const int r = i / HW, x = i - r * HW;
      int ix; float t, w[4];
      cubic_src(x, pn, HW, ix, t);
      cubic_coeffs(t, w);
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += w[k] * s.src[r][clampi(ix - 1 + k, 0, pn - 1)][lane];
      s.tmp[r][x][lane] = acc;
    }
    __syncthreads();
  }
  // vertical + zero padding: hup[dy][x+1][c], dy = 0 .. nrows+1  <->  output row y0-1+dy
  for (int i = grp; i < (nrows + 2) * (HW + 2); i += 8) {
    const int dy = i / (HW + 2), xp = i - dy * (HW + 2);
    const int yy = y0 - 1 + dy, x = xp - 1;
    float v = 0.0f;
    if (yy >= 0 && yy < HW && x >= 0 && x < HW) {
      if (up) {
        int iy; float t, w[4];
        cubic_src(yy, pn, HW, iy, t);
        cubic_coeffs(t, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) v += w[k] * s.tmp[clampi(iy - 1 + k, 0, pn - 1) - r0][x][lane];
      } else {
        v = s.src[yy - r0][x][lane];
      }
    }
    s.hup[dy][xp][lane] = v;
  }
  __syncthreads();
  // conv: thread (co = lane, pixel = grp, grp+8, ..) over the R x HW pixels of this CTA
  for (int px = grp; px < nrows * HW; px += 8) {
    const int ry = px / HW, x = px - ry * HW;
    float acc = s.bias[lane];
#pragma unroll 4
    for (int ci = 0; ci < kC; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
          acc += s.w[(ci * 9 + dy * 3 + dx) * (kC + 1) + lane] * s.hup[ry + dy][x + dx][ci];
    }
    // Phi (models/quant.py:205-206): (1-r)*h + r*conv(h); r = 0.5 in every released checkpoint
    s.outT[ry][lane][x] = (1.0f - resi) * s.hup[ry + 1][x + 1][lane] + resi * acc;
  }
  __syncthreads();
  // f_hat[b, c, y0+ry, :] += outT[ry][c][:]   (row-contiguous)
  for (int i = tid; i < nrows * kC * HW; i += 256) {
    const int ry = i / (kC * HW), rem = i - ry * kC * HW, c = rem / HW, x = rem - c * HW;
    const long long o = (((long long)b * kC + c) * HW + y0 + ry) * HW + x;
    f_hat[o] += s.outT[ry][c][x];
    if (f_rest != nullptr) f_rest[o] -= s.outT[ry][c][x];   // encode side: the reference's running residual (models/quant.py:163)
  }
}

// adaptive average pooling (F.interpolate(mode='area')): window [floor(o*HW/pn2), ceil((o+1)*HW/pn2))
__global__ void __launch_bounds__(256)
vq_area_down_kernel(const float* __restrict__ f_hat, int HW, int pn2, float* __restrict__ next_map) {
  const int b = blockIdx.x;
  const int n = kC * pn2 * pn2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / (pn2 * pn2), rem = i - c * pn2 * pn2, oy = rem / pn2, ox = rem - oy * pn2;
    const int y0 = (oy * HW) / pn2, y1 = ((oy + 1) * HW + pn2 - 1) / pn2;
    const int x0 = (ox * HW) / pn2, x1 = ((ox + 1) * HW + pn2 - 1) / pn2;
    const float* src = f_hat + ((long long)b * kC + c) * HW * HW;
    float acc = 0.0f;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) acc += src[yy * HW + xx];
    next_map[(long long)b * n + i] = acc / (float)((y1 - y0) * (x1 - x0));
  }
}

// x[r,t,:] = W @ nm[b,:,t] + bias + lvl_pos[t,:], r in {b, B+b}   (models/var.py:185-188)
constexpr int kEmbTok = 32;   // tokens per CTA: W_we (C x 32 fp32) is re-read from L2 once per CTA, so few large CTAs beat many small ones
__global__ void __launch_bounds__(256)
embed_next_map_kernel(const float* __restrict__ nm, int B, int l, int C, const float* __restrict__ W,
                      const float* __restrict__ bias, const float* __restrict__ lvl_pos, float* __restrict__ x,
                      int ldx_tokens, int tok_off) {
  __shared__ float tok[kEmbTok][kC];
  const int b = blockIdx.y, t0 = blockIdx.x * kEmbTok;
  const int nt = min(kEmbTok, l - t0);
  for (int i = threadIdx.x; i < kEmbTok * kC; i += blockDim.x) {
    const int tt = i / kC, c = i - tt * kC;
    tok[tt][c] = (tt < nt) ? nm[((long long)b * kC + c) * l + t0 + tt] : 0.0f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float w[kC];
    const float4* wr = reinterpret_cast<const float4*>(W + (long long)c * kC);
#pragma unroll
    for (int k = 0; k < kC / 4; ++k) {
      const float4 v = __ldg(wr + k);
      w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
    }
    const float bc = bias[c];
    for (int tt = 0; tt < nt; ++tt) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < kC; ++k) acc += w[k] * tok[tt][k];
      const float v = acc + bc + lvl_pos[(long long)(t0 + tt) * C + c];
      x[((long long)b * ldx_tokens + tok_off + t0 + tt) * C + c] = v;
      x[((long long)(B + b) * ldx_tokens + tok_off + t0 + tt) * C + c] = v;
    }
  }
}

__global__ void first_map_kernel(const float* __restrict__ cond, int first_l, int C, const float* __restrict__ pos_start,
                                 const float* __restrict__ lvl_pos, float* __restrict__ x, int ldx_tokens, int tok_off) {
  const int r = blockIdx.y, t = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    x[((long long)r * ldx_tokens + tok_off + t) * C + c] = cond[(long long)r * C + c] + pos_start[(long long)t * C + c] + lvl_pos[(long long)t * C + c];
}

// ---- encode side: nearest codebook entry, reference models/quant.py:155-157 -------------------------------
// d(r, v) = fma(-2, <z_r, e_v>, |z_r|^2 + |e_v|^2) with every dot product a sequential fma chain over c = 0..31 (the fixed order
// of oracle/spec_c:sdvar_spec_nearest_code), argmin with the lowest index on ties.  32 rows of z per CTA live in shared memory;
// a thread owns codes tid, tid+256, ... (one code = 32 registers) and keeps the running best of all 32 rows in registers.
constexpr int kNcRows = 32;
__global__ void __launch_bounds__(256)
vq_nearest_code_kernel(const float* __restrict__ z, const float* __restrict__ codebook, long long N, int V,
                       long long* __restrict__ idx_out) {
  __shared__ __align__(16) float sz[kNcRows][kC];
  __shared__ float szsq[kNcRows];
  __shared__ unsigned long long sbest[8][kNcRows];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * kNcRows;
  for (int i = tid; i < kNcRows * kC; i += 256) {
    const long long r = row0 + i / kC;
    sz[i / kC][i % kC] = r < N ? z[r * kC + i % kC] : 0.0f;
  }
  __syncthreads();
  if (tid < kNcRows) {
    float a = 0.0f;
#pragma unroll
    for (int c = 0; c < kC; ++c) a = __fmaf_rn(sz[tid][c], sz[tid][c], a);
    szsq[tid] = a;
  }
  __syncthreads();
  float bd[kNcRows];
  int bi[kNcRows];
#pragma unroll
  for (int r = 0; r < kNcRows; ++r) { bd[r] = INFINITY; bi[r] = 0x7FFFFFFF; }
  for (int v = tid; v < V; v += 256) {
    float e[kC];
    const float4* e4 = reinterpret_cast<const float4*>(codebook + (long long)v * kC);
#pragma unroll
    for (int q = 0; q < kC / 4; ++q) {
      const float4 t = __ldg(e4 + q);
      e[4 * q] = t.x; e[4 * q + 1] = t.y; e[4 * q + 2] = t.z; e[4 * q + 3] = t.w;
    }
    float esq = 0.0f;
#pragma unroll
    for (int c = 0; c < kC; ++c) esq = __fmaf_rn(e[c], e[c], esq);
#pragma unroll
    for (int r = 0; r < kNcRows; ++r) {
      float dot = 0.0f;
#pragma unroll
      for (int c = 0; c < kC; ++c) dot = __fmaf_rn(sz[r][c], e[c], dot);
      const float d = __fmaf_rn(-2.0f, dot, __fadd_rn(szsq[r], esq));
      if (d < bd[r]) { bd[r] = d; bi[r] = v; }     // codes are visited in increasing order: strict < keeps the lowest index
    }
  }
  // (distance key, index) packed so that the u64 minimum is "smallest distance, then lowest index"
#pragma unroll
  for (int r = 0; r < kNcRows; ++r) {
    unsigned long long w = ((unsigned long long)fkey(bd[r]) << 32) | (unsigned long long)(uint32_t)bi[r];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, w, off);
      w = o < w ? o : w;
    }
    if (lane == 0) sbest[warp][r] = w;
  }
  __syncthreads();
  if (tid < kNcRows && row0 + tid < N) {
    unsigned long long w = sbest[0][tid];
#pragma unroll
    for (int k = 1; k < 8; ++k) w = sbest[k][tid] < w ? sbest[k][tid] : w;
    idx_out[row0 + tid] = (long long)(uint32_t)(w & 0xFFFFFFFFull);
  }
}

}  // namespace sdvar

using namespace sdvar;

extern "C" int sdvar_vq_nearest_code(const float* z_NC, const float* codebook, long long N, int Cvae, int V, long long* idx_out,
                                     void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(z_NC && codebook && idx_out, "NULL argument");
  SDVAR_REQUIRE(Cvae == kC, "Cvae=%d unsupported (32)", Cvae);
  SDVAR_REQUIRE(N > 0 && V > 0, "bad geometry N=%lld V=%d", N, V);
  SDVAR_REQUIRE(((uintptr_t)codebook & 15) == 0, "codebook must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(st, FAM_VQ, (double)N * (Cvae * 4.0 + 8.0) + (double)V * Cvae * 4.0);
  const long long blocks = (N + kNcRows - 1) / kNcRows;
  vq_nearest_code_kernel<<<(unsigned)blocks, 256, 0, st>>>(z_NC, codebook, N, V, idx_out);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

extern "C" int sdvar_vq_next_input(const long long* idx_Bl, int B, int pn, int HW, int pn_next, int Cvae,
                                   const float* codebook, const float* phi_w, const float* phi_b, float resi_ratio,
                                   float* f_hat, float* next_map, float* f_rest, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(resi_ratio >= 0.0f && resi_ratio <= 1.0f, "resi_ratio=%f outside [0,1]", (double)resi_ratio);
  SDVAR_REQUIRE(idx_Bl && codebook && phi_w && phi_b && f_hat, "NULL argument");
  SDVAR_REQUIRE(Cvae == kC, "Cvae=%d unsupported (32)", Cvae);
  SDVAR_REQUIRE(B > 0 && pn >= 1 && pn <= HW && HW <= kMaxHW, "bad geometry pn=%d HW=%d", pn, HW);
  SDVAR_REQUIRE(pn_next >= 0 && pn_next <= HW, "bad pn_next=%d", pn_next);
  SDVAR_REQUIRE(pn_next == 0 || next_map != nullptr, "next_map is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(st, FAM_VQ, (double)B * (2.0 * Cvae * HW * HW * 4 + 8.0 * pn * pn + 4.0 * Cvae * pn_next * pn_next));
  constexpr int R = 4;
  if (HW <= 16) {
    using Smem = VqSmemT<16, R>;
    SDVAR_SET_SMEM_ONCE((vq_accumulate_kernel<16, R>), sizeof(Smem));
    vq_accumulate_kernel<16, R><<<dim3((HW + R - 1) / R, B), 256, sizeof(Smem), st>>>(idx_Bl, pn, HW, codebook, phi_w, phi_b, resi_ratio, f_hat, f_rest);
  } else {
    using Smem = VqSmemT<kMaxHW, R>;
    SDVAR_SET_SMEM_ONCE((vq_accumulate_kernel<kMaxHW, R>), sizeof(Smem));
    vq_accumulate_kernel<kMaxHW, R><<<dim3((HW + R - 1) / R, B), 256, sizeof(Smem), st>>>(idx_Bl, pn, HW, codebook, phi_w, phi_b, resi_ratio, f_hat, f_rest);
  }
  SDVAR_LAUNCH_CHECK();
  if (pn_next > 0) {
    vq_area_down_kernel<<<B, 256, 0, st>>>(f_hat, HW, pn_next, next_map);
    SDVAR_LAUNCH_CHECK();
  }
  return SDVAR_OK;
}

extern "C" int sdvar_vq_area_down(const float* f_hat, int B, int HW, int pn_next, int Cvae, float* next_map, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(f_hat && next_map, "NULL argument");
  SDVAR_REQUIRE(Cvae == kC, "Cvae=%d unsupported (32)", Cvae);
  SDVAR_REQUIRE(B > 0 && HW >= 1 && HW <= kMaxHW && pn_next >= 1 && pn_next <= HW, "bad geometry HW=%d pn_next=%d", HW, pn_next);
  ProfileScope prof((cudaStream_t)stream, FAM_VQ, (double)B * Cvae * 4.0 * (HW * HW + pn_next * pn_next));
  vq_area_down_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(f_hat, HW, pn_next, next_map);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

extern "C" int sdvar_embed_next_map(const float* next_map, int B, int l, int Cvae, int C, const float* W_we,
                                    const float* b_we, const float* lvl_pos, float* x, int ldx_tokens, int tok_off,
                                    void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(next_map && W_we && b_we && lvl_pos && x, "NULL argument");
  SDVAR_REQUIRE(Cvae == kC, "Cvae=%d unsupported (32)", Cvae);
  SDVAR_REQUIRE(B > 0 && l > 0 && C > 0 && ldx_tokens >= tok_off + l, "bad geometry");
  SDVAR_REQUIRE(((uintptr_t)W_we & 15) == 0, "W_we must be 16-byte aligned");
  ProfileScope prof((cudaStream_t)stream, FAM_EMBED, (double)B * l * (4.0 * Cvae + 8.0 * C));
  embed_next_map_kernel<<<dim3((l + kEmbTok - 1) / kEmbTok, B), 256, 0, (cudaStream_t)stream>>>(
      next_map, B, l, C, W_we, b_we, lvl_pos, x, ldx_tokens, tok_off);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

extern "C" int sdvar_first_map(const float* cond_2BC, int B2, int first_l, int C, const float* pos_start,
                               const float* lvl_pos, float* x, int ldx_tokens, int tok_off, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(cond_2BC && pos_start && lvl_pos && x && B2 > 0 && first_l > 0 && C > 0, "bad argument");
  first_map_kernel<<<dim3(first_l, B2), 256, 0, (cudaStream_t)stream>>>(cond_2BC, first_l, C, pos_start, lvl_pos, x,
                                                                          ldx_tokens, tok_off);
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}
