// groupnorm.cu -- GroupNorm(32 groups, eps, affine) [+ SiLU] on channels-last bf16 activations, for the VQVAE decoder that
// turns the path's f_hat into pixels (reference models/basic_vae.py:18-19 `Normalize`, :57-59 `F.silu(norm(x))`).
//
// The decoder is the boundary right after the hot path; its convolutions stay on cuDNN for now, but PyTorch's
// GroupNorm falls back to NCHW copies for channels-last bf16 (63 ms + 60 ms of layout copies per 64-image batch, measured), so
// this HBM-bound piece is done here: two passes over the tensor (statistics, then normalise+SiLU), 16-byte vector accesses,
// deterministic two-stage reduction (per-block partials, summed in a fixed order by every consumer).
#include <stdlib.h>

#include "common.cuh"

namespace sdvar {

__device__ __forceinline__ float tanh_approx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

constexpr int kGnThreads = 256;
constexpr int kGnMaxC = 1024;

// partial[(n*nblk + blk)*2*32 + g*2 + {0,1}] = sum / sum of squares over this block's pixels
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pre_bias, int HW, int C, int pix_per_blk,
                float* __restrict__ partial) {
  __shared__ float s_sum[32], s_sq[32];
  const int n = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
  const int vec_per_pix = C >> 3;                 // 8 channels (16 bytes) per vector
  const int cg = C >> 5;                          // channels per group
  if (threadIdx.x < 32) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
  __syncthreads();
  const int p0 = blk * pix_per_blk, p1 = min(p0 + pix_per_blk, HW);
  const long long base = (long long)n * HW * C;
  const int total = (p1 - p0) * vec_per_pix;
  // blockDim.x is a multiple of vec_per_pix (host guarantees it), so a thread keeps one fixed 8-channel vector index and
  // accumulates its 8 channels in registers over all its pixels; groups are combined once at the end
  float acc_s[8], acc_q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc_s[k] = 0.f; acc_q[k] = 0.f; }
  const int v = threadIdx.x % vec_per_pix;
  // conv bias folded in (statistics of x + b), and every value is taken relative to a per-(image, group) PIVOT -- the group's
  // first element -- so that the one-pass variance E[d^2] - E[d]^2 does not cancel when |mean| >> std
  float pb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k, c0 = (c / cg) * cg;
    const float piv = __bfloat162float(x[base + c0]) + (pre_bias ? pre_bias[c0] : 0.f);
    pb[k] = (pre_bias ? pre_bias[c] : 0.f) - piv;
  }
  // blockDim.x is a multiple of vec_per_pix: a thread's pixel advances by a fixed step, so no division in the loop, and four
  // independent 16-byte loads are in flight per thread (one load per iteration left the kernel latency-bound at 3.1 TB/s)
  (void)total;
  const int pstep = (int)blockDim.x / vec_per_pix;
  const __nv_bfloat16* xp = x + base + v * 8;
  auto accum = [&](const uint4& raw) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(h[k]);
      f.x += pb[2 * k]; f.y += pb[2 * k + 1];
      acc_s[2 * k] += f.x; acc_q[2 * k] += f.x * f.x;
      acc_s[2 * k + 1] += f.y; acc_q[2 * k + 1] += f.y * f.y;
    }
  };
  int pix = p0 + (int)threadIdx.x / vec_per_pix;
  for (; pix + 3 * pstep < p1; pix += 4 * pstep) {
    const uint4 r0 = *reinterpret_cast<const uint4*>(xp + (long long)pix * C);
    const uint4 r1 = *reinterpret_cast<const uint4*>(xp + (long long)(pix + pstep) * C);
    const uint4 r2 = *reinterpret_cast<const uint4*>(xp + (long long)(pix + 2 * pstep) * C);
    const uint4 r3 = *reinterpret_cast<const uint4*>(xp + (long long)(pix + 3 * pstep) * C);
    accum(r0); accum(r1); accum(r2); accum(r3);
  }
  for (; pix < p1; pix += pstep) accum(*reinterpret_cast<const uint4*>(xp + (long long)pix * C));
  // deterministic block reduction: per-thread channel partials go to shared memory, then 64 threads (32 groups x {sum, sumsq})
  // each add up their group's channels over the threads that own them, in a fixed order
  __shared__ float s_part[16][kGnThreads];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s_part[k][threadIdx.x] = acc_s[k]; s_part[8 + k][threadIdx.x] = acc_q[k]; }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    float tsum = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      const int vv = c >> 3, kk = (c & 7) + which * 8;
      for (int t = vv; t < (int)blockDim.x; t += vec_per_pix) tsum += s_part[kk][t];
    }
    if (which == 0) s_sum[g] = tsum; else s_sq[g] = tsum;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float* o = partial + ((long long)n * nblk + blk) * 64;
    o[2 * threadIdx.x] = s_sum[threadIdx.x];
    o[2 * threadIdx.x + 1] = s_sq[threadIdx.x];
  }
}

__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ pre_bias, int HW, int C, int pix_per_blk, int nblk_stats,
                const float* __restrict__ partial,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu, __nv_bfloat16* __restrict__ y) {
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.y, blk = blockIdx.x;
  const int cg = C >> 5, vec_per_pix = C >> 3;
  if (threadIdx.x < 32) {
    float s = 0.f, q = 0.f;
    const float* pp = partial + (long long)n * nblk_stats * 64 + 2 * threadIdx.x;
    for (int b = 0; b < nblk_stats; ++b) { s += pp[b * 64]; q += pp[b * 64 + 1]; }   // fixed order => deterministic
    const float cnt = (float)HW * (float)cg;
    const int c0 = threadIdx.x * cg;
    const float piv = __bfloat162float(x[(long long)n * HW * C + c0]) + (pre_bias ? pre_bias[c0] : 0.f);   // the pivot gn_stats used
    const float dm = s / cnt;
    const float mean = piv + dm;
    const float var = fmaxf(q / cnt - dm * dm, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  // blockDim.x is a multiple of vec_per_pix: a thread keeps ONE 8-channel vector index, so its scale / shift pairs live in
  // registers (y = x*a + b) and its pixel advances by a fixed step -- no division, no shared-memory reads in the loop
  const int v = (int)threadIdx.x % vec_per_pix;
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = v * 8 + k, g = c / cg;
    a[k] = s_rstd[g] * gamma[c];
    b[k] = beta[c] + ((pre_bias ? pre_bias[c] : 0.f) - s_mean[g]) * a[k];
  }
  const int p0 = blk * pix_per_blk, p1 = min(p0 + pix_per_blk, HW);
  const long long base = (long long)n * HW * C + v * 8;
  const int pstep = (int)blockDim.x / vec_per_pix;
  auto apply = [&](const uint4& raw, long long off) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h[k]);
      float y0 = f.x * a[2 * k] + b[2 * k];
      float y1 = f.y * a[2 * k + 1] + b[2 * k + 1];
      if (silu == 1) {          // x * sigmoid(x) with sigmoid(x) = 0.5 * tanh(x / 2) + 0.5: ONE MUFU op per element (tanh.approx) instead of two
        y0 = fmaf(0.5f * y0, tanh_approx(0.5f * y0), 0.5f * y0);      // (ex2 + rcp) -- at 2 per element the 256 x 256 layers were MUFU-bound
        y1 = fmaf(0.5f * y1, tanh_approx(0.5f * y1), 0.5f * y1);
      } else if (silu == 2) {   // A/B: the exp + divide form
        y0 = __fdividef(y0, 1.0f + __expf(-y0)); y1 = __fdividef(y1, 1.0f + __expf(-y1));
      }
      o[k] = pack_bf16x2(y0, y1);
    }
    *reinterpret_cast<uint4*>(y + off) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  int pix = p0 + (int)threadIdx.x / vec_per_pix;
  for (; pix + pstep < p1; pix += 2 * pstep) {
    const long long o0 = base + (long long)pix * C, o1 = base + (long long)(pix + pstep) * C;
    const uint4 r0 = *reinterpret_cast<const uint4*>(x + o0);
    const uint4 r1 = *reinterpret_cast<const uint4*>(x + o1);
    apply(r0, o0); apply(r1, o1);
  }
  for (; pix < p1; pix += pstep) {
    const long long o0 = base + (long long)pix * C;
    apply(*reinterpret_cast<const uint4*>(x + o0), o0);
  }
}

// out = h + bias [+ res]: the tail of a residual block (conv2 bias + skip connection, reference models/basic_vae.py:61) in one
// pass instead of cuDNN's separate bias kernel followed by an elementwise add.  One 16-byte vector (8 channels) per thread step.
__global__ void __launch_bounds__(256)
bias_residual_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ bias, const __nv_bfloat16* __restrict__ res,
                     long long nvec, int vec_per_pix, __nv_bfloat16* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const int v = (int)(i % vec_per_pix);
    const uint4 a = reinterpret_cast<const uint4*>(h)[i];
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * v), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * v + 1);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint4 r = make_uint4(0, 0, 0, 0);
    if (res) r = reinterpret_cast<const uint4*>(res)[i];
    const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&r);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __bfloat1622float2(ha[k]), fr = __bfloat1622float2(hr[k]);
      o[k] = pack_bf16x2(fa.x + bb[2 * k] + fr.x, fa.y + bb[2 * k + 1] + fr.y);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// nearest-neighbour 2x upsampling, channels-last (reference models/basic_vae.py:31 F.interpolate(scale_factor=2, mode='nearest')):
// each input vector is read once and written to its four output pixels.
__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int vec_per_pix, long long nvec, __nv_bfloat16* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const int v = (int)(i % vec_per_pix);
    const long long pix = i / vec_per_pix;
    const int w = (int)(pix % W);
    const long long nh = pix / W;            // n*H + h
    const uint4 a = reinterpret_cast<const uint4*>(x)[i];
    uint4* o = reinterpret_cast<uint4*>(y) + ((nh * 2) * (2 * W) + 2 * w) * vec_per_pix + v;
    o[0] = a;
    o[vec_per_pix] = a;
    o[(long long)2 * W * vec_per_pix] = a;
    o[(long long)2 * W * vec_per_pix + vec_per_pix] = a;
  }
}

}  // namespace sdvar

using namespace sdvar;

// x, y: (N, H*W, C) channels-last bf16 (y may alias x); pre_bias (nullable) fp32 [C] is added to x first; gamma/beta fp32 [C]; scratch >= N*nblk*64 floats with
// nblk = min(ceil(HW/32), 128).  32 groups (reference Normalize), C % 32 == 0, C % 8 == 0, C <= 1024.
extern "C" int sdvar_groupnorm_silu_nhwc(const sdvar_bf16* x, const float* pre_bias, int N, int HW, int C, const float* gamma,
                                         const float* beta, float eps, int silu, sdvar_bf16* y, float* scratch, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && y && gamma && beta && scratch, "NULL argument");
  SDVAR_REQUIRE(N > 0 && HW > 0 && C % 32 == 0 && C % 8 == 0 && C <= kGnMaxC, "bad geometry N=%d HW=%d C=%d", N, HW, C);
  SDVAR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "16-byte alignment");
  int nblk = (HW + 31) / 32;
  if (nblk > 128) nblk = 128;
  const int ppb = (HW + nblk - 1) / nblk;
  nblk = (HW + ppb - 1) / ppb;
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(st, FAM_MISC, (double)N * HW * C * 2.0 * 3.0);
  const int vpp = C >> 3;
  SDVAR_REQUIRE(vpp <= kGnThreads, "C too large");
  const int stat_threads = (kGnThreads / vpp) * vpp;
  gn_stats_kernel<<<dim3(nblk, N), stat_threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), pre_bias, HW, C, ppb, scratch);
  SDVAR_LAUNCH_CHECK();
  static const bool exp_form = getenv("SDVAR_GN_SILU_EXP") != nullptr;   // A/B switch
  gn_apply_kernel<<<dim3(nblk, N), stat_threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), pre_bias, HW, C, ppb, nblk, scratch, gamma, beta, eps,
                                                       silu ? (exp_form ? 2 : 1) : 0, reinterpret_cast<__nv_bfloat16*>(y));
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

// out[r, c] = h[r, c] + bias[c] (+ res[r, c]); rows = N*H*W pixels, channels-last bf16, C % 8 == 0; out may alias h or res.
extern "C" int sdvar_bias_residual_nhwc(const sdvar_bf16* h, const float* bias, const sdvar_bf16* res, long long rows, int C,
                                        sdvar_bf16* out, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(h && bias && out, "NULL argument");
  SDVAR_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, "bad geometry rows=%lld C=%d", rows, C);
  SDVAR_REQUIRE(((uintptr_t)h & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)res & 15) == 0 && ((uintptr_t)bias & 15) == 0,
                "16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nvec = rows * (C >> 3);
  ProfileScope prof(st, FAM_MISC, (double)nvec * 16.0 * (res ? 3.0 : 2.0));
  long long blocks = (nvec + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  bias_residual_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(h), bias,
                                                         reinterpret_cast<const __nv_bfloat16*>(res), nvec, C >> 3,
                                                         reinterpret_cast<__nv_bfloat16*>(out));
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}

// y (N, 2H, 2W, C) = nearest-neighbour 2x of x (N, H, W, C), channels-last bf16, C % 8 == 0; x and y must not overlap.
extern "C" int sdvar_upsample2x_nhwc(const sdvar_bf16* x, int N, int H, int W, int C, sdvar_bf16* y, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(x && y && x != y, "NULL or aliased argument");
  SDVAR_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad geometry N=%d H=%d W=%d C=%d", N, H, W, C);
  SDVAR_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nvec = (long long)N * H * W * (C >> 3);
  ProfileScope prof(st, FAM_MISC, (double)nvec * 16.0 * 5.0);
  long long blocks = (nvec + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  upsample2x_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), H, W, C >> 3, nvec,
                                                      reinterpret_cast<__nv_bfloat16*>(y));
  SDVAR_LAUNCH_CHECK();
  return SDVAR_OK;
}
