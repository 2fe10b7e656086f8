/*
 * sdvar_b200.h -- C ABI of libsdvar_b200.so, the sm_100a kernels behind SDVAR's speculative
 * draft-then-verify next-scale generation loop.
 *
 * The reference (lijrjyan/SDVAR) has no FFI/plugin boundary: the path sits behind Python methods
 * (SURVEY.md 8b).  This header is the boundary a maintainer binds with ctypes;
 * each entry point cites the reference code it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; all launches are asynchronous on it, there are no
 *     hidden synchronisations, host read-backs or allocations => every call is CUDA-graph capturable;
 *   - return value: 0 = success, negative = sdvar_status; sdvar_last_error() describes the failure;
 *   - the library never allocates or frees caller tensors; scratch is passed in;
 *   - sm_100a only: on any other device every compute entry returns SDVAR_ERR_ARCH (no fallback);
 *   - tensors are dense row-major; `bf16` is stored as uint16_t.
 *   - CFG layout: row-blocks [0,B) are the conditional half, [B,2B) the unconditional half
 *     (models/var.py:162, 199-200).
 */
#ifndef SDVAR_B200_H_
#define SDVAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDVAR_ABI_VERSION 2
#define SDVAR_MAX_SEG 16   /* max stages in one launch (a pyramid has 10) */
#define SDVAR_MAX_DEPTH 64 /* max transformer blocks per model */

typedef enum {
  SDVAR_OK = 0,
  SDVAR_ERR_ARG = -1,    /* bad shape / alignment / null pointer */
  SDVAR_ERR_ARCH = -2,   /* device is not sm_100 */
  SDVAR_ERR_CUDA = -3,   /* a CUDA runtime / driver call failed */
  SDVAR_ERR_UNSUPPORTED = -4
} sdvar_status;

typedef uint16_t sdvar_bf16;

/* ---- library ------------------------------------------------------------------------------- */
int sdvar_abi_version(void);
const char* sdvar_last_error(void);       /* thread-local, valid until the next failing call */
int sdvar_arch_check(int device);         /* 0 iff `device` is compute capability 10.x */
int sdvar_num_sms(int device);

/* ---- K3: fused logits epilogue --------------------------------------------------------------
 * replaces models/var.py:199-202 (CFG mix) + models/helpers.py:6-19 (sample_with_top_k_top_p_).
 * rows are (b,pos), b<B, pos<L; cond logits at row b*in_ld+in_off+pos and uncond logits at row
 * (B+b)*in_ld+in_off+pos of logits_2BLV (fp32, V % 1024 == 0, V <= 8192): in_ld/in_off select one stage's slice
 * of a multi-stage verify window (in_ld=L, in_off=0 for a dense (2B,L,V) tensor).  Outputs go to row b*out_ld+out_off+pos
 * of idx_out / mixed_out / prob_out (out_ld=L, out_off=0: dense (B,L,..)), so the draft's per-stage launches can fill the
 * window-shaped buffers K4 reads.  seg_begin_host[S+1] partitions [0,L) into stages;
 * stage j uses t1[j]=fl32(1+t_j), t2[j]=fl32(t_j), t_j = cfg*si/(K-1):  x = cond*t1 - uncond*t2.
 * top_k<=0 disables top-k; one_minus_top_p<0 disables top-p (else it is fl32(1-top_p)).
 * noise (B*L,V) is the pre-drawn Exp(1) tensor torch.multinomial would draw (row b*L+pos); NULL => no sampling
 * (filter only).  Outputs (each may be NULL): idx_out int64; mixed_out the mixed logits
 * with removed entries set to -inf (the reference masks them in place); prob_out the sampled
 * token's probability under the filtered distribution.  Arithmetic is bit-exact to
 * oracle/spec_c/sdvar_spec.c:sdvar_spec_sample.  No workspace. */
int sdvar_sample_cfg_topk_topp(const float* logits_2BLV, int B, int L, int in_ld, int in_off, int out_ld, int out_off, int V,
                               const int* seg_begin_host, int S,
                               const float* t1_host, const float* t2_host, int top_k, float one_minus_top_p,
                               const float* noise, long long* idx_out, float* mixed_out, float* prob_out,
                               void* stream);

/* ---- K4: speculative verify -----------------------------------------------------------------
 * north_star item 3; nearest reference code models/var.py:1160-1227 (top-1 rule, see
 * sdvar_verify_top1 below).  Per token row (b,pos): p=softmax(xt), q=softmax(xd) (both the
 * mixed+filtered logits, (B,L,V) fp32), accept iff u*q[d] < p[d]; on reject
 * out = argmax(max(0,p-q)/noise) (argmax(p/noise) if the residual is identically 0), else out=d.
 * A whole verify window is ONE launch: seg_begin_host[S+1] partitions [0,L) into its stages.
 * u (B*L) and noise (B*L,V): row b*L+pos when stage_major_aux == 0; when stage_major_aux != 0 they are the per-stage
 * draws laid end to end, row B*seg_begin[j] + b*l_j + (pos - seg_begin[j]) for pos in stage j.
 * Per (image, stage): first_reject (index within the stage of the first rejected token, l_j if none)
 * and n_accept; per image accepted_stages = #leading stages without a reject;
 * summary[0]=min_b accepted_stages, [1]=#accepted tokens, [2]=#rejected tokens, [3]=0.
 * p_d_out/q_d_out (B,L) may be NULL.  workspace: sdvar_verify_workspace_bytes(B,S) bytes, zero on first use (the kernel
 * leaves it zero).  Bit-exact to oracle/spec_c/sdvar_spec.c:sdvar_spec_verify. */
long long sdvar_verify_workspace_bytes(int B, int S);
int sdvar_verify_accept_resample(const float* xt, const float* xd, const long long* draft_idx, const float* u,
                                 const float* noise, int stage_major_aux, int B, int L, int V, const int* seg_begin_host, int S,
                                 long long* out_idx, unsigned char* accept, float* p_d_out, float* q_d_out,
                                 int* first_reject, int* n_accept, int* accepted_stages, int* summary,
                                 int* workspace, void* stream);

/* reference rule (models/var.py:1199-1222): match[b,pos] = (argmax_v xt == draft_idx);
 * n_match (B,S) int32 = per (image,stage) match counts (the caller applies the >= 0.5 threshold on
 * the batch mean, var.py:1203,1217). */
int sdvar_verify_top1(const float* xt, const long long* draft_idx, int B, int L, int V, const int* seg_begin_host,
                      int S, unsigned char* match, int* n_match, void* stream);

/* ---- K5: VQ next-input ----------------------------------------------------------------------
 * replaces models/var.py:205-211 + models/quant.py:187-196 (get_next_autoregressive_input),
 * :199-206 (Phi), :218-226 (PhiPartiallyShared index):
 *   h = codebook[idx] (B,l,Cvae) -> (B,Cvae,pn,pn) -> bicubic up to HWxHW (skipped when pn==HW)
 *   d = (1-r)*h + r*(conv3x3(h; phi_w, phi_b)),  r = resi_ratio = |quant_resi| (0.5 in the released checkpoints)
 *   f_hat += d                                                  (in place, fp32 (B,Cvae,HW,HW))
 *   f_rest -= d  when f_rest != NULL                            (encode side: the running residual of models/quant.py:163)
 *   next_map = area-down(f_hat) to pn_next x pn_next            ((B,Cvae,pn_next,pn_next) fp32)
 * pn_next == 0 => last stage, next_map not written.  phi_w (Cvae,Cvae,3,3), phi_b (Cvae) are the
 * Phi module selected by the caller.  Cvae must be 32, HW <= 32.  No scratch is needed. */
int sdvar_vq_next_input(const long long* idx_Bl, int B, int pn, int HW, int pn_next, int Cvae, const float* codebook,
                        const float* phi_w, const float* phi_b, float resi_ratio, float* f_hat, float* next_map,
                        float* f_rest, void* stream);

/* the area-down half of the step above on its own (models/quant.py:192, F.interpolate(mode='area')): next_map
 * (B,Cvae,pn_next,pn_next) = adaptive average of f_hat (B,Cvae,HW,HW) over windows [floor(o*HW/pn), ceil((o+1)*HW/pn)).  Same
 * kernel, so a stage input rebuilt from a committed f_hat is bit-identical to the one sdvar_vq_next_input returned. */
int sdvar_vq_area_down(const float* f_hat, int B, int HW, int pn_next, int Cvae, float* next_map, void* stream);

/* test hook: the kernels' exponential (oracle/spec_c: sdvar_spec_expf) applied element-wise through the packed fp32x2 and the
 * scalar code path, so the arithmetic spec can be pinned bit for bit on adversarial inputs.  x[n] <= 0 (or -inf). */
int sdvar_debug_spec_expf(const float* x, long long n, float* y_packed, float* y_scalar, void* stream);

/* encode side: nearest codebook entry per row, reference models/quant.py:155-157
 * (`d = |z|^2 + |e|^2 - 2 z e^T; argmin`).  z_NC (N,Cvae) fp32, codebook (V,Cvae) fp32, idx_out (N) int64.
 * d = fma(-2, <z,e>, |z|^2 + |e|^2), dot products as sequential fma chains over c (oracle/spec_c: sdvar_spec_nearest_code),
 * lowest index on ties.  Cvae must be 32. */
int sdvar_vq_nearest_code(const float* z_NC, const float* codebook, long long N, int Cvae, int V, long long* idx_out, void* stream);

/* stage input map (models/var.py:185-188):  x[r, t, :] = W_we @ next_map[b, :, t] + b_we + lvl_pos[t, :]
 * for r in {b, B+b} (the CFG repeat), next_map (B,Cvae,l) fp32, W_we (C,Cvae), lvl_pos (l,C) slice for the
 * stage, x (2B, ldx_tokens, C) fp32 written at token offset tok_off with ldx_tokens tokens per image. */
int sdvar_embed_next_map(const float* next_map, int B, int l, int Cvae, int C, const float* W_we, const float* b_we,
                         const float* lvl_pos, float* x, int ldx_tokens, int tok_off, void* stream);

/* first stage map (models/var.py:179-183): x[r,t,:] = cond[r,:] + pos_start[t,:] + lvl_pos[t,:] */
int sdvar_first_map(const float* cond_2BC, int B2, int first_l, int C, const float* pos_start, const float* lvl_pos,
                    float* x, int ldx_tokens, int tok_off, void* stream);

/* ---- transformer pieces ----------------------------------------------------------------------
 * LayerNorm (eps, no affine) + adaLN modulate, fp32 in -> bf16 out (models/basic_var.py:157-158,173):
 *   out[r,:] = LN(x[r,:]) * (1 + scale[img(r),:]) + shift[img(r),:],  img(r) = r / tokens_per_img;
 * scale/shift are rows of an fp32 matrix with leading dimension ld_mod (the adaLN output).
 * slot_map (device int32[#images of the pass], or NULL = identity): per-image resources (adaLN rows here, KV-cache slots in
 * the QKV epilogue and the attention) of pass-image i live at slot slot_map[i].  This is how a SUB-BATCH of images that
 * share a stage runs as one dense pass while every image keeps its own cache slot (per-image ragged acceptance,
 * SURVEY.md 8f #2; replaces the batch-global accept_length of models/var.py:1349-1350). */
int sdvar_ln_modulate(const float* x, int M, int C, int tokens_per_img, const float* scale, const float* shift,
                      int ld_mod, const int* slot_map, float eps, sdvar_bf16* out, void* stream);

/* out = silu(x) as bf16 (the SiLU in front of every ada_lin, models/basic_var.py:147,170) */
int sdvar_silu_bf16(const float* x, long long n, sdvar_bf16* out, void* stream);
int sdvar_f32_to_bf16(const float* x, long long n, sdvar_bf16* out, void* stream);

/* image hand-over: fp32 images in [0,1] (B,3,H,W) -> uint8 = trunc(clamp(x,0,1)*255), what the reference's notebook writes
 * to PNG (sdvar_colab_test.py:235-236) and what the data-parallel gather ships (4x fewer bytes than fp32).
 * hwc == 0: out (B,3,H,W); hwc != 0: out (B,H,W,3), the layout of the FID .npz (utils/misc.py:360-381). */
int sdvar_image_to_u8(const float* img_B3HW, int B, int H, int W, int hwc, uint8_t* out, void* stream);

/* K1: D = epilogue(A[M,K] @ W[N,K]^T), bf16 operands, fp32 accumulation in TMEM (tcgen05.mma),
 * operands staged by TMA.  Replaces F.linear at models/basic_var.py:52,93,119,156 and
 * models/var.py:125.  K % 64 == 0, N % 32 == 0, A/W 16-byte aligned rows. */
typedef enum {
  SDVAR_EPI_F32 = 0,        /* out_f32[M,N] = acc + bias                                  (head, ada_lin) */
  SDVAR_EPI_BF16 = 1,       /* out_bf16[M,N] = acc + bias                                                 */
  SDVAR_EPI_GELU_BF16 = 2,  /* out_bf16[M,N] = gelu_tanh(acc + bias)                       (fc1)          */
  SDVAR_EPI_RESID_F32 = 3,  /* out_f32[M,N] += (acc + bias) * gate[img(r), :]              (proj, fc2)    */
  SDVAR_EPI_QKV = 4         /* q/k l2-norm + scale, KV-cache append                        (mat_qkv)      */
} sdvar_epilogue;

typedef struct {
  int epilogue;              /* sdvar_epilogue */
  const float* bias;         /* [N] or NULL */
  float* out_f32;            /* F32 / RESID_F32 */
  sdvar_bf16* out_bf16;      /* BF16 / GELU_BF16 */
  int ldo;                   /* leading dimension of out (elements) */
  /* RESID_F32: gate row = gate + img(r)*ld_gate, img(r) = r / tokens_per_img */
  const float* gate;
  int ld_gate;
  int tokens_per_img;
  const int* slot_map;       /* device int32[#images] or NULL: gate row / KV-cache slot of pass-image i (see sdvar_ln_modulate) */
  /* QKV (models/basic_var.py:93-109): N = 3*H*64; rows r = img*Lq + t.  q -> q_out[img,h,t,:] bf16
   * (normalised, times exp(min(scale_mul[h], ln 100)) when l2norm), k -> k_cache[img,h,kv_off+t,:]
   * (normalised), v -> vT_cache[img,h,:,kv_off+t] (transposed). bias = cat(q_bias,0,v_bias). */
  sdvar_bf16* q_out;         /* (imgs, H, Lq, 64) */
  sdvar_bf16* k_cache;       /* (imgs, H, Lmax, 64) */
  sdvar_bf16* vT_cache;      /* (imgs, H, 64, Lmax_pad) */
  const float* scale_mul;    /* [H] raw parameter (log-scale) */
  int H, Lq, Lmax, Lmax_pad, kv_off, l2norm;
} sdvar_gemm_epilogue;

int sdvar_gemm_bf16(const sdvar_bf16* A, int lda, const sdvar_bf16* W, int ldw, int M, int N, int K,
                    const sdvar_gemm_epilogue* epi_host, void* stream);

/* K2: attention over the KV cache written by the QKV epilogue (models/basic_var.py:107-117).
 * q (imgs,H,Lq,64) bf16; k_cache (imgs,H,Lmax,64); vT_cache (imgs,H,64,Lmax_pad); out (imgs*Lq, H*64)
 * bf16.  Query token t of the launch belongs to window stage j (seg_begin_host over [0,Lq)) and sees
 * keys [0, kv_off + seg_begin[j+1]) -- block-causal inside the window, everything before it
 * (models/var.py:108-113; incremental decode is S=1).  softmax scale `scale` (1.0 with l2 norm).
 * logit_bound_log (device, [H], may be NULL): when given, the caller guarantees |q.k| <= exp(min(logit_bound_log[h], ln 100)) <= 40
 * for every head (true for l2-normalised attention, where it is the scale_mul parameter): the kernel then uses that bound
 * as the softmax reference point and runs the one-pass ping-pong variant; NULL selects the general two-pass kernel.
 * slot_map / cache_slots: pass-image i reads the K/V of cache slot slot_map[i] (< cache_slots = image slots the caches hold);
 * NULL / 0 = identity with cache_slots = imgs. */
int sdvar_attention(const sdvar_bf16* q, const sdvar_bf16* k_cache, const sdvar_bf16* vT_cache, int imgs, int H,
                    int Lq, int Lmax, int Lmax_pad, int kv_off, const int* seg_begin_host, int S, float scale,
                    const float* logit_bound_log, const int* slot_map, int cache_slots, sdvar_bf16* out, void* stream);

/* ---- decoder boundary ------------------------------------------------------------
 * GroupNorm(32 groups, eps, affine) optionally followed by SiLU on channels-last bf16 activations (reference
 * models/basic_vae.py:18-19, 57-59).  x, y: (N, H*W, C) with C contiguous; y may alias x.  scratch: N*128*64 floats.
 * pre_bias (nullable, fp32 [C]) is added to x before the statistics: the bias of the convolution that produced x, so that
 * convolution can run bias-free (no separate bias pass over the activation). */
int sdvar_groupnorm_silu_nhwc(const sdvar_bf16* x, const float* pre_bias, int N, int HW, int C, const float* gamma, const float* beta,
                              float eps, int silu, sdvar_bf16* y, float* scratch, void* stream);
/* out = h + bias[c] (+ res): conv bias + skip connection of a residual block (models/basic_vae.py:61) in one pass.
 * rows = N*H*W pixels, C % 8 == 0, channels-last bf16; res may be NULL; out may alias h or res. */
int sdvar_bias_residual_nhwc(const sdvar_bf16* h, const float* bias, const sdvar_bf16* res, long long rows, int C, sdvar_bf16* out,
                             void* stream);
/* Decoder convolution as a tcgen05 implicit GEMM (replaces nn.Conv2d(k=3,padding=1) / nn.Conv2d(k=1) at models/basic_vae.py:22-27,
 * 44-52, 75-76, 171-196, models/vqvae.py:38-39).  x (N,H,W,Cin) channels-last bf16, Cin % 32 == 0; w_packed (taps, Cout, Cin) bf16
 * = weight.permute(2,3,0,1) of the (Cout,Cin,kh,kw) parameter, taps = 9 or 1; bias fp32 [Cout] or NULL; res (N,H,W,Cout) bf16 skip
 * connection or NULL.  Exactly one output: y (N,H,W,Cout) bf16 (Cout % 8 == 0), or y_f32_nchw (N,Cout,H,W) fp32 clamped to
 * [lo,hi] (conv_out + the clamp of models/vqvae.py:63).  128 consecutive pixels in (n,y,x) order must form a box: W divides 128
 * or is a multiple of it, H likewise for the rows left.  Zero padding comes from the TMA unit's out-of-bounds fill. */
int sdvar_conv_nhwc(const sdvar_bf16* x, int N, int H, int W, int Cin, const sdvar_bf16* w_packed, int taps, int Cout,
                    const float* bias, const sdvar_bf16* res, sdvar_bf16* y, float* y_f32_nchw, float lo, float hi, void* stream);
/* Upsample2x of the decoder in one step (models/basic_vae.py:28-33: F.interpolate(scale_factor=2, mode='nearest') then a 3x3
 * convolution): y (N,2H,2W,Cout) = conv3x3(nearest2x(x)) + bias, computed as four 2x2 convolutions on the low-resolution x
 * (N,H,W,Cin), one per output parity (a,b).  w_par (16, Cout, Cin) bf16: matrix (a*2+b)*4 + (u*2+v) is the sum of the 3x3 taps
 * that fall on input pixel (Y+a-1+u, X+b-1+v) of output pixel (2Y+a, 2X+b).  Same geometry rules as sdvar_conv_nhwc. */
int sdvar_conv_up2x_nhwc(const sdvar_bf16* x, int N, int H, int W, int Cin, const sdvar_bf16* w_par, int Cout, const float* bias,
                         sdvar_bf16* y, void* stream);
/* nearest-neighbour 2x upsampling (models/basic_vae.py:31), channels-last bf16: x (N,H,W,C) -> y (N,2H,2W,C). */
int sdvar_upsample2x_nhwc(const sdvar_bf16* x, int N, int H, int W, int C, sdvar_bf16* y, void* stream);

/* ---- whole transformer pass (the launch sequence of one stage / one verify window) -----------
 * Device-pointer table of one VAR model in the engine's packed layout (bf16 weights, fp32 vectors).
 * Replaces the python block loop models/var.py:195-197 / :973-976 / :1051-1055. */
typedef struct {
  int depth, C, H, V, Cvae, l2norm;
  float eps, attn_scale;
  const sdvar_bf16* w_qkv[SDVAR_MAX_DEPTH];   /* (3C, C) */
  const float* b_qkv[SDVAR_MAX_DEPTH];        /* (3C) = cat(q_bias, 0, v_bias) */
  const float* scale_mul[SDVAR_MAX_DEPTH];    /* (H) */
  const sdvar_bf16* w_proj[SDVAR_MAX_DEPTH];  /* (C, C) */
  const float* b_proj[SDVAR_MAX_DEPTH];
  const sdvar_bf16* w_fc1[SDVAR_MAX_DEPTH];   /* (4C, C) */
  const float* b_fc1[SDVAR_MAX_DEPTH];
  const sdvar_bf16* w_fc2[SDVAR_MAX_DEPTH];   /* (C, 4C) */
  const float* b_fc2[SDVAR_MAX_DEPTH];
  const sdvar_bf16* w_head;                   /* (V, C) */
  const float* b_head;
  int attn_fixed_max;                         /* 1: l2norm and exp(min(scale_mul, ln 100)) <= 40 for every head (one-pass attention) */
} sdvar_var_weights;

typedef struct {
  int imgs;                 /* 2B */
  int Lq;                   /* tokens per image in this pass */
  int Lmax, Lmax_pad;       /* KV cache geometry */
  int kv_off;               /* tokens already in the cache */
  int S;                    /* window stages */
  int seg_begin[SDVAR_MAX_SEG + 1];
  const int* slot_map;      /* device int32[imgs] or NULL: cache / adaLN slot of pass-image i (sub-batch passes) */
  int cache_slots;          /* image slots held by k_cache / vT_cache / ada / head_mod (0 = imgs) */
  float* x;                 /* (imgs*Lq, C) fp32 residual stream, in/out */
  const float* ada;         /* adaLN rows [gamma1 gamma2 scale1 scale2 shift1 shift2] (6C fp32) of block i, image r at
                               ada + i*ada_block_stride + r*ada_img_stride (strides in floats) */
  long long ada_block_stride;
  long long ada_img_stride;
  const float* head_mod;    /* (imgs, 2C) fp32: scale, shift of AdaLNBeforeHead */
  sdvar_bf16* k_cache[SDVAR_MAX_DEPTH];
  sdvar_bf16* vT_cache[SDVAR_MAX_DEPTH];
  sdvar_bf16* xm;           /* (imgs*Lq, C)  bf16 scratch: modulated LN output / attention output */
  sdvar_bf16* q;            /* (imgs, H, Lq, 64) bf16 scratch */
  sdvar_bf16* attn;         /* (imgs*Lq, C) bf16 scratch */
  sdvar_bf16* hidden;       /* (imgs*Lq, 4C) bf16 scratch */
  float* logits;            /* (imgs*Lq, V) fp32 out, or NULL to skip the head */
} sdvar_pass;

int sdvar_var_forward(const sdvar_var_weights* w_host, const sdvar_pass* pass_host, void* stream);

/* Optional per-family device timing for bench.py's roofline leg (off by default, zero cost when off).  After
 * sdvar_profile_begin() every entry point brackets its launches with a cudaEvent pair on the launch stream;
 * sdvar_profile_end() synchronises the device and returns, per family, the summed milliseconds, the summed
 * ALGORITHMIC work (FLOPs for GEMM/ATTN, bytes for the others, as defined in DESIGN.md) and the launch count. */
#define SDVAR_PROFILE_FAMILIES 9 /* 0 GEMM, 1 ATTN, 2 LN, 3 SAMPLE, 4 VERIFY, 5 VQ, 6 EMBED, 7 MISC, 8 CONV */
int sdvar_profile_begin(void);
int sdvar_profile_end(double* ms, double* work, long long* launches);

/* number of kernels the library has launched since load (gpu_launches accounting in bench.py) */
long long sdvar_launch_count(void);
/* add n to that counter: called by a host that replays a captured CUDA graph of n library launches */
void sdvar_count_launches(long long n);

#ifdef __cplusplus
}
#endif
#endif /* SDVAR_B200_H_ */
