// forward.cu -- sdvar_var_forward: the launch sequence of one transformer pass (one stage of incremental
// decode, or one multi-stage verify window) over the packed weights of one VAR model.
//
// Replaces the python block loops models/var.py:195-197, :973-976, :1051-1055 and AdaLNSelfAttn.forward
// (models/basic_var.py:152-159): 7 kernels per block, all asynchronous on the caller's stream, so a whole pass
// is one C call and can be captured into a CUDA graph.  ada_lin(cond) depends only on the class label and is
// computed once per generation by the caller, not once per stage as in the reference.
#include "common.cuh"

using namespace sdvar;

extern "C" int sdvar_var_forward(const sdvar_var_weights* w, const sdvar_pass* ps, void* stream) {
  if (int rc = check_arch()) return rc;
  SDVAR_REQUIRE(w && ps, "NULL argument");
  SDVAR_REQUIRE(w->depth > 0 && w->depth <= SDVAR_MAX_DEPTH, "depth=%d out of range", w->depth);
  SDVAR_REQUIRE(w->C == w->H * 64, "C=%d must equal 64*H (H=%d)", w->C, w->H);
  SDVAR_REQUIRE(ps->imgs > 0 && ps->Lq > 0 && ps->S >= 1 && ps->S <= SDVAR_MAX_SEG, "bad pass geometry");
  SDVAR_REQUIRE(ps->x && ps->ada && ps->xm && ps->q && ps->attn && ps->hidden, "NULL pass buffer");
  SDVAR_REQUIRE(ps->slot_map != nullptr || ps->cache_slots == 0 || ps->cache_slots == ps->imgs, "cache_slots without a slot map");
  const int C = w->C, M = ps->imgs * ps->Lq;
  const int ldm = (int)ps->ada_img_stride;
  SDVAR_REQUIRE(ps->ada_img_stride >= 6 * C && ps->ada_img_stride % 4 == 0 && ps->ada_block_stride % 4 == 0, "bad adaLN strides");
  for (int i = 0; i < w->depth; ++i) {
    const float* ada = ps->ada + (size_t)i * ps->ada_block_stride;  // rows: [gamma1 gamma2 scale1 scale2 shift1 shift2]
    int rc;
    if ((rc = sdvar_ln_modulate(ps->x, M, C, ps->Lq, ada + 2 * C, ada + 4 * C, ldm, ps->slot_map, w->eps, ps->xm, stream))) return rc;
    sdvar_gemm_epilogue e{};
    e.epilogue = SDVAR_EPI_QKV;
    e.bias = w->b_qkv[i];
    e.q_out = ps->q; e.k_cache = ps->k_cache[i]; e.vT_cache = ps->vT_cache[i];
    e.scale_mul = w->scale_mul[i];
    e.slot_map = ps->slot_map;
    e.H = w->H; e.Lq = ps->Lq; e.Lmax = ps->Lmax; e.Lmax_pad = ps->Lmax_pad; e.kv_off = ps->kv_off; e.l2norm = w->l2norm;
    if ((rc = sdvar_gemm_bf16(ps->xm, C, w->w_qkv[i], C, M, 3 * C, C, &e, stream))) return rc;
    if ((rc = sdvar_attention(ps->q, ps->k_cache[i], ps->vT_cache[i], ps->imgs, w->H, ps->Lq, ps->Lmax, ps->Lmax_pad,
                              ps->kv_off, ps->seg_begin, ps->S, w->attn_scale,
                              (w->attn_fixed_max && w->l2norm) ? w->scale_mul[i] : nullptr, ps->slot_map, ps->cache_slots, ps->attn, stream)))
      return rc;
    sdvar_gemm_epilogue r{};
    r.epilogue = SDVAR_EPI_RESID_F32;
    r.bias = w->b_proj[i];
    r.out_f32 = ps->x; r.ldo = C;
    r.gate = ada; r.ld_gate = ldm; r.tokens_per_img = ps->Lq; r.slot_map = ps->slot_map;
    if ((rc = sdvar_gemm_bf16(ps->attn, C, w->w_proj[i], C, M, C, C, &r, stream))) return rc;
    if ((rc = sdvar_ln_modulate(ps->x, M, C, ps->Lq, ada + 3 * C, ada + 5 * C, ldm, ps->slot_map, w->eps, ps->xm, stream))) return rc;
    sdvar_gemm_epilogue g{};
    g.epilogue = SDVAR_EPI_GELU_BF16;
    g.bias = w->b_fc1[i];
    g.out_bf16 = ps->hidden; g.ldo = 4 * C;
    if ((rc = sdvar_gemm_bf16(ps->xm, C, w->w_fc1[i], C, M, 4 * C, C, &g, stream))) return rc;
    r.bias = w->b_fc2[i];
    r.gate = ada + C;
    if ((rc = sdvar_gemm_bf16(ps->hidden, 4 * C, w->w_fc2[i], 4 * C, M, C, 4 * C, &r, stream))) return rc;
  }
  if (ps->logits != nullptr) {
    SDVAR_REQUIRE(ps->head_mod && w->w_head, "head requested without head_mod / w_head");
    int rc;
    if ((rc = sdvar_ln_modulate(ps->x, M, C, ps->Lq, ps->head_mod, ps->head_mod + C, 2 * C, ps->slot_map, w->eps, ps->xm, stream))) return rc;
    sdvar_gemm_epilogue f{};
    f.epilogue = SDVAR_EPI_F32;
    f.bias = w->b_head;
    f.out_f32 = ps->logits; f.ldo = w->V;
    if ((rc = sdvar_gemm_bf16(ps->xm, C, w->w_head, C, M, w->V, C, &f, stream))) return rc;
  }
  return SDVAR_OK;
}
