// Coarse entry points of the two transformer stacks on the hot path, so that a non-Python host can run "one ViT forward" / "one prefill"
// with ONE call each (SURVEY.md 8(b) proposal: d3d_vit_l14_336_forward, d3d_phi3_prefill):
//   d3d_vit_forward  : CLIPEncoder.forward + VisionTransformer.forward (resnet_encoders.py:273-284, clip/model.py:219-238), also the LLaVA
//                      tower's hidden_states[-2] (Policy_Dynam3D_VLN.py:441-452) with n_layers_run = n_layers - 1, project = 0;
//   d3d_phi3_prefill : the prefill of llava.generate over packed inputs_embeds (POL:456-463) -> last-token logits, optionally keeping every
//                      layer's packed QKV matrix as the KV cache of the greedy decode.
// Host code only: every layer is the same sequence of library kernels the Python engines issue (clip_vit.py / phi3.py keep their layer loops
// for the per-stage profile and the precise mode), so results are bit-identical to the per-kernel path.
#include <math.h>

#include "common.cuh"

namespace {

int attention_auto(const void* qkv, int64_t ld, int64_t n_rows, void* out, int64_t ldo, const int* cu, int n_seq, int max_len, int H, int Dh,
                   int causal, int kind, void* stream) {
  const float scale = (float)(1.0 / sqrt((double)Dh));  // the double expression the Python engines (and HF) evaluate, rounded once
  if ((Dh == 64 || Dh == 96) && max_len >= 256) return d3d_attention_tc(qkv, ld, n_rows, out, ldo, cu, n_seq, max_len, H, Dh, causal, kind, scale, stream);
  if ((Dh == 64 || Dh == 96) && max_len >= 64) return d3d_attention_mma(qkv, ld, out, ldo, cu, n_seq, max_len, H, Dh, causal, kind, scale, stream);
  return d3d_attention_simt(qkv, ld, out, ldo, cu, n_seq, max_len, H, Dh, causal, kind, scale, stream);
}

int gemm(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K, int in_kind, int out_kind, const float* bias,
         int act, const float* residual, int64_t ldres, void* stream, bool skinny = false) {
  d3d_gemm_args a;
  a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.in_kind = in_kind; a.out_kind = out_kind;
  a.bias = bias; a.act = act; a.residual = residual; a.ldres = ldres;
  return skinny ? d3d_gemm_skinny(&a, stream) : d3d_gemm(&a, stream);
}

}  // namespace

extern "C" int d3d_vit_forward(const d3d_vit_model* m, const uint8_t* img, int N, int Hin, int Win, int n_layers_run, int project,
                               const d3d_vit_scratch* s, void* stream) {
  D3D_REQUIRE(m && img && s && N > 0, "args");
  D3D_REQUIRE(n_layers_run >= 0 && n_layers_run <= m->n_layers, "n_layers_run");
  D3D_REQUIRE(!project || (m->ln_post_g && m->proj), "this tower has no ln_post / proj (HF vision tower): call with project = 0");
  const int W = m->width, T = N * m->tokens, g2 = m->tokens - 1, kind = m->kind, Dh = W / m->n_heads;
  const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f}, stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
  D3D_TRY(d3d_preprocess_im2col(img, N, Hin, Win, m->resolution, m->patch, mean, stdv, s->cols, m->kpad, kind, stream));
  D3D_TRY(gemm(s->cols, m->kpad, m->conv_w, m->kpad, s->conv, W, N * g2, W, m->kpad, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, nullptr, 0, stream));
  D3D_TRY(d3d_vit_embed_ln(s->conv, m->cls, m->pos, m->ln_pre_g, m->ln_pre_b, 1e-5f, N, m->tokens, W, s->X, stream));
  for (int l = 0; l < n_layers_run; ++l) {
    const d3d_vit_layer& p = m->layers[l];
    D3D_TRY(d3d_layernorm(s->X, W, nullptr, p.ln1_g, p.ln1_b, 1e-5f, T, W, D3D_ACT_NONE, nullptr, 0, s->A16, W, kind, stream));
    D3D_TRY(gemm(s->A16, W, p.w_qkv, W, s->qkv, 3 * W, T, 3 * W, W, kind, kind, p.b_qkv, D3D_ACT_NONE, nullptr, 0, stream));
    D3D_TRY(attention_auto(s->qkv, 3 * W, T, s->att, W, s->cu, N, m->tokens, m->n_heads, Dh, 0, kind, stream));
    D3D_TRY(gemm(s->att, W, p.w_o, W, s->X, W, T, W, W, kind, D3D_OUT_F32, p.b_o, D3D_ACT_NONE, s->X, W, stream));
    D3D_TRY(d3d_layernorm(s->X, W, nullptr, p.ln2_g, p.ln2_b, 1e-5f, T, W, D3D_ACT_NONE, nullptr, 0, s->A16, W, kind, stream));
    D3D_TRY(gemm(s->A16, W, p.w_fc, W, s->h, 4 * W, T, 4 * W, W, kind, kind, p.b_fc, D3D_ACT_QUICK_GELU, nullptr, 0, stream));
    D3D_TRY(gemm(s->h, 4 * W, p.w_pr, 4 * W, s->X, W, T, W, 4 * W, kind, D3D_OUT_F32, p.b_pr, D3D_ACT_NONE, s->X, W, stream));
  }
  if (!project) return 0;  // fp32 hidden state [N, tokens, width] in s->X
  D3D_TRY(d3d_layernorm(s->X, W, nullptr, m->ln_post_g, m->ln_post_b, 1e-5f, T, W, D3D_ACT_NONE, nullptr, 0, s->A16, W, kind, stream));
  return gemm(s->A16, W, m->proj, W, s->out, m->out_dim, T, m->out_dim, W, kind, kind, nullptr, D3D_ACT_NONE, nullptr, 0, stream);
}

extern "C" int d3d_phi3_prefill(const d3d_lm_model* m, float* X, int T, const int* cu_seqlens, const int* positions, int n_seq, int max_len,
                                const int* last_rows, const float* inv_freq, void* const* qkv_layers_h, int64_t ld_qkv,
                                const d3d_lm_scratch* s, int trim_last_layer, float* logits, void* stream) {
  D3D_REQUIRE(m && X && cu_seqlens && positions && last_rows && inv_freq && s && logits && T > 0 && n_seq > 0, "args");
  const int Hd = m->hidden, F = m->ffn, kind = m->kind, H = m->n_heads, Dh = m->head_dim;
  D3D_TRY(d3d_rope_table(positions, inv_freq, T, Dh, s->rope_tab, stream));  // cos/sin per token, shared by all layers
  // Only each sequence's LAST token feeds the logits (POL:463 prefill): in the final layer everything after the attention is row-wise, so
  // o_proj / MLP run on the n_seq last rows only (K and V of all rows are still produced: they are the KV cache).
  const bool trim = trim_last_layer && n_seq <= 16 && m->n_layers > 0;
  for (int l = 0; l < m->n_layers; ++l) {
    const d3d_lm_layer& p = m->layers[l];
    void* qkv = qkv_layers_h ? qkv_layers_h[l] : s->qkv;
    const int64_t ldq = qkv_layers_h ? ld_qkv : 3 * (int64_t)Hd;
    D3D_TRY(d3d_rmsnorm(X, Hd, nullptr, p.rms1, m->eps, T, Hd, nullptr, 0, s->A16, Hd, kind, stream));
    D3D_TRY(gemm(s->A16, Hd, p.w_qkv, Hd, qkv, ldq, T, 3 * Hd, Hd, kind, kind, nullptr, D3D_ACT_NONE, nullptr, 0, stream));
    D3D_TRY(d3d_rope_apply(qkv, ldq, s->rope_tab, T, H, Dh, kind, stream));
    D3D_TRY(attention_auto(qkv, ldq, T, s->att, Hd, cu_seqlens, n_seq, max_len, H, Dh, 1, kind, stream));
    if (trim && l == m->n_layers - 1) {
      D3D_TRY(d3d_gather_rows16(s->att, Hd, last_rows, s->att_last, Hd, n_seq, Hd, stream));
      D3D_TRY(d3d_scatter_rows(X, Hd, last_rows, s->x_last, Hd, nullptr, n_seq, Hd, stream));
      D3D_TRY(gemm(s->att_last, Hd, p.w_o, Hd, s->x_last, Hd, n_seq, Hd, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, s->x_last, Hd, stream, true));
      D3D_TRY(d3d_rmsnorm(s->x_last, Hd, nullptr, p.rms2, m->eps, n_seq, Hd, nullptr, 0, s->last16, Hd, kind, stream));
      D3D_TRY(gemm(s->last16, Hd, p.w_gu, Hd, s->h, F, n_seq, 2 * F, Hd, kind, kind, nullptr, D3D_ACT_SWIGLU, nullptr, 0, stream, true));
      D3D_TRY(gemm(s->h, F, p.w_down, F, s->x_last, Hd, n_seq, Hd, F, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, s->x_last, Hd, stream, true));
      D3D_TRY(d3d_rmsnorm(s->x_last, Hd, nullptr, m->norm, m->eps, n_seq, Hd, nullptr, 0, s->last16, Hd, kind, stream));
      return gemm(s->last16, Hd, m->lm_head, Hd, logits, m->vocab, n_seq, m->vocab, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, nullptr, 0, stream);
    }
    D3D_TRY(gemm(s->att, Hd, p.w_o, Hd, X, Hd, T, Hd, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, X, Hd, stream));
    D3D_TRY(d3d_rmsnorm(X, Hd, nullptr, p.rms2, m->eps, T, Hd, nullptr, 0, s->A16, Hd, kind, stream));
    D3D_TRY(gemm(s->A16, Hd, p.w_gu, Hd, s->h, F, T, 2 * F, Hd, kind, kind, nullptr, D3D_ACT_SWIGLU, nullptr, 0, stream));
    D3D_TRY(gemm(s->h, F, p.w_down, F, X, Hd, T, Hd, F, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, X, Hd, stream));
  }
  D3D_TRY(d3d_rmsnorm(X, Hd, last_rows, m->norm, m->eps, n_seq, Hd, nullptr, 0, s->last16, Hd, kind, stream));
  return gemm(s->last16, Hd, m->lm_head, Hd, logits, m->vocab, n_seq, m->vocab, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, nullptr, 0, stream);
}

// Shared tail of a layer for compact rows: o_proj (+residual) -> rmsnorm -> gate/up + SwiGLU -> down (+residual)
static int lm_layer_tail(const d3d_lm_model* m, const d3d_lm_layer& p, float* X, int T, const d3d_lm_scratch* s, void* stream) {
  const int Hd = m->hidden, F = m->ffn, kind = m->kind;
  D3D_TRY(gemm(s->att, Hd, p.w_o, Hd, X, Hd, T, Hd, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, X, Hd, stream));
  D3D_TRY(d3d_rmsnorm(X, Hd, nullptr, p.rms2, m->eps, T, Hd, nullptr, 0, s->A16, Hd, kind, stream));
  D3D_TRY(gemm(s->A16, Hd, p.w_gu, Hd, s->h, F, T, 2 * F, Hd, kind, kind, nullptr, D3D_ACT_SWIGLU, nullptr, 0, stream));
  return gemm(s->h, F, p.w_down, F, X, Hd, T, Hd, F, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, X, Hd, stream);
}

extern "C" int d3d_phi3_prefill_chunk(const d3d_lm_model* m, float* X, int T, int n_seq, int max_len, const d3d_lm_chunk* c, void* const* cache_layers_h,
                                      int64_t ld_cache, int64_t cache_rows, void* att_cache, const float* inv_freq, const d3d_lm_scratch* s,
                                      const int* last_rows, int trim_last_layer, float* logits, void* stream) {
  D3D_REQUIRE(m && X && c && cache_layers_h && att_cache && inv_freq && s && T > 0 && n_seq > 0, "args");
  D3D_REQUIRE(c->seq_start && c->seq_len && c->rows && c->positions, "chunk tables");
  D3D_REQUIRE(!last_rows || logits, "logits buffer");
  const int Hd = m->hidden, F = m->ffn, kind = m->kind, H = m->n_heads, Dh = m->head_dim;
  D3D_REQUIRE(Dh == 64 || Dh == 96, "the chunked prefill runs on the tcgen05 attention (head_dim 64 / 96)");
  const float scale = (float)(1.0 / sqrt((double)Dh));
  D3D_TRY(d3d_rope_table(c->positions, inv_freq, T, Dh, s->rope_tab, stream));
  const bool trim = last_rows && trim_last_layer && n_seq <= 16 && m->n_layers > 0;
  for (int l = 0; l < m->n_layers; ++l) {
    const d3d_lm_layer& p = m->layers[l];
    void* cache = cache_layers_h[l];
    D3D_TRY(d3d_rmsnorm(X, Hd, nullptr, p.rms1, m->eps, T, Hd, nullptr, 0, s->A16, Hd, kind, stream));
    D3D_TRY(gemm(s->A16, Hd, p.w_qkv, Hd, s->qkv, 3 * Hd, T, 3 * Hd, Hd, kind, kind, nullptr, D3D_ACT_NONE, nullptr, 0, stream));
    D3D_TRY(d3d_rope_apply(s->qkv, 3 * Hd, s->rope_tab, T, H, Dh, kind, stream));
    D3D_TRY(d3d_scatter_rows16(s->qkv, 3 * Hd, cache, ld_cache, c->rows, T, 3 * Hd, stream));
    D3D_TRY(d3d_attention_tc_ex(cache, ld_cache, cache_rows, att_cache, Hd, c->seq_start, c->seq_len, n_seq, max_len, c->q_tile_begin, c->q_tile_end, H,
                                Dh, 1, kind, scale, stream));
    if (trim && l == m->n_layers - 1) {
      // last-token rows only: att_last[i] = att_cache[rows[last_rows[i]]] needs a two-level index; gather the chunk's rows first
      D3D_TRY(d3d_gather_rows16(att_cache, Hd, c->rows, s->att, Hd, T, Hd, stream));
      D3D_TRY(d3d_gather_rows16(s->att, Hd, last_rows, s->att_last, Hd, n_seq, Hd, stream));
      D3D_TRY(d3d_scatter_rows(X, Hd, last_rows, s->x_last, Hd, nullptr, n_seq, Hd, stream));
      D3D_TRY(gemm(s->att_last, Hd, p.w_o, Hd, s->x_last, Hd, n_seq, Hd, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, s->x_last, Hd, stream, true));
      D3D_TRY(d3d_rmsnorm(s->x_last, Hd, nullptr, p.rms2, m->eps, n_seq, Hd, nullptr, 0, s->last16, Hd, kind, stream));
      D3D_TRY(gemm(s->last16, Hd, p.w_gu, Hd, s->h, F, n_seq, 2 * F, Hd, kind, kind, nullptr, D3D_ACT_SWIGLU, nullptr, 0, stream, true));
      D3D_TRY(gemm(s->h, F, p.w_down, F, s->x_last, Hd, n_seq, Hd, F, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, s->x_last, Hd, stream, true));
      D3D_TRY(d3d_rmsnorm(s->x_last, Hd, nullptr, m->norm, m->eps, n_seq, Hd, nullptr, 0, s->last16, Hd, kind, stream));
      return gemm(s->last16, Hd, m->lm_head, Hd, logits, m->vocab, n_seq, m->vocab, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, nullptr, 0, stream);
    }
    D3D_TRY(d3d_gather_rows16(att_cache, Hd, c->rows, s->att, Hd, T, Hd, stream));
    D3D_TRY(lm_layer_tail(m, p, X, T, s, stream));
  }
  if (!last_rows) return 0;
  D3D_TRY(d3d_rmsnorm(X, Hd, last_rows, m->norm, m->eps, n_seq, Hd, nullptr, 0, s->last16, Hd, kind, stream));
  return gemm(s->last16, Hd, m->lm_head, Hd, logits, m->vocab, n_seq, m->vocab, Hd, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, nullptr, 0, stream);
}
