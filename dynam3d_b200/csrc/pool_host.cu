// Host-side orchestration (inside the library) of the packed pooling pass, so that the Python layer makes ONE call per
// (view, level) instead of ~40: position-feature rows -> MLP (Linear, LayerNorm, GELU, Linear) -> + content feature /
// aggregate token -> 2-layer post-norm TransformerEncoder -> final LayerNorm on the first token of every sequence.
// Reference: FF:580-597, 662-688, 717-730, 743-756 (one encoder call per segment / zone there).
#include "common.cuh"

// in-library (C++ linkage): pool_kernels.cu
int d3d_pool_assemble_cast(const float* emb, const int64_t* seq_fts, int fts_is_f32, const int* tok_seq, const int* tok_src, const float* agg,
                           int T, int D, float* X, void* X16, int kind, void* stream);

extern "C" {
int d3d_gemm(const d3d_gemm_args* args_h, void* stream);
int d3d_layernorm(const float* x, int64_t ldx, const int* row_index, const float* gamma, const float* beta, float eps, int T, int D, int act,
                  float* out32, int64_t ld32, void* out16, int64_t ld16, int kind16, void* stream);
int d3d_cast16(const float* in, int64_t ldi, void* out, int64_t ldo, int T, int D, int kind, void* stream);
int d3d_attention_simt(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H, int Dh,
                       int causal, int kind, float scale, void* stream);
int d3d_attention_mma(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H, int Dh,
                      int causal, int kind, float scale, void* stream);
int d3d_pool_features(const int64_t* seq_xyz, const int64_t* seq_dir, const int64_t* seq_scale, const float* centre, const int* tok_seq,
                      const int* tok_src, int T, int mode, void* out16, int kind, void* stream);
int d3d_pool_assemble(const float* emb, const int64_t* seq_fts, int fts_is_f32, const int* tok_seq, const int* tok_src, const float* agg, int T,
                      int D, float* X, void* stream);
}

namespace {

int gemm(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N, int K, int kind, int out_kind,
         const float* bias, int act, const float* res, int64_t ldres, void* st) {
  d3d_gemm_args a;
  a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
  a.in_kind = kind; a.out_kind = out_kind; a.bias = bias; a.act = act; a.residual = res; a.ldres = ldres;
  return d3d_gemm(&a, st);
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t d3d_pool_workspace_bytes(int T, int d_model, int d_hidden_mlp) {
  const size_t t = (size_t)T;
  size_t b = 0;
  b += al(t * 8 * 2);                        // A0 feature rows
  b += al(t * (size_t)d_hidden_mlp * 4);     // MLP hidden fp32
  b += al(t * (size_t)d_hidden_mlp * 2);     // MLP hidden 16-bit / A16
  b += al(t * (size_t)d_model * 4);          // emb
  b += al(t * (size_t)d_model * 4);          // X
  b += al(t * (size_t)d_model * 3 * 2);      // qkv
  b += al(t * (size_t)d_model * 2);          // att
  b += al(t * (size_t)d_model * 4 * 2);      // ffn hidden
  return b + 4096;
}

extern "C" int d3d_mlp_ln_gelu(const d3d_mlp* m_h, const void* A0, int64_t lda, int T, void* hidden32, void* hidden16, float* out, int64_t ldo,
                               void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(m_h && A0 && hidden32 && hidden16 && out, "args");
  const d3d_mlp& m = *m_h;
  D3D_TRY(gemm(A0, lda, m.w0, m.k_pad, hidden32, m.d_hidden, T, m.d_hidden, m.k_pad, m.kind, D3D_OUT_F32, m.b0, D3D_ACT_NONE, nullptr, 0, stream));
  D3D_TRY(d3d_layernorm((const float*)hidden32, m.d_hidden, nullptr, m.ln_g, m.ln_b, 1e-5f, T, m.d_hidden, D3D_ACT_GELU, nullptr, 0, hidden16,
                        m.d_hidden, m.kind, stream));
  D3D_TRY(gemm(hidden16, m.d_hidden, m.w3, m.d_hidden, out, ldo, T, m.d_out, m.d_hidden, m.kind, D3D_OUT_F32, m.b3, D3D_ACT_NONE, nullptr, 0, stream));
  return 0;
}

extern "C" int d3d_pool_tokens(const d3d_pool_level* lvl_h, const int64_t* seq_ptrs /*[4][n_seq]: xyz, dir, scale, fts*/, const float* centre,
                               const int* tok_seq, const int* tok_src, const int* cu_seqlens, int T, int n_seq, int max_len, int mode,
                               int fts_is_f32, void* workspace, size_t workspace_bytes, float* out, void* stream) {
  if (n_seq == 0 || T == 0) return 0;
  D3D_REQUIRE(lvl_h && seq_ptrs && centre && tok_seq && tok_src && cu_seqlens && workspace && out, "args");
  const d3d_pool_level& L = *lvl_h;
  const int D = L.d_model, kind = L.mlp.kind;
  D3D_REQUIRE(L.mlp.d_out == D && D % L.n_head == 0, "level dims");
  if (workspace_bytes < d3d_pool_workspace_bytes(T, D, L.mlp.d_hidden)) {
    d3d_set_error("pool workspace too small: %zu < %zu", workspace_bytes, d3d_pool_workspace_bytes(T, D, L.mlp.d_hidden));
    return D3D_ENOMEM;
  }
  char* p = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  auto take = [&](size_t bytes) { char* r = p; p += al(bytes); return (void*)r; };
  const size_t t = (size_t)T;
  void* A0 = take(t * 8 * 2);
  void* h32 = take(t * (size_t)L.mlp.d_hidden * 4);
  void* h16 = take(t * (size_t)L.mlp.d_hidden * 2);
  float* emb = (float*)take(t * D * 4);
  float* X = (float*)take(t * D * 4);
  void* qkv = take(t * D * 3 * 2);
  void* att = take(t * D * 2);
  void* ffn = take(t * D * 4 * 2);
  void* A16 = h16;  // re-used once the MLP is done (d_hidden >= d_model)
  D3D_REQUIRE(L.mlp.d_hidden >= D, "mlp hidden narrower than d_model");

  D3D_TRY(d3d_pool_features(seq_ptrs, seq_ptrs + n_seq, seq_ptrs + 2 * (size_t)n_seq, centre, tok_seq, tok_src, T, mode, A0, kind, stream));
  D3D_TRY(d3d_mlp_ln_gelu(&L.mlp, A0, 8, T, h32, h16, emb, D, stream));
  D3D_TRY(d3d_pool_assemble_cast(emb, seq_ptrs + 3 * (size_t)n_seq, fts_is_f32, tok_seq, tok_src, L.agg, T, D, X, A16, kind, stream));
  const int Dh = D / L.n_head;
  const float scale = 1.0f / sqrtf((float)Dh);
  for (int l = 0; l < L.n_layers; ++l) {
    const d3d_encoder_layer& e = L.layers[l];
    D3D_TRY(gemm(A16, D, e.w_in, D, qkv, 3 * D, T, 3 * D, D, kind, kind, e.b_in, D3D_ACT_NONE, nullptr, 0, stream));
    if (max_len >= 64 && Dh == 64)  // mixed lengths: short sequences one warp per (sequence, head), long ones on the 128-row kernel
      D3D_TRY(d3d_attention_mixed(qkv, 3 * D, att, D, cu_seqlens, n_seq, max_len, L.n_head, Dh, kind, scale, stream));
    else if (max_len >= 64)
      D3D_TRY(d3d_attention_mma(qkv, 3 * D, att, D, cu_seqlens, n_seq, max_len, L.n_head, Dh, 0, kind, scale, stream));
    else
      D3D_TRY(d3d_attention_simt(qkv, 3 * D, att, D, cu_seqlens, n_seq, max_len, L.n_head, Dh, 0, kind, scale, stream));
    D3D_TRY(gemm(att, D, e.w_out, D, X, D, T, D, D, kind, D3D_OUT_F32, e.b_out, D3D_ACT_NONE, X, D, stream));
    D3D_TRY(d3d_layernorm(X, D, nullptr, e.n1_g, e.n1_b, 1e-5f, T, D, D3D_ACT_NONE, X, D, A16, D, kind, stream));
    D3D_TRY(gemm(A16, D, e.w1, D, ffn, 4 * D, T, 4 * D, D, kind, kind, e.b1, D3D_ACT_GELU, nullptr, 0, stream));
    D3D_TRY(gemm(ffn, 4 * D, e.w2, 4 * D, X, D, T, D, 4 * D, kind, D3D_OUT_F32, e.b2, D3D_ACT_NONE, X, D, stream));
    D3D_TRY(d3d_layernorm(X, D, nullptr, e.n2_g, e.n2_b, 1e-5f, T, D, D3D_ACT_NONE, X, D, A16, D, kind, stream));
  }
  D3D_TRY(d3d_layernorm(X, D, cu_seqlens, L.norm_g, L.norm_b, L.norm_eps, n_seq, D, D3D_ACT_NONE, out, D, nullptr, 0, kind, stream));
  return 0;
}
