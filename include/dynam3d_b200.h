/* libdynam3d_b200.so -- C ABI of the B200-native Dynam3D per-step hot path.
 *
 * The reference (MrZihan/Dynam3D) is 100% Python and has NO FFI boundary of its own; its native code lives in
 * un-vendored wheels (torch_kdtree, tinycudann, open3d, cuBLAS via torch).  Each entry below names the reference
 * Python call site it replaces (file:line under /root/reference):
 *   FF   = Dynam3D_VLN/vlnce_baselines/models/feature_fields.py
 *   POL  = Dynam3D_VLN/vlnce_baselines/models/Policy_Dynam3D_VLN.py
 *   ENC  = Dynam3D_VLN/vlnce_baselines/models/encoders/resnet_encoders.py
 *   CLIPM= Dynam3D_VLN/vlnce_baselines/models/encoders/clip/model.py
 *   PFF  = Dynam3D_Pretrain/src_3dff/models/feature_fields.py
 *
 * Conventions: plain pointers and sizes only (no torch types).  Pointers are DEVICE pointers unless the
 * parameter name ends in `_h` (host).  Every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 * returns 0 on success or a negative D3D_E* code; `d3d_last_error()` returns the message of the last failure on
 * the calling thread.  No call allocates device memory except d3d_workspace_* (explicit) -- scratch is passed in.
 * One host thread per stream; kernels are sm_100a only and the library refuses to run elsewhere (no fallback).
 */
#ifndef DYNAM3D_B200_H
#define DYNAM3D_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3D_OK 0
#define D3D_EINVAL (-1)   /* bad argument / unsupported shape */
#define D3D_ECUDA (-2)    /* CUDA runtime / driver error */
#define D3D_EARCH (-3)    /* device is not sm_100 */
#define D3D_ENOMEM (-4)   /* workspace too small */

/* 16-bit storage kinds for tensor-core operands */
#define D3D_F16 0
#define D3D_BF16 1
/* output kinds */
#define D3D_OUT_F32 2

/* GEMM epilogue activations */
#define D3D_ACT_NONE 0
#define D3D_ACT_QUICK_GELU 1 /* x*sigmoid(1.702x), CLIPM:162-164 */
#define D3D_ACT_GELU 2       /* exact erf GELU, nn.GELU() */
#define D3D_ACT_SILU 3
#define D3D_ACT_SWIGLU 4     /* out[:, j] = silu(acc[:, 2j]) * acc[:, 2j+1] on row-interleaved gate/up weights */
#define D3D_ACT_LEAKY_RELU 5 /* slope 0.01: tinycudann CutlassMLP "LeakyReLU" (PFF:221-243) */

const char* d3d_last_error(void);
int d3d_version(void);
/* 0 when device `dev` is an sm_100 part, D3D_EARCH otherwise */
int d3d_check_device(int dev);
int d3d_sm_count(void);
/* number of kernels this library has launched in the calling process (bench.py reports the difference over the timed region) */
long long d3d_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core GEMM (tcgen05.mma + TMEM accumulators, TMA-fed, persistent):  C = epi(A @ W^T)
 *   A [M,K] row-major 16-bit (lda elements), W [N,K] row-major 16-bit (nn.Linear layout, ldw), fp32 accumulate.
 *   epi: (+bias[n]) -> act -> (+residual[m,n] fp32) -> store as out_kind (D3D_F16 / D3D_BF16 / D3D_OUT_F32).
 *   Replaces every nn.Linear / nn.MultiheadAttention projection / HF Linear on the path
 *   (CLIPM:167-188,234-236; FF:134-161; POL:83-111; HF LlamaMLP/LlamaAttention via POL:123-127), i.e. the
 *   cuBLAS calls of torch 1.13.  K*2 and ld*2 bytes must be multiples of 16; pointers 16-byte aligned.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const void* A; int64_t lda;
  const void* W; int64_t ldw;
  void* C; int64_t ldc;
  int M, N, K;
  int in_kind;            /* D3D_F16 | D3D_BF16 (A and W) */
  int out_kind;           /* D3D_F16 | D3D_BF16 | D3D_OUT_F32 */
  const float* bias;      /* [N] or NULL */
  int act;                /* D3D_ACT_* */
  const float* residual;  /* [M, ldres] fp32 or NULL; may alias C when out_kind == D3D_OUT_F32 */
  int64_t ldres;
} d3d_gemm_args;
int d3d_gemm(const d3d_gemm_args* args_h, void* stream);
/* Large problems run on CTA pairs (tcgen05.mma.cta_group::2, 256x256 tiles per 2-CTA cluster).  mode: 1 = on (default),
 * 0 = single-CTA kernels only (A/B comparisons in tools/gemm_bench.py), -1 = re-read D3D_GEMM_PAIR from the environment. */
int d3d_gemm_set_pair_mode(int mode);

/* Plain CUDA-core reference GEMM with the same contract (debug / self-check only, slow). */
int d3d_gemm_simt(const d3d_gemm_args* args_h, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry of the 3D-token path (HBM-bound fp32; bit-exact against oracle/geometry.py).
 * ------------------------------------------------------------------------------------------------ */

/* preprocess_depth on full-resolution images (POL:171-186, call site POL:350): zeros -> column max, then
 * (lo*100 + d*(hi-lo)*100)/100 metres.  obs/out [n_img,H,W] fp32. */
int d3d_depth_preprocess(const float* obs, float* out, int n_img, int H, int W, float lo, float hi, void* stream);

/* cv2.resize(INTER_NEAREST) to the gh x gw patch grid + preprocess_depth (POL:336-341).  obs [n_img,H,W] fp32,
 * out [batch*views, gh*gw].  row_idx_h/col_idx_h: host tables of source indices (cv2's floor(x*src/dst)).
 * literal_q1 != 0 reproduces the reference's `observations['depth'][b][i]` indexing:
 * out[b,i,r,c] = obs[b, i, row_idx[r]] (image row i), in which case col_idx_h is ignored. */
int d3d_depth_patch_grid(const float* obs, float* out, int batch, int views, int H, int W, int gh, int gw, int literal_q1,
                         const int* row_idx_h, const int* col_idx_h, float lo, float hi, void* stream);

/* project_depth_to_3d_habitat + world offset (FF:276-293, 548-554).  depth [n,W*H] metres; pose [n,6] =
 * (x, y, z internal frame, cos(theta), sin(theta), theta) fp32 with theta = ix*(-pi/6)+heading; tan_x_h[W],
 * tan_z_h[H], neg_atan_x_h[W]: host tables built with the reference's expressions (FF:283-286); tan_h = tan(hfov/2).
 * Outputs xyz [n,W*H,3], dir [n,W*H] (mod 2pi), scale [n,W*H]. */
int d3d_unproject_habitat(const float* depth, const float* pose, int n_units, int W, int H, const float* tan_x_h,
                          const float* tan_z_h, const float* neg_atan_x_h, float tan_h, float* xyz, float* dir, float* scale,
                          void* stream);

/* get_patch_3d_info (FF:296-326): out5 [5,n,W*H] = rel_x, rel_y, rel_z, direction mod 2pi, scale. */
int d3d_patch_3d_info(const float* depth, int n_units, int W, int H, const float* tan_x_h, const float* tan_z_h,
                      const float* neg_atan_x_h, float tan_h, float* out5, void* stream);

/* delete_old_features_from_camera_frustum, numeric part (FF:88-115, 347-360) for all views of a step at once:
 * a stored patch is culled if for ANY view it projects inside the HxW image, lies in [near, far] and in front of
 * the observed depth + eps.  Culled rows are tombstoned in place (xyz=-10000, dir=scale=0, fts16 row=0);
 * mask[n] (1 = culled now) and *n_deleted are written for the host bookkeeping (FF:362-393).
 * depth [n_views,H,W] metres; cam [n_views,5] = (x, y, z internal, cos(-heading), sin(-heading)). */
int d3d_frustum_cull(float* xyz, float* dir, float* scale, void* fts16, int n_patches, int fts_dim, const float* depth,
                     int n_views, int H, int W, const float* cam, float fx, float fy, float cx, float cy, float near_,
                     float far_, float eps, uint8_t* mask, int* n_deleted, void* stream);

/* The same cull for ALL episodes of a rank in one launch, with device-side compaction of the culled row indices so that the host
 * bookkeeping (d3d_ffh_cull_list) touches only what was deleted -- the long-horizon form (147 k stored patches per episode, FF:329-396).
 * jobs [n_jobs] (device): per-episode pool base addresses and stored-patch count; depth [n_jobs, n_views, H, W]; cam [n_jobs, n_views, 5];
 * del_idx [n_jobs, del_cap] int32 receives the culled rows of each episode in no particular order, n_del [n_jobs] their count
 * (entries beyond del_cap are dropped: size del_cap >= the largest n_patches to make that impossible). */
typedef struct { uint64_t xyz, dir, scale, fts16; int32_t n_patches; int32_t pad; } d3d_cull_job;
int d3d_frustum_cull_batched(const d3d_cull_job* jobs, int n_jobs, int max_patches, int fts_dim, const float* depth, int n_views, int H, int W,
                             const float* cam, float fx, float fy, float cx, float cy, float near_, float far_, float eps, int* del_idx,
                             int del_cap, int* n_del, void* stream);

/* Posed-dataset form of the cull (get_frustum_mask FF:64-84, call site FF:343-344; z-test FF:349-353): cam25 [n_views,25] fp32 =
 * world->camera view matrix (4x4 row-major) followed by the intrinsics (3x3 row-major).  Same tombstoning / outputs as above. */
int d3d_frustum_cull_matrix(float* xyz, float* dir, float* scale, void* fts16, int n_patches, int fts_dim, const float* depth,
                            int n_views, int H, int W, const float* cam25, float near_, float far_, float eps, uint8_t* mask,
                            int* n_deleted, void* stream);

/* Posed-dataset unprojection (a4'): project_depth_to_3d (FF:50-60: open3d create_from_depth_image on the uint16 depth + nearest resize to
 * the gh x gw grid, replacing the 8 joblib/open3d CPU threads of FF:130,518), R @ p + T, get_heading_angle and the patch scale (FF:250-259,
 * 536-546).  depth [n_views,H,W] uint16; view_params [n_views,16] double (device) = fx, fy, cx, cy, R[9] row-major, T[3];
 * row_idx_h / col_idx_h: host tables of F.interpolate(mode='nearest') source indices; tan_abs = |tan(rel_direction[0][-1])| of get_rays.
 * *n_invalid counts pixels open3d would drop (z >= depth_trunc): the reference raises there, the caller should too. */
int d3d_unproject_pinhole(const uint16_t* depth, int n_views, int H, int W, const double* view_params, int gh, int gw,
                          const int* row_idx_h, const int* col_idx_h, float depth_scale, float depth_trunc, float tan_abs, float* xyz,
                          float* dir, float* scale, int* n_invalid, void* stream);

/* Exact K-NN in 3-D, replaces torch_kdtree build_kd_tree + query (FF:246,606,610; PFF:364,540,584):
 * squared L2 ((dx*dx+dy*dy)+dz*dz in fp32), ascending, lowest index on ties.  k <= 8, n_ref >= k.
 * refs [n_ref,3], queries [n_q,3] -> out_d2 [n_q,k] fp32, out_idx [n_q,k] int32. */
int d3d_knn3d(const float* refs, int n_ref, const float* queries, int n_q, int k, float* out_d2, int* out_idx, void* stream);

/* centroid of gathered points per sequence (FF:583, 663, 715): fp64 accumulate, one rounding; empty -> NaN.
 * member[cu_seqlens[s] .. cu_seqlens[s+1]) are row indices into xyz [*,3]; out [n_seq,3]. */
int d3d_seq_centroid(const float* xyz, const int* member, const int* cu_seqlens, int n_seq, float* out, void* stream);

/* get_environment_features for one episode and one token level (FF:818-862): gather rows `ids` (dict order) from
 * pos [*,3] / fts [*,width] fp32, move to the agent frame, keep rows with |rel| <= radius, order preserved.
 * agent [5] = (x, y, z internal, cos(-heading), sin(-heading)).  out_rel [n_ids,3], out_fts [n_ids,width], *out_count. */
int d3d_env_export(const float* pos, const float* fts, const int* ids, int n_ids, const float* agent, float radius, int width,
                   float* out_rel, float* out_fts, int* out_count, void* stream);
/* All (episode, token kind) exports of a step in ONE launch (one block per job).  jobs: device array of 72-byte records
 * { const float* pos; const float* fts; int64 ids_off; float* out_rel; float* out_fts; float agent[5]; float radius; int n_ids; int pad; };
 * ids_all: the jobs' id lists back to back (job i reads ids_all[ids_off .. ids_off + n_ids)); out_count [n_jobs]. */
int d3d_env_export_batched(const void* jobs, const int* ids_all, int n_jobs, int width, int* out_count, void* stream);

/* FastSAM post-processing (FF:411-422): masks [n_img, M, H, W] u8 (0/1, in FastSAM's order) -> dense segment labels
 * out [n_img, gh*gw] int64 in 0..G-1 and n_seg[n_img]: later masks overwrite earlier ones, uncovered pixels take label 0,
 * torch 'nearest' resize through the host index tables (floor(dst * in/out)), labels renumbered in ascending order. */
int d3d_segm_relabel(const uint8_t* masks, int n_img, int M, int H, int W, int gh, int gw, const int* row_idx_h,
                     const int* col_idx_h, int64_t* out, int* n_seg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation / elementwise / gather kernels between the GEMMs (fp32 math; 16-bit outputs feed GEMM A operands).
 * Row widths D must be one of 128, 256, 512, 768, 1024, 3072, 4096.
 * ------------------------------------------------------------------------------------------------ */

/* torch.nn.LayerNorm (CLIPM:153-159; FF:141,146,151,155,159; POL:85-109): y = act(LN(x[row_index?row_index[r]:r])*gamma+beta),
 * act in {D3D_ACT_NONE, D3D_ACT_GELU}; writes out32 [T,ld32] and/or out16 [T,ld16] (either may be NULL). */
int d3d_layernorm(const float* x, int64_t ldx, const int* row_index, const float* gamma, const float* beta, float eps, int T,
                  int D, int act, float* out32, int64_t ld32, void* out16, int64_t ld16, int kind16, void* stream);

/* HF LlamaRMSNorm / Phi3RMSNorm (llava language model, POL:123-127): y = w * x * rsqrt(mean(x^2)+eps). */
int d3d_rmsnorm(const float* x, int64_t ldx, const int* row_index, const float* w, float eps, int T, int D, float* out32,
                int64_t ld32, void* out16, int64_t ld16, int kind16, void* stream);

/* HF apply_rotary_pos_emb (rotate_half form) in place on q and k of a packed [T, 3*H*Dh] 16-bit QKV buffer;
 * pos [T] int32 token positions, inv_freq [Dh/2] fp32. */
int d3d_rope(void* qkv, int64_t ld, const int* pos, const float* inv_freq, int T, int H, int Dh, int kind, void* stream);
/* The same rotation split in two: the per-token table tab [T, Dh] fp32 = [cos | sin] of pos*inv_freq is built once per prefill and
 * applied by every layer with 16-byte vector accesses (16-bit QKV; identical results to d3d_rope). */
int d3d_rope_table(const int* pos, const float* inv_freq, int T, int Dh, float* tab, void* stream);
int d3d_rope_apply(void* qkv, int64_t ld, const float* tab, int T, int H, int Dh, int kind, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Greedy decode with a KV cache (POL:463 `llava.generate(max_new_tokens=20, do_sample=False)`).
 * The prefill keeps every layer's packed QKV matrix [rows_cap, 3*hidden] (K already rotated): that IS the cache.  Decode step s
 * appends the new tokens' rows at t_prefill + s*n_seq + b.  HBM-bound on the weights: d3d_gemm_skinny streams each weight matrix once.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* rms1; const void* w_qkv; const void* w_o; const float* rms2;
  const void* w_gu;   /* row-interleaved gate/up [2*ffn, hidden] */
  const void* w_down;
} d3d_lm_layer;
typedef struct {
  int n_layers, hidden, n_heads, head_dim, ffn, vocab;
  int kind;                    /* D3D_F16 | D3D_BF16: weights, activations and cache */
  float eps;
  const d3d_lm_layer* layers;  /* HOST array [n_layers] of device pointers */
  const float* norm; const void* lm_head; const void* embed;
} d3d_lm_model;
/* C = epi(A @ W^T) for 1..16 activation rows (same argument struct and epilogue contract as d3d_gemm; K % 32 == 0). */
int d3d_gemm_skinny(const d3d_gemm_args* args_h, void* stream);
int d3d_gemm_skinny_set_config(int cfg);  /* 0 = default heuristic, 1..9 = tuning variants (tools/skinny_bench.py) */
/* out[r] = index of the FIRST maximum of x[r, :n] (torch.argmax / HF greedy search). */
int d3d_argmax_rows(const float* x, int64_t ld, int rows, int n, int* out, void* stream);
/* Attention of the step's n_seq new query rows over their sequences' cached keys / values (prefill rows [cu[b], cu[b+1]) + decode rows). */
int d3d_decode_attention(const void* qkv, int64_t ld, const int* cu_seqlens, int n_seq, int t_prefill, int step, int H, int Dh, int kind,
                         float scale, void* out, int64_t ldo, void* stream);
/* The same with the RoPE of the step's own rows fused: rope_tab = d3d_rope_table of the sequences' positions ([n_seq, Dh] fp32), the rows
 * t_prefill + step * n_seq + b still un-rotated.  q is rotated in registers, the new key in place (the cache row ends up as d3d_rope_apply
 * would have left it).  rope_tab == NULL: rows already rotated (= d3d_decode_attention). */
int d3d_decode_attention_rope(void* qkv, int64_t ld, const int* cu_seqlens, int n_seq, int t_prefill, int step, int H, int Dh, int kind,
                              float scale, const float* rope_tab, void* out, int64_t ldo, void* stream);
/* One decode step for all sequences: embeds tokens_in [n_seq], runs the layers (qkv_layers_h: HOST array of the per-layer cache base
 * pointers), writes logits [n_seq, vocab] fp32 and next_tokens [n_seq].  Scratch: x32 [n_seq,hidden] f32, a16 [n_seq,hidden], att16
 * [n_seq,hidden], h16 [n_seq,ffn] 16-bit, rope_tab [n_seq,head_dim] f32, pos [n_seq] int32. */
int d3d_lm_decode_step(const d3d_lm_model* m_h, void* const* qkv_layers_h, int64_t ld_qkv, const int* cu_seqlens, int n_seq, int t_prefill,
                       int step, const int* tokens_in, const float* inv_freq, float* x32, void* a16, void* att16, void* h16, float* rope_tab,
                       int* pos, float* logits, int* next_tokens, void* stream);
/* Switches of the decode step, for A/B measurements (default 1).  bit 0: the kernels of a step are chained with programmatic dependent
 * launch (weights of kernel i+1 stream while kernel i runs); bit 1: the register-staged CUDA-core decode attention + separate RoPE kernel
 * instead of the staged tensor-core one; bit 2: no L2 prefetch of the next kernel's weights; bits 3, 4: TIMING EXPERIMENTS ONLY (results are
 * wrong): every weight tile / key block is read from the same rows, i.e. the step without its HBM traffic. */
int d3d_lm_decode_set_pdl(int on);

/* dst[r, :D] = src[idx[r], :D] for n 16-bit rows (last-token rows of the final Phi-3 layer). */
int d3d_gather_rows16(const void* src, int64_t lds, const int* idx, void* dst, int64_t ldd, int n, int D, void* stream);
/* dst[dst_idx[r], :D] = src[r, :D] for n 16-bit rows (compact QKV rows of a prefill chunk -> their rows of the strided KV cache). */
int d3d_scatter_rows16(const void* src, int64_t lds, void* dst, int64_t ldd, const int* dst_idx, int n, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Coarse entry points: one call per transformer stack (csrc/forward_host.cu).  Bit-identical to issuing the per-kernel entries.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  const float* ln1_g; const float* ln1_b; const void* w_qkv; const float* b_qkv; const void* w_o; const float* b_o;
  const float* ln2_g; const float* ln2_b; const void* w_fc; const float* b_fc; const void* w_pr; const float* b_pr;
} d3d_vit_layer;
typedef struct {
  int n_layers, width, n_heads, patch, tokens, kpad, resolution, out_dim;
  int kind;                      /* D3D_F16 | D3D_BF16 */
  const void* conv_w;            /* [width, kpad] patch embedding (conv1 14x14/14 as a GEMM) */
  const float* cls; const float* pos; const float* ln_pre_g; const float* ln_pre_b;
  const d3d_vit_layer* layers;   /* HOST array [n_layers] of device pointers */
  const float* ln_post_g; const float* ln_post_b; const void* proj; /* NULL for the HF vision tower (hidden states only) */
} d3d_vit_model;
typedef struct {                 /* device scratch, sized for N images: T = N * tokens */
  void* cols;                    /* [N*(tokens-1), kpad] 16-bit */
  float* conv;                   /* [N*(tokens-1), width] */
  float* X;                      /* [T, width] residual stream; the hidden state when project = 0 */
  void* A16; void* qkv; void* att; void* h; void* out;   /* [T, width | 3 width | width | 4 width | out_dim] 16-bit */
  const int* cu;                 /* [N+1] = i * tokens */
} d3d_vit_scratch;
/* CLIPEncoder.forward + VisionTransformer.forward (ENC:273-284, CLIPM:219-238) for N u8 NHWC images: preprocessing, patch embedding,
 * n_layers_run residual blocks, then (project != 0) ln_post + proj into s->out [N, tokens, out_dim] (row 0 of an image = CLS), or
 * (project == 0) the fp32 hidden state in s->X -- the LLaVA tower's hidden_states[-2] with n_layers_run = n_layers - 1 (POL:441-452). */
int d3d_vit_forward(const d3d_vit_model* m_h, const uint8_t* img, int N, int Hin, int Win, int n_layers_run, int project,
                    const d3d_vit_scratch* s_h, void* stream);

typedef struct {                 /* device scratch for T packed tokens of n_seq sequences */
  void* A16; void* qkv; void* att; void* h;              /* [T, hidden | 3 hidden | hidden | ffn] 16-bit (qkv unused with a KV cache) */
  float* rope_tab;               /* [T, head_dim] */
  void* last16;                  /* [n_seq, hidden] 16-bit */
  void* att_last; float* x_last; /* [n_seq, hidden] 16-bit / fp32: last-token rows of the final layer */
} d3d_lm_scratch;
/* Prefill of llava.generate over packed inputs_embeds (POL:456-463): X [T, hidden] fp32 (overwritten: the residual stream), cu_seqlens
 * [n_seq+1], positions [T], last_rows [n_seq] -> logits [n_seq, vocab] fp32 of every sequence's last token.  qkv_layers_h != NULL: HOST array
 * of per-layer [>= T, ld_qkv] 16-bit buffers that keep the packed QKV matrices (K rotated) = the KV cache of d3d_lm_decode_step.
 * trim_last_layer != 0: the final layer's o_proj / MLP run on the last-token rows only (same logits, ~2 % fewer FLOPs). */
int d3d_phi3_prefill(const d3d_lm_model* m_h, float* X, int T, const int* cu_seqlens, const int* positions, int n_seq, int max_len,
                     const int* last_rows, const float* inv_freq, void* const* qkv_layers_h, int64_t ld_qkv, const d3d_lm_scratch* s_h,
                     int trim_last_layer, float* logits, void* stream);

/* Chunked prefill over a strided KV cache: the prompt of a navigation step is [2 text tokens | 576 patch tokens | instance | zone | text]
 * (POL:456) and, attention being causal, the first tokens do not depend on the 3D memory -- their prefill can run while the (host-
 * synchronised) memory update is in flight.  One call processes T compact rows X [T, hidden] (fp32 residual stream, overwritten) of n_seq
 * sequences: rows[t] = row of the KV cache that compact row t occupies (sequence s lives at cache rows seq_start[s] + position),
 * positions[t] its token position, seq_len[s] the number of keys visible to this pass; the pass computes the query tiles (128 positions)
 * [q_tile_begin, q_tile_end) of every sequence, i.e. its compact rows must be exactly the positions of those tiles (< seq_len).
 * cache_layers_h: HOST array of per-layer [cache_rows, ld_cache] 16-bit packed QKV buffers (zero-initialised once); att_cache
 * [cache_rows, hidden] 16-bit scratch.  last_rows (compact rows of the sequences' last tokens) != NULL: also final norm + lm_head ->
 * logits [n_seq, vocab].  Results are bit-identical to d3d_phi3_prefill (same kernels per row / tile; trim_last_layer as there). */
typedef struct {
  const int* seq_start; const int* seq_len; const int* rows; const int* positions;
  int q_tile_begin, q_tile_end;
} d3d_lm_chunk;
int d3d_phi3_prefill_chunk(const d3d_lm_model* m_h, float* X, int T, int n_seq, int max_len, const d3d_lm_chunk* c_h, void* const* cache_layers_h,
                           int64_t ld_cache, int64_t cache_rows, void* att_cache, const float* inv_freq, const d3d_lm_scratch* s_h,
                           const int* last_rows, int trim_last_layer, float* logits, void* stream);
/* Persistent GEMM kernels launch on at most `n` SMs (0 = all): leaves SMs free for a concurrent latency-bound stream (the view loop of the
 * 3D-memory update) while a throughput-bound stream runs transformer layers. */
int d3d_gemm_set_sm_limit(int n);

/* Device timing of every d3d_gemm launch between begin and end (CUDA events on the launch stream), per kernel variant
 * [0] gemm_tcgen05_pair_kernel<256> (CTA pairs), [1] gemm_tcgen05_kernel<256>, [2] gemm_tcgen05_kernel<128>: algorithmic FLOPs (2 M N K),
 * milliseconds and launch counts -- bench.py's `roofline` of the dominant kernel.  end synchronises the device. */
int d3d_gemm_profile_begin(void);
int d3d_gemm_profile_end(double* flops3, float* ms3, int* launches3);

/* embed_tokens (POL:439): out[t, :D] = table16[ids[t], :D] as fp32. */
int d3d_embed_gather(const void* table, int kind, const int* ids, int T, int D, float* out, int64_t ldo, void* stream);

/* CLIPEncoder preprocessing + im2col of the 14x14/14 patch embedding (ENC:267-284; CLIPM:220-222):
 * img u8 NHWC [N,Hin,Win,3] -> bicubic resize to RxR (A=-0.75, align_corners=False, rounded to u8 like torchvision;
 * identity when Hin==Win==R) -> /255 -> (x-mean)/std -> 16-bit -> out [N*(R/patch)^2, kpad], col = c*p*p+ky*p+kx,
 * zero padded to kpad (multiple of 8). */
int d3d_preprocess_im2col(const uint8_t* img, int N, int Hin, int Win, int R, int patch, const float* mean3_h,
                          const float* std3_h, void* out, int kpad, int kind, void* stream);

/* One pass of Pillow's ImagingResample (8 bits per channel, fixed-point taps, clip8) along H (axis 0) or W (axis 1) of u8 NHWC images:
 * the resize inside HF's CLIPImageProcessor, i.e. the LLaVA tower's input path `llava_processor(images=rgb)` (Policy_Dynam3D_VLN.py:438).
 * bounds [out_size,2] (first source index, tap count) and kk [out_size,ksize] int32 are DEVICE tables (ops.pil_bicubic_tables). */
int d3d_pil_resample_pass(const uint8_t* src, uint8_t* dst, int N, int H, int W, int C, int out_size, int axis, const int* bounds,
                          const int* kk, int ksize, void* stream);
/* ViT token assembly + ln_pre (CLIPM:223-225): out[n,0] = LN(cls+pos[0]); out[n,1+i] = LN(conv[n*(tokens-1)+i]+pos[1+i]). */
int d3d_vit_embed_ln(const float* conv, const float* cls, const float* pos, const float* gamma, const float* beta, float eps,
                     int N, int tokens, int D, float* out, void* stream);

/* dst[dst_idx?dst_idx[r]:r] = src[src_idx?src_idx[r]:r] for n fp32 rows of width D (slot writes FF:644-648,688,730,756). */
int d3d_scatter_rows(const float* src, int64_t lds, const int* src_idx, float* dst, int64_t ldd, const int* dst_idx, int n, int D,
                     void* stream);
/* a[i] += b[i] over n_rows*D contiguous fp32 (patch_features + patch_position_fts, POL:453). */
int d3d_add_inplace(float* a, const float* b, int64_t n_rows, int D, void* stream);
/* fp32 -> 16-bit cast of a [T,D] matrix. */
int d3d_cast16(const float* in, int64_t ldi, void* out, int64_t ldo, int T, int D, int kind, void* stream);

/* Variable-length multi-head self-attention, fp32 CUDA-core online-softmax version (nn.MultiheadAttention inside
 * nn.TransformerEncoderLayer, FF:134-137; also exact fallback for CLIPM:181-183 / HF attention).
 * qkv [T, 3*H*Dh] 16-bit packed (q|k|v), out [T, H*Dh] 16-bit, cu_seqlens [n_seq+1] int32, Dh in {64, 96}. */
int d3d_attention_simt(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                       int Dh, int causal, int kind, float scale, void* stream);

/* Tensor-core flash attention (mma.sync m16n8k16, online softmax, fp32 accumulate) with the same contract as
 * d3d_attention_simt; used for CLIP ViT (CLIPM:181-183) and the causal Phi-3 prefill.  Rows must be 16-byte aligned. */
int d3d_attention_mma(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                      int Dh, int causal, int kind, float scale, void* stream);
/* Non-causal, head_dim 64, for packed batches that mix many short sequences with a few long ones (the patch -> instance pooling pass of a
 * step: ~1 500 sequences of ~37 tokens): sequences of <= 64 tokens run one warp per (sequence, head), the others on the 128-row kernel. */
int d3d_attention_mixed(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H, int Dh,
                        int kind, float scale, void* stream);

/* tcgen05 flash attention (S = QK^T and O = PV on the 5th-gen tensor cores, accumulators in TMEM, Q/K/V tiles by TMA), head_dim 64
 * (CLIP ViT / LLaVA tower) or 96 (Phi-3: 64 + 32 column swizzle atoms), non-causal or causal; same packed-QKV contract as
 * d3d_attention_simt; n_rows = total rows T of the qkv matrix (for the TMA descriptors). */
int d3d_attention_tc(const void* qkv, int64_t ld, int64_t n_rows, void* out, int64_t ldo, const int* cu_seqlens, int n_seq,
                     int max_len, int H, int Dh, int causal, int kind, float scale, void* stream);
/* The same kernel over sequences that start at arbitrary rows seq_start[s] with explicit lengths seq_len[s] (NULL: packed, len = start[s+1] -
 * start[s]) and for the query tiles (128 rows) [q_tile_begin, q_tile_end) of every sequence only -- the chunked prefill over a strided KV
 * cache (d3d_phi3_prefill_chunk): keys [0, len) of the sequence, queries of the selected tiles. */
int d3d_attention_tc_ex(const void* qkv, int64_t ld, int64_t n_rows, void* out, int64_t ldo, const int* seq_start, const int* seq_len,
                        int n_seq, int max_len, int q_tile_begin, int q_tile_end, int H, int Dh, int causal, int kind, float scale, void* stream);
/* Tile shape of the tcgen05 attention per head dim: key halves per tile = softmax threads per query row.  1: 64-key tiles, one thread per
 * row, more resident CTAs per SM (default at head_dim 96: two CTAs instead of one); 2: 128-key tiles, two threads per row (default at 64). */
int d3d_attention_tc_set_halves(int halves_d64, int halves_d96);
/* Split-operand ("precise", <= 1e-3 logits) attention on tcgen05: S = Qh Kh^T + Ql Kh^T + Qh Kl^T, O += Ph Vh + Pl Vh + Ph Vl, fp32
 * accumulate and output.  qkv_hl: [n_rows, >= lo_off + 3 H Dh] fp16, hi parts in columns [0, 3 H Dh), lo parts lo_off columns to the right
 * (d3d_split16, 2 terms); out: fp32 [n_rows, H Dh], ldo in floats.  head_dim 64 / 96; sequences of any length (built for >= 256). */
int d3d_attention_split_tc(const void* qkv_hl, int64_t ld, int64_t n_rows, int64_t lo_off, float* out, int64_t ldo, const int* cu_seqlens, int n_seq,
                           int max_len, int H, int Dh, int causal, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Token builders for the layer-wise pooling (patch -> instance -> zone) and the merge discriminator.
 * Sequences of several episodes are packed in one batch; seq_* are per-sequence base addresses (device pointers
 * stored as int64) of the owning episode's pools; tok_seq[t] / tok_src[t] give each packed token's sequence and its
 * row in that pool (-1 = the learned aggregate token that heads every sequence).
 * ------------------------------------------------------------------------------------------------ */

/* 7-d / 4-d position feature rows (FF:584-591 mode 0; FF:719-723 mode 1), 8 wide 16-bit (GEMM A operand, K padded):
 * [xyz-centre, |xyz|, sin(dir), cos(dir), scale, 0] or [xyz-centre, |xyz|, 0,0,0,0]; centre [n_seq,3]. */
int d3d_pool_features(const int64_t* seq_xyz, const int64_t* seq_dir, const int64_t* seq_scale, const float* centre,
                      const int* tok_seq, const int* tok_src, int T, int mode, void* out16, int kind, void* stream);

/* X[t] = aggregate token (src<0) or emb[t] + fts[src] (FF:592-593, 674-676, 725-727); fts rows fp16 or fp32 of width D. */
int d3d_pool_assemble(const float* emb, const int64_t* seq_fts, int fts_is_f32, const int* tok_seq, const int* tok_src,
                      const float* agg, int T, int D, float* X, void* stream);

/* instance_merge_discriminator input (FF:613-617): row (g,j) = [inst_fts[idx[g,j]] | view_fts[g] | centre[g]-inst_pos[idx[g,j]] | 0],
 * 16-bit rows of ldo >= 2*D+3 elements; idx [G,K] int32. */
int d3d_disc_input(const float* inst_fts, const float* inst_pos, const int* idx, const float* view_fts, const float* centre,
                   int G, int K, int D, int ldo, void* out16, int kind, void* stream);

/* Policy-side A operands (POL:432-435): patch info rows [rel_x, rel_y, rel_z, sin(dir), cos(dir), scale, 0, 0] from the
 * [5,n] planes of d3d_patch_3d_info; [a | b] concatenation cast to 16 bit; 3-d positions padded to 8. */
int d3d_patch_info_rows(const float* info5, int64_t n, void* out16, int kind, void* stream);
int d3d_concat2_cast(const float* a, const float* b, int n, int D, void* out16, int kind, void* stream);
int d3d_pos3_rows(const float* x, int n, void* out16, int kind, void* stream);

/* Pointer-table (batched over episodes) variants: one launch serves every episode of the rank. */
/* n independent 16-byte-aligned block copies (append a view's patches to the episode pools, FF:557-570). */
int d3d_copy_blocks(const int64_t* src_ptr, const int64_t* dst_ptr, const int64_t* nbytes, int n, void* stream);
/* *(float*)dst_row_ptr[r] <- src[src_idx?src_idx[r]:r] (fp32 rows of width D): slot writes across episodes (FF:644-648,688,730,756). */
int d3d_scatter_rows_ptr(const float* src, int64_t lds, const int* src_idx, const int64_t* dst_row_ptr, int n, int D, void* stream);
/* 2-NN of every query against ITS OWN reference set ref_ptr[q] ([n_ref[q],3] fp32), same arithmetic / tie-break as d3d_knn3d;
 * missing neighbours: d2=+inf, idx=-1.  out [n_q,2]. */
int d3d_knn2_batched(const int64_t* ref_ptr, const int* n_ref, const float* queries, int n_q, float* out_d2, int* out_idx, void* stream);
/* d3d_disc_input with per-query instance pools (fts_ptr[q], pos_ptr[q]); rows with idx<0 are zero. */
int d3d_disc_input_batched(const int64_t* fts_ptr, const int64_t* pos_ptr, const int* idx, const float* view_fts, const float* centre,
                           int Q, int K, int D, int ldo, void* out16, int kind, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One-call pooling pass (host orchestration inside the library): all segments / zones of a view, all episodes.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {            /* nn.Sequential(Linear, LayerNorm, GELU, Linear) (FF:139-143,148-152,157-161; POL:83-111) */
  const void* w0; const float* b0;      /* [d_hidden, k_pad] 16-bit (K zero-padded to a multiple of 8), [d_hidden] */
  const float* ln_g; const float* ln_b; /* [d_hidden] */
  const void* w3; const float* b3;      /* [d_out, d_hidden] 16-bit, [d_out] */
  int k_pad, d_hidden, d_out, kind;
} d3d_mlp;
typedef struct {            /* one nn.TransformerEncoderLayer, post-norm, GELU (FF:134-137) */
  const void* w_in; const float* b_in;    /* [3d, d], [3d] */
  const void* w_out; const float* b_out;  /* [d, d], [d] */
  const float* n1_g; const float* n1_b;
  const void* w1; const float* b1;        /* [4d, d], [4d] */
  const void* w2; const float* b2;        /* [d, 4d], [d] */
  const float* n2_g; const float* n2_b;
} d3d_encoder_layer;
typedef struct {            /* position MLP + aggregate token + 2-layer encoder + final norm of one pooling level */
  d3d_mlp mlp;
  const float* agg;                       /* [d] learned aggregate token (FF:145,154) */
  d3d_encoder_layer layers[2];
  const float* norm_g; const float* norm_b; float norm_eps;
  int n_layers, d_model, n_head;
} d3d_pool_level;

size_t d3d_pool_workspace_bytes(int T, int d_model, int d_hidden_mlp);
/* out[T, ldo] = Linear(GELU(LN(Linear(A0)))) with fp32 accumulate; hidden32 [T,d_hidden] fp32 and hidden16 [T,d_hidden] 16-bit scratch. */
int d3d_mlp_ln_gelu(const d3d_mlp* m_h, const void* A0, int64_t lda, int T, void* hidden32, void* hidden16, float* out, int64_t ldo,
                    void* stream);
/* features -> MLP -> assemble -> encoder -> out [n_seq, d_model] (token 0 of every sequence).  seq_ptrs [4, n_seq] int64 device
 * table (xyz, dir, scale, fts base addresses per sequence); other arguments as d3d_pool_features / d3d_pool_assemble. */
int d3d_pool_tokens(const d3d_pool_level* lvl_h, const int64_t* seq_ptrs, const float* centre, const int* tok_seq, const int* tok_src,
                    const int* cu_seqlens, int T, int n_seq, int max_len, int mode, int fts_is_f32, void* workspace,
                    size_t workspace_bytes, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-side bookkeeping of the 3D token memory (CPU code in the same library; ALL pointers below are HOST pointers).
 * Owns what the reference keeps in Python dicts (FF:164-177) for every episode of a Feature_Fields and plans each view:
 * d3d_ffh_begin_view -> [pooling / K-NN / discriminator kernels] -> d3d_ffh_finish_view -> d3d_ffh_fetch_view -> [slot writes,
 * merged-instance and zone pooling kernels].  Literal FF:362-393, 433-475, 623-756, 759-812 incl. Q2, Q3, Q5, Q9.
 * ------------------------------------------------------------------------------------------------ */
void* d3d_ffh_create(int batch_size, int num_proposal, float zone_len);
void d3d_ffh_destroy(void* h);
int d3d_ffh_reset(void* h, int batch_size);                 /* FF:186-206 */
int d3d_ffh_pop(void* h, int index);                        /* FF:210-229 */
int d3d_ffh_counts(void* h, int b, int64_t* counts8);       /* n_patch, n_p2i, n_inst, live inst, n_zone, live zones, tree, last K */
int d3d_ffh_cull(void* h, int b, const uint8_t* mask, int64_t n, int64_t* dead_inst, int* n_dead_inst, int64_t* dead_zone,
                 int* n_dead_zone); /* Same bookkeeping from the compacted list of culled rows (d3d_frustum_cull_batched): rows [n] in any order. */
int d3d_ffh_cull_list(void* h, int b, const int32_t* rows, int64_t n, int64_t* dead_inst, int* n_dead_inst, int64_t* dead_zone, int* n_dead_zone);
                        /* FF:362-393 */
int d3d_ffh_set_tree(void* h);                              /* FF:396 */
int d3d_ffh_begin_view(void* h, const float* xyz, const int64_t* segm, int P, const int64_t* stage_off, int64_t* base_rows, int* n_seg,
                       int* seq_owner, int* members, int* cu_m, int* tok_src, int* tok_seq, int* cu_tok, int* n_ref, int* info);
/* Whole-step planning: the patch -> instance pooling of a view is independent of the memory state, so all V views are planned (and pooled
 * on the device) as ONE packed batch; begin_view_refs then makes view ix the view in flight for finish_view / fetch_view. */
int d3d_ffh_begin_step(void* h, const float* xyz /*[V,B,P,3]*/, const int64_t* segm /*[V,B,P]*/, int P, int V, int64_t* base_rows,
                       int* view_seq_start, int* seq_owner, int* members, int* cu_m, int* tok_src, int* tok_seq, int* cu_tok, int* info);
int d3d_ffh_begin_view_refs(void* h, int ix, int* n_ref);
int d3d_ffh_finish_view(void* h, const float* res12, int* sizes10, int64_t* after3);
/* Per-view runtime (the interpreter-free form of the view loop): see csrc/ff_host.cu.  d3d_ff_pools: device base addresses of one
 * episode's pools; d3d_ff_runtime: weights, workspace, the pinned/device upload ring and scratch owned by the caller.
 *   d3d_ff_view_pre : K-NN proposals + merge discriminator for view ix, result copy, previous view's deferred zone pass, wait, planner
 *                     (same sizes10 / after3 as d3d_ffh_finish_view);  the caller then grows pools if after3 asks for more slots;
 *   d3d_ff_view_post: slot writes + merged-instance pooling pass, upload of the zone pass (issued by the next pre / d3d_ff_run_deferred). */
typedef struct { int64_t patch_pos, patch_dir, patch_scale, patch_fts, inst_pos, inst_fts, zone_pos, zone_fts; } d3d_ff_pools;
typedef struct {
  const d3d_pool_level* level_inst; const d3d_pool_level* level_zone; const d3d_mlp* disc;
  void* workspace; size_t workspace_bytes;                 /* d3d_pool_tokens workspace */
  void* stage_dev; void* stage_host; size_t stage_bytes;   /* upload ring: device buffer + pinned host mirror (>= 4 views of uploads) */
  float* res_dev; float* res_host;                         /* [max_seq, 12] */
  float* d2; int* idx;                                     /* [max_seq, 2] */
  void* disc_in; float* disc_h32; void* disc_h16; float* disc_out;  /* [2*max_seq, k_pad] 16-bit, [2*max_seq, hidden] f32 / 16-bit, [2*max_seq, 4] f32 */
  float* out_merge; float* out_zone;                       /* [max_seq, d_model] f32 */
  void* event;                                             /* cudaEvent_t */
  int max_seq;
} d3d_ff_runtime;
int d3d_ff_view_pre(void* h, int ix, const d3d_ff_runtime* rt, const d3d_ff_pools* pools_h, const float* centres_step, const float* view_fts_step,
                    int* sizes10, int64_t* after3, void* stream);
int d3d_ff_view_post(void* h, const d3d_ff_runtime* rt, const d3d_ff_pools* pools_h, float* centres_step, float* view_fts_step, void* stream);
int d3d_ff_run_deferred(void* h, const d3d_ff_runtime* rt, void* stream);
/* Growable device arrays with stable addresses for the episode pools (csrc/vmm_pool.cu): a reserved virtual range, physical chunks committed
 * on demand (CUDA virtual memory management).  Replaces the reference's grow-by-concatenation (feature_fields.py:557-570, 643-648, 715-730).
 * create: reserve va_bytes (chunk_bytes = commit granularity, 0 = 32 MiB); ensure: commit until >= bytes are mapped, zero-filling new chunks
 * on `stream` (no copy, no pointer change, no device synchronisation); destroy: the caller synchronises first. */
void* d3d_vmm_create(size_t va_bytes, size_t chunk_bytes, int device);
int d3d_vmm_ensure(void* pool, size_t bytes, void* stream);
uint64_t d3d_vmm_base(void* pool);
size_t d3d_vmm_mapped(void* pool);
size_t d3d_vmm_reserved(void* pool);
int d3d_vmm_destroy(void* pool);
/* Device timing of the view runtime's phases (CUDA events on the launch stream), for bench.py's per-stage roofline table.
 * begin: start recording; end: synchronise and return, per phase [knn, disc, new_slots, merge_pool, zone_pool, result_copy]: milliseconds,
 * algorithmic work (bytes for knn / new_slots / result_copy, FLOPs for the others) and the number of recorded scopes. */
int d3d_ff_profile_begin(void);
int d3d_ff_profile_end(float* ms6, double* work6, int* launches6);
void* d3d_event_create(void);   /* cudaEvent_t without timing, for d3d_ff_runtime.event */
void d3d_event_destroy(void* e);

int d3d_ffh_fetch_view(void* h, int* new_src, int* new_owner, int64_t* new_iid, int* mg_owner, int64_t* mg_iid, float* mg_pos,
                       int* mg_tok_src, int* mg_tok_seq, int* mg_cu, int* zn_owner, int64_t* zn_slot, int* zn_keys, float* zn_pos,
                       int* zn_tok_src, int* zn_tok_seq, int* zn_cu);
int d3d_ffh_zone_key_array(void* h, int b, float* out);
int d3d_ffh_get_map(void* h, int b, int which, int64_t* ids, int64_t* lens, int64_t* cat, int64_t* sizes2);
int d3d_ffh_live_ids(void* h, int b, int which, int64_t* ids, int64_t* n_out);  /* dict-order keys of map `which` (ids may be NULL) */
int d3d_ffh_get_p2i(void* h, int b, int64_t* out);
int d3d_ffh_get_patch_pos(void* h, int b, float* out);
int d3d_ffh_get_zone_keys(void* h, int b, float* keys, int64_t* ids, int64_t* n);
int d3d_ffh_get_last(void* h, int b, float* d2, int* idx, uint8_t* merge, int64_t* n);

/* ------------------------------------------------------------------------------------------------
 * "Precise" pipeline (parity evidence against the reference's fp32 CPU path; not the production mode): activations stay
 * fp32 between kernels (every producer above accepts kind = D3D_OUT_F32) and each tensor-core GEMM runs on split operands:
 * d3d_split16 writes [hi | lo (| hi)] along K (hi = fp16(x), lo = fp16(x - hi)), the weight is stored as [W_hi | W_hi (| W_lo)],
 * so the unchanged tcgen05 GEMM computes A_hi W_hi + A_lo W_hi (+ A_hi W_lo) with fp32 accumulation (~2^-22 operand precision).
 * ------------------------------------------------------------------------------------------------ */
int d3d_split16(const float* in, int64_t ldi, void* out16, int64_t ldo, int T, int K, int terms, void* stream);
/* fp32-in / fp32-out variant of d3d_attention_simt (no 16-bit rounding of q, k, v or the output). */
int d3d_attention_f32(const float* qkv, int64_t ld, float* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                      int Dh, int causal, float scale, void* stream);
/* Attention on the tensor cores with split fp16x2 operands (x = hi + lo; S = Qh Kh + Ql Kh + Qh Kl, O += Ph Vh + Pl Vh + Ph Vl, fp32 softmax,
 * fp32 output): ~22-bit products at mma.sync speed -- the precise mode's attention for sequences of >= 64 tokens.  qkv_hl [T, ld] fp16 is the
 * packed QKV matrix split by d3d_split16: hi halves in columns [0, 3*H*Dh), lo halves at column lo_off. */
int d3d_attention_split(const void* qkv_hl, int64_t ld, int64_t lo_off, float* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                        int Dh, int causal, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pretrain novel-view patch renderer, render_view_3d_patch (PFF:494-625), habitat mode.  The two torch_kdtree queries are
 * d3d_knn3d (K = 4); the tinycudann CutlassMLPs run on d3d_gemm (bias-free, D3D_ACT_LEAKY_RELU).
 * ------------------------------------------------------------------------------------------------ */
/* ray sample points (PFF:408-422, 523-528), evaluated in fp64 like the numpy code and rounded once: rel_y [n_samples] fp64,
 * tan_x / tan_z [n_rays] fp32 per-ray tangents -> out_xyz [n_rays*n_samples, 3] fp32. */
int d3d_ray_points_habitat(const double* rel_y, const float* tan_x, const float* tan_z, int n_rays, int n_samples, double cos_h,
                           double sin_h, double cam_x, double cam_y, double cam_z, float* out_xyz, void* stream);
/* PFF:542-552: sqrt, radius cut-off (idx <- -1 in place), density proxy 1/sum(dist), top-n_top samples per ray (ties: lowest index). */
int d3d_ray_topk(const float* d2, int* idx, int n_rays, int n_samples, int K, float radius, int n_top, int* topk, void* stream);
int d3d_gather_samples(const float* ray_xyz, const int* topk, int n_rays, int n_samples, int n_top, float* out, void* stream);
/* PFF:586-615: neighbour gather of the selected samples into GEMM operands: pos_rows [n_points*K, 8] (kind pos_kind) and
 * feat_rows16 [n_points*K, D] fp16; idx is updated in place with the radius cut-off. */
int d3d_nerf_gather(const float* d2, int* idx, const float* sample_xyz, const float* patch_xyz, const float* patch_dir,
                    const float* patch_scale, const void* patch_fts16, const float* ray_dir, int n_points, int n_top, int K, int D,
                    float radius, float far_, float cam_dir, float cos_neg, float sin_neg, void* pos_rows, int pos_kind,
                    void* feat_rows16, void* stream);
/* out16 = fp16(fp16(a32) + b16) (PFF:478-482). */
int d3d_add_half(const float* a32, const void* b16, void* out16, int64_t n, void* stream);
/* raw2feature (PFF:446-474): feat [n_rays*n_top, D] fp32, density [n_rays*n_top] -> feature_map [n_rays, D] (L2-normalised), depth_map [n_rays]. */
int d3d_volume_render(const float* feat, const float* density, const int* topk, const float* rel_dist, int n_rays, int n_samples,
                      int n_top, int D, float* feature_map, float* depth_map, void* stream);

/* ---- candidate-waypoint predictor: BinaryDistPredictor_TRM (TRM_net.py:9-88) + heat-map NMS (POL:226-247,
 * waypoint_pred/utils.py:37-66).  The depth-encoder embeddings ([B*12, 128, 4, 4], ENC:15-109) are an INPUT.  Linear layers: d3d_gemm on
 * split operands (fp32-class accuracy), LayerNorm: d3d_layernorm (eps 1e-12). ---- */
/* in-place ReLU on n fp32 values */
int d3d_wp_relu(float* x, int64_t n, void* stream);
/* BERT self-attention over the n_img <= 12 views of every episode with an additive mask ([n_img, n_img] fp32, 0 / -10000 as
 * WBERT:184-185 builds it from utils.py:90-102): qkv [n_episodes*n_img, 3*H*Dh] fp32 rows = [q | k | v], out [n_episodes*n_img, H*Dh]. Dh = 64. */
int d3d_wp_neighbor_attention(const float* qkv, const float* add_mask, int n_episodes, int n_img, int H, int Dh, float scale, float* out, void* stream);
/* logits [n_episodes, n_angles, n_classes] fp32 (after the HEATMAP_OFFSET roll, TRM:84-86) -> prob (softmax over the whole map; may be NULL)
 * and nms [n_episodes, n_angles, n_classes]: the map wrapped by one angle row, `max_predictions` rounds of {first arg-max; keep its
 * probability; zero the box |dx| <= sigma_x (circular over the class axis), |dy| <= sigma_y around (x, y = index / n_classes as a FLOAT)},
 * un-wrapped.  Non-zero cells of nms are the candidate waypoints. */
int d3d_wp_heatmap_nms(const float* logits, int n_episodes, int n_angles, int n_classes, int max_predictions, float sigma_x, float sigma_y,
                       float* prob, float* nms, void* stream);

#ifdef __cplusplus
}
#endif
#endif
