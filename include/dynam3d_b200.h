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

const char* d3d_last_error(void);
int d3d_version(void);
/* 0 when device `dev` is an sm_100 part, D3D_EARCH otherwise */
int d3d_check_device(int dev);
int d3d_sm_count(void);

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

/* Plain CUDA-core reference GEMM with the same contract (debug / self-check only, slow). */
int d3d_gemm_simt(const d3d_gemm_args* args_h, void* stream);

#ifdef __cplusplus
}
#endif
#endif
