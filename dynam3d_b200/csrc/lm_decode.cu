// Greedy decode with a KV cache for the llava-phi-3-mini language model (POL:463: llava.generate(max_new_tokens=20, do_sample=False);
// SURVEY.md 8(f) rank 1).  A decode step multiplies 8..16 activation rows with every weight matrix once, so it is HBM-bound on the
// weights (7.6 GB per token): the kernels here are written for the memory roofline, not for the tensor pipe.
//
//   skinny_gemm_kernel  : out[M<=16, N] = epi(A[M,K] @ W[N,K]^T).  One CTA per 16 weight rows, 8 warps split K; every thread streams
//                         16-byte pieces of its weight row straight from HBM (L1 no-allocate) and feeds mma.sync.m16n8k16 with a
//                         K-permuted fragment layout (the SAME permutation on A and W, so no shuffles or smem staging are needed);
//                         fp32 partials of the 8 warps are summed in a fixed order in smem (deterministic), then +bias / SwiGLU /
//                         +fp32 residual and the store.
//   decode_attn_kernel  : one CTA per (sequence, head); keys / values = the sequence's prefill rows plus its rows of the earlier
//                         decode steps in the per-layer QKV cache; 16 warps, each loads 8 keys' K and V rows back to back (8-byte
//                         loads), batched online softmax, 16-way merge in smem.
//   argmax_kernel       : first maximum of every logit row (torch.argmax / HF greedy tie-break).
//   d3d_lm_decode_step  : the whole step (32 layers) behind ONE C call, so the host issues ~260 launches without interpreter overhead.
// Programmatic dependent launch: every kernel of the step is launched with programmaticStreamSerialization and triggers its dependents at
// once, so kernel i+1 is resident while kernel i still runs: its producer warp streams WEIGHTS (which no kernel writes) into its shared-memory
// ring, and only the threads that touch activations execute griddepcontrol.wait.  HBM stays busy across the ~260 kernel boundaries of a
// step instead of draining and refilling at each (a 19..100 MB weight matrix is a 3..15 us stream; the boundary cost was of that order).
#include <cmath>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

extern "C" int d3d_rmsnorm(const float* x, int64_t ldx, const int* row_index, const float* w, float eps, int T, int D, float* out32,
                           int64_t ld32, void* out16, int64_t ld16, int kind16, void* stream);
extern "C" int d3d_rope_table(const int* pos, const float* inv_freq, int T, int Dh, float* tab, void* stream);
extern "C" int d3d_rope_apply(void* qkv, int64_t ld, const float* tab, int T, int H, int Dh, int kind, void* stream);

namespace {

// PDL device side: no-ops when the kernel was launched without the attribute
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool g_decode_pdl = true;
bool g_decode_prefetch = true;
int g_decode_dbg = 0;  // timing experiments only (wrong results): 1 = GEMM weights always from rows 0..15 (no HBM traffic), 2 = attention keys always block 0

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_decode_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

constexpr int SK_WARPS = 8;
// tuning knobs (template parameters of the kernel, chosen per problem in skinny()):
//   SK_NT  n-tiles (8 weight rows each) per CTA;  SK_U  K super-steps (128 elements each) whose loads a warp issues before it multiplies

template <bool NOALLOC>
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 v;
  if (NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1, int kind) {
  if (kind == D3D_BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// L2 prefetch of the NEXT kernel's weights, issued by the producer warps once their own rows are on their way: the tail of a kernel and
// the boundary to the next one (~6..8 us with HBM otherwise idle) are spent pulling the next matrix into the 126 MB L2
struct NextWeights {
  const void* p;
  long long bytes;
  int dbg_same_rows;  // timing experiment: stream rows 0..15 for every tile
  int stamp_slot;     // debug builds: launch counter
};
__device__ __forceinline__ void l2_prefetch_share(const NextWeights& nx, int lane) {
  if (nx.bytes <= 0) return;
  constexpr long long CH = 4096;
  const long long per = ((nx.bytes / gridDim.x) + CH - 1) / CH * CH;
  const long long beg = per * blockIdx.x;
  const long long end = beg + per < nx.bytes ? beg + per : nx.bytes;
  for (long long off = beg + lane * CH; off < end; off += 32 * CH) {
    const unsigned sz = (unsigned)((end - off < CH ? end - off : CH) & ~15LL);
    if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)nx.p + off), "r"(sz) : "memory");
  }
}

// RMSNorm of one fp32 row by one warp: the arithmetic and summation order of rmsnorm_kernel (nn_kernels.cu; float4 per lane strided by 32,
// sequential per-lane sum, butterfly), with the loads issued in batches of 8 -- a plain runtime loop serialises 24 L2 round trips per pass
// (measured: 11 us for 8 rows).  CG: read through L2 (rows written by other CTAs of the same grid).
template <bool CG>
__device__ __forceinline__ void warp_rmsnorm_row(const float* __restrict__ x, const float* __restrict__ w, float eps, int D, uint16_t* __restrict__ out16,
                                                 int kind, int lane) {
  const float4* xr = reinterpret_cast<const float4*>(x);
  const float4* wr = reinterpret_cast<const float4*>(w);
  const int n4 = D / 128;
  auto ldx = [&](int c4) { return CG ? __ldcg(xr + c4) : xr[c4]; };
  float q = 0.f;
  for (int i0 = 0; i0 < n4; i0 += 8) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = i0 + j < n4 ? ldx(lane + 32 * (i0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i0 + j < n4) q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float r = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
  for (int i0 = 0; i0 < n4; i0 += 8) {
    float4 v[8], g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool in = i0 + j < n4;
      v[j] = in ? ldx(lane + 32 * (i0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[j] = in ? wr[lane + 32 * (i0 + j)] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i0 + j < n4)
        reinterpret_cast<uint2*>(out16)[lane + 32 * (i0 + j)] =
            make_uint2(pack16x2(v[j].x * r * g[j].x, v[j].y * r * g[j].y, kind), pack16x2(v[j].z * r * g[j].z, v[j].w * r * g[j].w, kind));
  }
}

struct SkinnyEpi {
  void* C; long long ldc;
  const float* bias; const float* residual; long long ldres;
  int act, out_kind;
  // Fused RMSNorm of the finished fp32 rows C[M, N] (the residual stream) into the 16-bit operand of the next GEMM, done by the LAST CTA to
  // finish (arrival counter): a separate 1-CTA kernel between two GEMMs cost ~11 us of dependency latency per instance in the step's chain
  const float* norm_w; uint16_t* norm_out; float norm_eps; int norm_kind; int* norm_counter;
};

// one thread per (tile, row, column pair): fixed-order sum over the 8 K-slices (deterministic), then bias / activation / residual / store
template <int NT>
__device__ __forceinline__ void skinny_epilogue(const float (*red)[NT][16][8], int n_base, int M, int N, const SkinnyEpi& ep) {
  const int e = threadIdx.x;
  if (e < NT * 16 * 4) {
    const int t = e / 64, r = (e / 4) % 16, cp = e % 4;
    const int n = n_base + t * 8 + cp * 2;
    if (r < M && n < N) {
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) {
        v0 += red[w][t][r][cp * 2];
        v1 += red[w][t][r][cp * 2 + 1];
      }
      const bool has1 = n + 1 < N;
      if (ep.bias) {
        v0 += ep.bias[n];
        if (has1) v1 += ep.bias[n + 1];
      }
      if (ep.act == D3D_ACT_SWIGLU) {  // row-interleaved gate/up: columns (2j, 2j+1) -> output column j
        const float o = __fdividef(v0, 1.0f + __expf(-v0)) * v1;
        const long long oc = n >> 1;
        if (ep.out_kind == D3D_OUT_F32) ((float*)ep.C)[(long long)r * ep.ldc + oc] = o;
        else st16(ep.C, (size_t)((long long)r * ep.ldc + oc), o, ep.out_kind);
      } else {
        if (ep.act == D3D_ACT_QUICK_GELU) { v0 = quick_gelu(v0); v1 = quick_gelu(v1); }
        else if (ep.act == D3D_ACT_GELU) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
        else if (ep.act == D3D_ACT_SILU) { v0 = silu(v0); v1 = silu(v1); }
        else if (ep.act == D3D_ACT_LEAKY_RELU) { v0 = v0 > 0.f ? v0 : 0.01f * v0; v1 = v1 > 0.f ? v1 : 0.01f * v1; }
        if (ep.residual) {
          v0 += ep.residual[(long long)r * ep.ldres + n];
          if (has1) v1 += ep.residual[(long long)r * ep.ldres + n + 1];
        }
        if (ep.out_kind == D3D_OUT_F32) {
          ((float*)ep.C)[(long long)r * ep.ldc + n] = v0;
          if (has1) ((float*)ep.C)[(long long)r * ep.ldc + n + 1] = v1;
        } else {
          st16(ep.C, (size_t)((long long)r * ep.ldc + n), v0, ep.out_kind);
          if (has1) st16(ep.C, (size_t)((long long)r * ep.ldc + n + 1), v1, ep.out_kind);
        }
      }
    }
  }
}

// Fragment trick: a K "super-step" is 128 elements.  Thread (g = lane/4, kq = lane%4) loads the 64 contiguous bytes (32 K-elements at
// k0 = 128*ss + 32*kq) of A row g / g+8 and of W row n0+g, so a warp instruction group reads 256 contiguous bytes of each of its 8 weight
// rows (DRAM-friendly).  Every 16-byte piece feeds two MMAs: elements {0,1} are K-slots (2kq, 2kq+1), {2,3} slots (2kq+8, 2kq+9) of the
// first, {4,5} / {6,7} of the second.  A and W use the same slot -> k map, so the products pair up and each k is used exactly once.
// HI: rows 8..15 of A exist (M > 8); otherwise their fragment registers are zero and never loaded.
template <bool HI, int SK_NT, int SK_U, bool NOALLOC>
__global__ void __launch_bounds__(SK_WARPS * 32) skinny_gemm_kernel(const uint16_t* __restrict__ A, long long lda, const uint16_t* __restrict__ W,
                                                                    long long ldw, int M, int N, int K, int kind, SkinnyEpi ep) {
  __shared__ float red[SK_WARPS][SK_NT][16][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, kq = lane & 3;
  const int n_base = blockIdx.x * (8 * SK_NT);
  float acc[SK_NT][4];
#pragma unroll
  for (int t = 0; t < SK_NT; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
  const bool lo_ok = g < M, hi_ok = HI && g + 8 < M;
  const uint16_t* a_lo_p = A + (long long)(lo_ok ? g : 0) * lda + kq * 32;
  const uint16_t* a_hi_p = A + (long long)(hi_ok ? g + 8 : 0) * lda + kq * 32;
  const uint16_t* w_p[SK_NT];
  bool w_ok[SK_NT];
#pragma unroll
  for (int t = 0; t < SK_NT; ++t) {
    const int n = n_base + t * 8 + g;
    w_ok[t] = n < N;
    w_p[t] = W + (long long)(w_ok[t] ? n : 0) * ldw + kq * 32;
  }
  // the warp owns super-steps warp, warp+8, ...; it issues the loads of SK_U super-steps back to back, then multiplies.
  // K % 128 != 0: the 32-element pieces beyond K are skipped (K % 32 == 0 is required)
  const int ssteps = (K + 127) >> 7;
  for (int ss0 = warp; ss0 < ssteps; ss0 += SK_WARPS * SK_U) {
    uint4 wv[SK_U][SK_NT][4], al[SK_U][4], ah[SK_U][4];
#pragma unroll
    for (int u = 0; u < SK_U; ++u) {
      const int off = (ss0 + u * SK_WARPS) << 7;
      const bool in = off + kq * 32 < K;
#pragma unroll
      for (int t = 0; t < SK_NT; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) wv[u][t][c] = (in && w_ok[t]) ? ldg_stream16<NOALLOC>(w_p[t] + off + c * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        al[u][c] = (in && lo_ok) ? __ldg(reinterpret_cast<const uint4*>(a_lo_p + off + c * 8)) : make_uint4(0, 0, 0, 0);
        if (HI) ah[u][c] = (in && hi_ok) ? __ldg(reinterpret_cast<const uint4*>(a_hi_p + off + c * 8)) : make_uint4(0, 0, 0, 0);
        else ah[u][c] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < SK_U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int t = 0; t < SK_NT; ++t) {
          mma16816(acc[t], al[u][c].x, ah[u][c].x, al[u][c].y, ah[u][c].y, wv[u][t][c].x, wv[u][t][c].y, kind);
          mma16816(acc[t], al[u][c].z, ah[u][c].z, al[u][c].w, ah[u][c].w, wv[u][t][c].z, wv[u][t][c].w, kind);
        }
  }
#pragma unroll
  for (int t = 0; t < SK_NT; ++t) {
    red[warp][t][g][kq * 2] = acc[t][0];
    red[warp][t][g][kq * 2 + 1] = acc[t][1];
    red[warp][t][g + 8][kq * 2] = acc[t][2];
    red[warp][t][g + 8][kq * 2 + 1] = acc[t][3];
  }
  __syncthreads();
  skinny_epilogue<SK_NT>(red, n_base, M, N, ep);
}

// ---- bulk-copy variant of the skinny GEMM -------------------------------------------------------------------------------------
// The register-staged kernel above keeps the LSU busy (eight 64-byte row pieces per warp load) and tops out near 3.6 TB/s.  Here a
// producer warp streams the CTA's 16 weight rows with cp.async.bulk (1 KB per row and stage, mbarrier complete_tx) into a 6-stage
// shared-memory ring -- no LSU wavefronts, no register staging, ~96 KB in flight per CTA -- and 8 consumer warps read their fragments
// with conflict-free 16-byte LDS (row pitch 1088 B) using the same K-permuted fragment trick.
constexpr int BK_ROWS = 16;                 // weight rows per CTA (2 n-tiles)
// K elements per stage = 2 KB per row and bulk copy.  The copy engine takes ~75 cycles per cp.async.bulk whatever its size (clock stamps,
// tools/skinny_stamps.py): with 1 KB copies a producer needed 1 250 cycles per 16-row stage -- 13 B / clk, at two CTAs per SM just the
// SM's share of HBM -- and a decode step whose weights all sat in L2 was hardly faster than one that streamed them.
constexpr int BK_KC = 1024;
constexpr int BK_U = BK_KC / 256;           // 32-element K-steps per warp and stage (8 warps split a stage's K)
constexpr int BK_PITCH = BK_KC * 2 + 64;    // 2112 B: consecutive rows start 16 banks apart -> 2 rows x 64 B per LDS wavefront, no conflicts
constexpr int BK_STAGES = 3;
constexpr int BK_STAGE_BYTES = BK_ROWS * BK_PITCH;
constexpr int BK_SMEM = BK_STAGES * BK_STAGE_BYTES + 128;

__device__ __forceinline__ void bk_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void bk_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bk_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "BK_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra BK_DONE;\n\t"
      "bra BK_WAIT;\n\t"
      "BK_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

#ifdef D3D_SK_STAMPS
// debug build only (make EXTRA=-DD3D_SK_STAMPS): clock stamps of CTA 0 (consumer warp 0 / the producer), read by tools/skinny_stamps.py
__device__ long long g_sk_stamps[128];
__device__ unsigned long long g_sk_cta_times[8][320][4];  // [launch % 8][CTA]: globaltimer at CTA start, after pdl_wait, at consumer end; smid
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define SK_GT(slot) do { if ((threadIdx.x & 31) == 0 && (threadIdx.x == 0 || (slot) == 0) && blockIdx.x < 320) g_sk_cta_times[nx.stamp_slot & 7][blockIdx.x][(slot)] = gtime(); } while (0)
#define SK_STAMP(cond, slot) do { if (cond) g_sk_stamps[(slot)] = clock64(); } while (0)
#else
#define SK_STAMP(cond, slot) do { } while (0)
#define SK_GT(slot) do { } while (0)
#endif

template <bool HI>
__global__ void __launch_bounds__(288) skinny_bulk_kernel(const uint16_t* __restrict__ A, long long lda, const uint16_t* __restrict__ W, long long ldw,
                                                          int M, int N, int K, int kind, SkinnyEpi ep, NextWeights nx) {
  extern __shared__ __align__(128) uint8_t bk_smem[];
  __shared__ float red[SK_WARPS][2][16][8];
  const uint32_t sbase = smem_u32(bk_smem);
  const uint32_t bar0 = sbase + BK_STAGES * BK_STAGE_BYTES;  // full[s] at +8s, empty[s] at +8(STAGES+s)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = (K + BK_KC - 1) / BK_KC;
  // Every CTA owns a contiguous range of N / gridDim.x (+-1) weight rows, walked in tiles of <= 16 rows: all CTAs stream the same number
  // of bytes.  (Whole 16-row tiles dealt round-robin left 44 SMs with two tiles and 104 with one on the N = 3072 outputs, and a last
  // round with 136 of 296 CTAs busy on the gate/up matrix.)  Range ends are even: a SwiGLU column pair never straddles two tiles.
  const int r0 = (int)(((long long)N * blockIdx.x / gridDim.x) & ~1LL);
  const int r1 = blockIdx.x + 1 == gridDim.x ? N : (int)(((long long)N * (blockIdx.x + 1) / gridDim.x) & ~1LL);
  pdl_trigger();  // the next kernel of the chain may become resident and start on ITS weights
  SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0, 0);
  if (threadIdx.x == 0) { SK_GT(0); }
#ifdef D3D_SK_STAMPS
  if (threadIdx.x == 0 && blockIdx.x < 320) { unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid)); g_sk_cta_times[nx.stamp_slot & 7][blockIdx.x][3] = smid; }
#endif
  if (threadIdx.x == 0) {
    for (int s = 0; s < BK_STAGES; ++s) { bk_mbar_init(bar0 + 8 * s, 1); bk_mbar_init(bar0 + 8 * (BK_STAGES + s), SK_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // persistent: the CTA walks row tiles blockIdx.x, +gridDim.x, ...; the ring position runs on across tiles, so the producer keeps
  // ~96 KB of weight rows in flight while the consumers reduce and store the previous tile
  if (warp == SK_WARPS) {
    // ===== producer: one row copy per lane =====
    int it = 0;
    for (int n_base = r0; n_base < r1; n_base += BK_ROWS) {
      const int rows = min(BK_ROWS, r1 - n_base);
      for (int ks = 0; ks < n_stages; ++ks, ++it) {
        const int s = it % BK_STAGES;
        const uint32_t ph = (uint32_t)(it / BK_STAGES) & 1u;
        bk_mbar_wait(bar0 + 8 * (BK_STAGES + s), ph ^ 1u);
        SK_STAMP(blockIdx.x == 0 && lane == 0 && it < 12, 64 + it);
        const int k0 = ks * BK_KC;
        const uint32_t bytes = (uint32_t)min(BK_KC, K - k0) * 2u;
        if (lane == 0) bk_mbar_expect_tx(bar0 + 8 * s, bytes * (uint32_t)rows);
        __syncwarp();
        if (lane < rows) bulk_g2s(sbase + s * BK_STAGE_BYTES + lane * BK_PITCH, W + (long long)((nx.dbg_same_rows ? 0 : n_base) + lane) * ldw + k0, bytes, bar0 + 8 * s);
      }
    }
    l2_prefetch_share(nx, lane);
  } else {
    // ===== consumers: warp w owns K elements [128w, 128w + 128) of every stage =====
    pdl_wait();  // A (and the residual / C rows) belong to the kernels before this one; the producer above never touches them
    SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0, 1);
    SK_GT(1);
    const int g = lane >> 2, kq = lane & 3;
    const bool lo_ok = g < M, hi_ok = HI && g + 8 < M;
    const uint16_t* a_lo_p = A + (long long)(lo_ok ? g : 0) * lda + warp * (BK_U * 32) + kq * 8;
    const uint16_t* a_hi_p = A + (long long)(hi_ok ? g + 8 : 0) * lda + warp * (BK_U * 32) + kq * 8;
    // A fragments come from L2 (~700 cycles): the fragments of stage it + 1 are requested before stage it is waited for, so that round trip
    // runs under a stage's worth of work instead of in front of every stage (clock stamps: 1 900 -> ~? cycles per 32 KB stage)
    const int total = ((r1 - r0 + BK_ROWS - 1) / BK_ROWS) * n_stages;
    uint4 al[BK_U], ah[BK_U], nl[BK_U], nh[BK_U];
    auto load_a = [&](int it_, uint4 (&l)[BK_U], uint4 (&h)[BK_U]) {
      const int ks_ = it_ % n_stages;
      const int kb = ks_ * BK_KC + warp * (BK_U * 32);
#pragma unroll
      for (int u = 0; u < BK_U; ++u) {
        const bool in = it_ < total && kb + u * 32 + kq * 8 < K;
        l[u] = (in && lo_ok) ? __ldg(reinterpret_cast<const uint4*>(a_lo_p + ks_ * BK_KC + u * 32)) : make_uint4(0, 0, 0, 0);
        h[u] = (HI && in && hi_ok) ? __ldg(reinterpret_cast<const uint4*>(a_hi_p + ks_ * BK_KC + u * 32)) : make_uint4(0, 0, 0, 0);
      }
    };
    load_a(0, al, ah);
    float acc[4][4];
    // one stage: `cl / ch` hold its A fragments, `nl_ / nh_` receive the next stage's (the caller alternates the two register sets, so no
    // register copy -- which would wait for the load -- sits at the end of a stage)
    auto do_stage = [&](int it, uint4 (&cl)[BK_U], uint4 (&ch)[BK_U], uint4 (&nl_)[BK_U], uint4 (&nh_)[BK_U]) {
      const int tile = it / n_stages, ks = it - tile * n_stages;
      const int n_base = r0 + tile * BK_ROWS;
      const int rows = min(BK_ROWS, r1 - n_base);
      const bool ok0 = g < rows, ok1 = g + 8 < rows;
      if (ks == 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
      }
      const int s = it % BK_STAGES;
      const uint32_t ph = (uint32_t)(it / BK_STAGES) & 1u;
      const int k0 = ks * BK_KC + warp * (BK_U * 32);
      SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0 && it < 12, 100 + it);
      load_a(it + 1, nl_, nh_);
      SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0 && it < 24, 2 + 2 * it);
      bk_mbar_wait(bar0 + 8 * s, ph);
      SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0 && it < 24, 3 + 2 * it);
      const uint32_t st = sbase + s * BK_STAGE_BYTES + (uint32_t)(warp * (BK_U * 64) + kq * 16);
#pragma unroll
      for (int u = 0; u < BK_U; ++u) {
        const bool in = k0 + u * 32 + kq * 8 < K;
        uint4 w0 = make_uint4(0, 0, 0, 0), w1 = make_uint4(0, 0, 0, 0);
        if (in && ok0) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0.x), "=r"(w0.y), "=r"(w0.z), "=r"(w0.w) : "r"(st + g * BK_PITCH + u * 64));
        if (in && ok1) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w1.x), "=r"(w1.y), "=r"(w1.z), "=r"(w1.w) : "r"(st + (g + 8) * BK_PITCH + u * 64));
        // four independent accumulator chains (the two 16-element halves of a fragment go to separate sums, added at the tile's end)
        mma16816(acc[0], cl[u].x, ch[u].x, cl[u].y, ch[u].y, w0.x, w0.y, kind);
        mma16816(acc[1], cl[u].x, ch[u].x, cl[u].y, ch[u].y, w1.x, w1.y, kind);
        mma16816(acc[2], cl[u].z, ch[u].z, cl[u].w, ch[u].w, w0.z, w0.w, kind);
        mma16816(acc[3], cl[u].z, ch[u].z, cl[u].w, ch[u].w, w1.z, w1.w, kind);
      }
      SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0 && it < 12, 76 + it);
      __syncwarp();
      if (lane == 0) bk_mbar_arrive(bar0 + 8 * (BK_STAGES + s));
      SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0 && it < 12, 88 + it);
      if (ks == n_stages - 1) {  // end of a row tile
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          red[warp][t][g][kq * 2] = acc[t][0] + acc[t + 2][0];
          red[warp][t][g][kq * 2 + 1] = acc[t][1] + acc[t + 2][1];
          red[warp][t][g + 8][kq * 2] = acc[t][2] + acc[t + 2][2];
          red[warp][t][g + 8][kq * 2 + 1] = acc[t][3] + acc[t + 2][3];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 consumer warps only: the producer keeps streaming the next tile
        SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0, 52 + (tile ? 3 : 0));
        skinny_epilogue<2>(red, n_base, M, r1, ep);  // columns >= r1 belong to the next CTA
        SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0, 53 + (tile ? 3 : 0));
        asm volatile("bar.sync 1, 256;" ::: "memory");  // `red` is rewritten by the next tile
        SK_STAMP(blockIdx.x == 0 && threadIdx.x == 0, 54 + (tile ? 3 : 0));
      }
    };
    for (int it = 0; it < total; it += 2) {
      do_stage(it, al, ah, nl, nh);
      if (it + 1 < total) do_stage(it + 1, nl, nh, al, ah);
    }
    SK_GT(2);
    if (ep.norm_w) {  // kernel-uniform
      // every CTA (also one with an empty row range) arrives once its rows of C are written; the last one normalises all M rows
      __shared__ int s_last;
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) s_last = atomicAdd(ep.norm_counter, 1) == (int)gridDim.x - 1;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_last) {
        __threadfence();
        if (threadIdx.x == 0) *ep.norm_counter = 0;  // ready for the next launch (stream order)
        for (int row = warp; row < M; row += SK_WARPS)
          warp_rmsnorm_row<true>((const float*)ep.C + (long long)row * ep.ldc, ep.norm_w, ep.norm_eps, N, ep.norm_out + (long long)row * N, ep.norm_kind, lane);
      }
    }
  }
}

// ---- decode attention over the per-layer QKV cache ----
// cache row layout: [q (H*Dh) | k (H*Dh) | v (H*Dh)] 16-bit; sequence b owns the prefill rows [cu[b], cu[b+1]) and the decode rows
// t_prefill + s*n_seq + b for s = 0..step (the last one is the query's own row).  Dh % 32 == 0 (96 or 64), Dh <= 128.
constexpr int DA_WARPS = 16;
constexpr int DA_U = 8;  // keys whose K and V rows a warp loads back to back (bytes in flight: the kernel is HBM-bound)
template <int DH>
__global__ void __launch_bounds__(DA_WARPS * 32, 1) decode_attn_kernel(const uint16_t* __restrict__ qkv, long long ld, const int* __restrict__ cu,
                                                                   int n_seq, int t_prefill, int step, int H, int kind, float scale,
                                                                   uint16_t* __restrict__ out, long long ldo) {
  constexpr int ACTIVE = DH / 4;  // lanes that hold 4 consecutive channels (8-byte loads); the other lanes idle (head_dim 96: 24 of 32)
  __shared__ float s_m[DA_WARPS], s_l[DA_WARPS], s_acc[DA_WARPS][DH];
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool on = lane < ACTIVE;
  const int p0 = cu[b], n_pre = cu[b + 1] - p0;
  const int n_keys = n_pre + step + 1;
  const long long q_row = (long long)t_prefill + (long long)step * n_seq + b;
  auto load4 = [&](long long row, long long col, float (&f)[4]) {
    const uint2 v = on ? *reinterpret_cast<const uint2*>(qkv + row * ld + col + lane * 4) : make_uint2(0, 0);
    const float2 a = unpack16x2(v.x, kind), c = unpack16x2(v.y, kind);
    f[0] = a.x; f[1] = a.y; f[2] = c.x; f[3] = c.y;
  };
  float q[4];
  load4(q_row, (long long)h * DH, q);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] *= scale;
  float m = -INFINITY, l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  const long long k_col = (long long)(H + h) * DH, v_col = (long long)(2 * H + h) * DH;
  for (int j0 = warp * DA_U; j0 < n_keys; j0 += DA_WARPS * DA_U) {
    float kf[DA_U][4], vf[DA_U][4], s[DA_U];
#pragma unroll
    for (int u = 0; u < DA_U; ++u) {
      const int j = min(j0 + u, n_keys - 1);  // clamped: the duplicate is masked below
      const long long row = j < n_pre ? (long long)(p0 + j) : (long long)t_prefill + (long long)(j - n_pre) * n_seq + b;
      load4(row, k_col, kf[u]);
      load4(row, v_col, vf[u]);
    }
#pragma unroll
    for (int u = 0; u < DA_U; ++u) s[u] = (q[0] * kf[u][0] + q[1] * kf[u][1]) + (q[2] * kf[u][2] + q[3] * kf[u][3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < DA_U; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    float mx = m;
#pragma unroll
    for (int u = 0; u < DA_U; ++u) {
      if (j0 + u >= n_keys) s[u] = -INFINITY;
      mx = fmaxf(mx, s[u]);
    }
    const float c = __expf(m - mx);  // 0 on the first batch (m = -inf)
    l *= c;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] *= c;
#pragma unroll
    for (int u = 0; u < DA_U; ++u) {
      const float p = __expf(s[u] - mx);
      l += p;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += p * vf[u][i];
    }
    m = mx;
  }
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  if (on) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s_acc[warp][lane * 4 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < DH) {
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) gm = fmaxf(gm, s_m[w]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < DA_WARPS; ++w) {
      const float c = s_m[w] == -INFINITY ? 0.f : __expf(s_m[w] - gm);
      num += c * s_acc[w][threadIdx.x];
      den += c * s_l[w];
    }
    st16(out, (size_t)((long long)b * ldo + (long long)h * DH + threadIdx.x), num / den, kind);
  }
}

// Staged decode attention.  Every warp streams ITS key blocks (16 keys = K and V head slices, 4*DH bytes per key) through a private double
// buffer in shared memory with 16-byte cp.async -- 12 KB in flight per warp instead of the 3 KB the register-staged kernel can hold, 192 KB
// per SM at two CTAs.  (cp.async.bulk per 192-byte row piece was tried and lost: ~3 000 bulk requests per CTA saturate the copy engine.)
// The arithmetic of a block runs on the tensor cores in TRANSPOSED form, so that the 16 keys -- not the single query -- fill the M side:
//   S^T[16 keys x 8] = K[16 x DH] . q^T[DH x 8]     (column 0 = the query, columns 1..7 zero)      DH/16 mma.m16n8k16
//   O^T[DH x 8]     += V^T[DH x 16 keys] . P^T[16 x 8]                                             DH/16 mma.m16n8k16
// K fragments by ldmatrix, V^T fragments by ldmatrix.trans (rows padded to 4*DH + 16 bytes: conflict-free), fp32 accumulate, P rounded to
// 16 bit like in the prefill kernels.  The CUDA-core version (4 channels per lane, shuffle-reduced dot products) issued ~500 instructions
// per block and warp: with 16 warps per SM the kernel was bound by issue slots, ~3 400 cycles per block, whether or not the keys came
// from HBM (clock stamps, tools/decode_attn_stamps.py).
// Under programmatic dependent launch a warp issues its first two blocks -- OLD cache rows, which nothing in flight writes -- before
// griddepcontrol.wait, i.e. while this step's QKV GEMM / RoPE still run; only the block with the newest key and the query row wait.
constexpr int DC_WARPS = 8;
constexpr int DC_KEYS = 16;
__device__ __forceinline__ void dc_cp_async16(uint32_t dst, const void* src, int src_bytes = 16) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void dc_ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void dc_ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <int DH>
__global__ void __launch_bounds__(DC_WARPS * 32) decode_attn_staged_kernel(uint16_t* qkv, long long ld, const int* __restrict__ cu, int n_seq,
                                                                           int t_prefill, int step, int H, int kind, float scale,
                                                                           uint16_t* __restrict__ out, long long ldo, const float* __restrict__ rope_tab,
                                                                           int dbg_same_block) {
  static_assert(DH % 16 == 0, "head_dim % 16");
  constexpr int ROWB = DH * 4;            // K then V head slice of one key, bytes
  constexpr int PITCH = ROWB + 16;        // shared-memory row: 8 consecutive rows start 4 banks apart (ldmatrix conflict-free)
  constexpr int CPK = ROWB / 16;          // 16-byte chunks per key
  constexpr int HALF = CPK / 2;           // ... of which the first half is K
  constexpr int BLOCK_BYTES = DC_KEYS * PITCH;
  constexpr int KS = DH / 16;             // k-steps of S^T = m-tiles of O^T
  extern __shared__ __align__(128) uint8_t dc_smem[];
  __shared__ float s_m[DC_WARPS], s_l[DC_WARPS], s_acc[DC_WARPS][DH];
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
#ifdef D3D_SK_STAMPS
  const bool st_on = b == 0 && h == 0 && threadIdx.x == 0;
  if (st_on) g_sk_stamps[0] = clock64();
  if (threadIdx.x == 0) g_sk_cta_times[0][(b * gridDim.y + h) % 320][0] = gtime();
#endif
  const int p0 = cu[b], n_pre = cu[b + 1] - p0;  // cu_seqlens: uploaded before the prefill, constant during generation
  const int n_keys = n_pre + step + 1;
  const int n_blk = (n_keys + DC_KEYS - 1) / DC_KEYS;
  const uint32_t wbase = smem_u32(dc_smem) + (uint32_t)warp * 2u * BLOCK_BYTES;
  const long long k_col = (long long)(H + h) * DH, v_col = (long long)(2 * H + h) * DH;
  bool waited = false;
  // Chunk c = lane + 32 i of a block is (key c / CPK, 16-byte piece c % CPK): the pattern repeats every PER iterations (= KPP keys), so the
  // per-lane pieces are three constants and a block is 12 address adds -- computed per chunk (divisions, row select, 64-bit multiplies)
  // the issue of one block cost ~2 800 cycles with 16 warps per SM
  constexpr int PER = CPK == 24 ? 3 : 1;
  constexpr int KPP = PER * 32 / CPK;  // keys per pattern: 4 (head_dim 96), 2 (64), 1 (128)
  static_assert(DC_KEYS % KPP == 0 && (PER * 32) % CPK == 0, "chunk pattern");
  uint32_t doff[PER];
  long long soff[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i, key = c / CPK, ch = c - key * CPK;
    doff[i] = (uint32_t)(key * PITCH + ch * 16);
    soff[i] = (long long)key * ld + (ch < HALF ? k_col + ch * 8 : v_col + (ch - HALF) * 8);
  }
  auto issue = [&](int blk, int buf) {  // one commit group per call (possibly empty)
    if (blk < n_blk) {
      if (blk == n_blk - 1 && !waited) { pdl_wait(); waited = true; }  // the newest key row comes from this step's QKV GEMM + RoPE
      const int j0 = blk * DC_KEYS;
      const uint32_t dst0 = wbase + (uint32_t)(buf * BLOCK_BYTES);
      if (j0 + DC_KEYS <= n_pre && !dbg_same_block) {  // whole block inside the (contiguous) prefill rows
        const uint16_t* src0 = qkv + (long long)(p0 + j0) * ld;
#pragma unroll
        for (int gk = 0; gk < DC_KEYS / KPP; ++gk)
#pragma unroll
          for (int i = 0; i < PER; ++i) dc_cp_async16(dst0 + (uint32_t)(gk * KPP * PITCH) + doff[i], src0 + (long long)(gk * KPP) * ld + soff[i]);
      } else {
        for (int c = lane; c < DC_KEYS * CPK; c += 32) {
          const int key = c / CPK, ch = c - key * CPK;
          int j = j0 + key;
          const bool ok = j < n_keys;  // rows past the end are zero-filled: 0 * stale shared memory must stay finite
          if (!ok || dbg_same_block) j = ok ? key : 0;
          const long long row = j < n_pre ? (long long)(p0 + j) : (long long)t_prefill + (long long)(j - n_pre) * n_seq + b;
          const uint16_t* src = qkv + row * ld + (ch < HALF ? k_col + ch * 8 : v_col + (ch - HALF) * 8);
          dc_cp_async16(dst0 + (uint32_t)(key * PITCH + ch * 16), src, ok ? 16 : 0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(warp, 0);
  issue(warp + DC_WARPS, 1);
  if (!waited) { pdl_wait(); waited = true; }
#ifdef D3D_SK_STAMPS
  if (st_on) g_sk_stamps[1] = clock64();
  if (threadIdx.x == 0) g_sk_cta_times[0][(b * gridDim.y + h) % 320][1] = gtime();
#endif
  const int g = lane >> 2, t = lane & 3;
  // B fragments of q^T (column 0 only: lanes with g == 0): b0 = dims 16 ks + 2t, +1; b1 = dims 16 ks + 2t + 8, +9
  uint32_t qb[KS][2];
  {
    const uint16_t* qp = qkv + ((long long)t_prefill + (long long)step * n_seq + b) * ld + (long long)h * DH;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      qb[ks][0] = g == 0 ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 2 * t) : 0u;
      qb[ks][1] = g == 0 ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 2 * t + 8) : 0u;
    }
  }
  // RoPE of the step's own row, fused (rope_tab != NULL: the row arrives un-rotated from the QKV GEMM): channel i pairs with i + DH/2, i.e.
  // k-step ks with ks + KS/2 in the SAME lane; the arithmetic and 16-bit rounding of rope_apply_kernel (nn_kernels.cu), so the cache row
  // equals what a prefill would have written.  q is rotated in registers; the key is rotated by the warp that owns the last block (below).
  const float* tb = rope_tab ? rope_tab + (long long)b * DH : nullptr;  // [cos (DH/2) | sin (DH/2)] of this sequence's position
  auto rot = [&](uint32_t& lo, uint32_t& hi, int i0) {  // channels (i0, i0 + 1) and their partners
    const float2 a = unpack16x2(lo, kind), bb = unpack16x2(hi, kind);
    const float c0 = tb[i0], c1 = tb[i0 + 1], s0 = tb[DH / 2 + i0], s1 = tb[DH / 2 + i0 + 1];
    lo = pack16x2(__fsub_rn(__fmul_rn(a.x, c0), __fmul_rn(bb.x, s0)), __fsub_rn(__fmul_rn(a.y, c1), __fmul_rn(bb.y, s1)), kind);
    hi = pack16x2(__fadd_rn(__fmul_rn(bb.x, c0), __fmul_rn(a.x, s0)), __fadd_rn(__fmul_rn(bb.y, c1), __fmul_rn(a.y, s1)), kind);
  };
  if (tb && g == 0) {
#pragma unroll
    for (int ks = 0; ks < KS / 2; ++ks) {
      rot(qb[ks][0], qb[ks + KS / 2][0], ks * 16 + 2 * t);
      rot(qb[ks][1], qb[ks + KS / 2][1], ks * 16 + 2 * t + 8);
    }
  }
  float m = -INFINITY, l = 0.f;
  float o[KS][4];  // O^T tiles: [0] / [2] in lanes t == 0 are channels 16 mt + g / + g + 8
#pragma unroll
  for (int i = 0; i < KS; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  int buf = 0;
  for (int blk = warp; blk < n_blk; blk += DC_WARPS, buf ^= 1) {
#ifdef D3D_SK_STAMPS
    if (st_on && blk / DC_WARPS < 10) g_sk_stamps[2 + 3 * (blk / DC_WARPS)] = clock64();
#endif
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
#ifdef D3D_SK_STAMPS
    if (st_on && blk / DC_WARPS < 10) g_sk_stamps[3 + 3 * (blk / DC_WARPS)] = clock64();
#endif
    const uint32_t st = wbase + (uint32_t)(buf * BLOCK_BYTES);
    const int nvalid = min(DC_KEYS, n_keys - blk * DC_KEYS);
    if (tb && blk == n_blk - 1) {  // the newest key: rotate it in shared memory and in the cache row (for the steps to come)
      if (lane < DH / 4) {
        const uint32_t ka = st + (uint32_t)((nvalid - 1) * PITCH + lane * 4);
        uint32_t lo, hi;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(ka));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(ka + DH));
        rot(lo, hi, 2 * lane);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ka), "r"(lo) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ka + DH), "r"(hi) : "memory");
        uint16_t* kr = qkv + ((long long)t_prefill + (long long)step * n_seq + b) * ld + k_col + 2 * lane;
        *reinterpret_cast<uint32_t*>(kr) = lo;
        *reinterpret_cast<uint32_t*>(kr + DH / 2) = hi;
      }
      __syncwarp();
    }
    // ---- S^T = K q^T ----
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const uint32_t ka = st + (uint32_t)((lane & 15) * PITCH + (lane >> 4) * 16);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a0, a1, a2, a3;
        dc_ldsm_x4(ka + ks * 32, a0, a1, a2, a3);
        mma16816(sc, a0, a1, a2, a3, qb[ks][0], qb[ks][1], kind);
      }
    }
    // column 0 lives in the lanes t == 0: sc[0] = key g, sc[2] = key g + 8; broadcast over the 4 lanes of a group, then reduce over g
    float s_lo = __shfl_sync(0xffffffffu, sc[0], lane & ~3) * scale, s_hi = __shfl_sync(0xffffffffu, sc[2], lane & ~3) * scale;
    if (g >= nvalid) s_lo = -INFINITY;
    if (g + 8 >= nvalid) s_hi = -INFINITY;
    float mx = fmaxf(s_lo, s_hi);
#pragma unroll
    for (int ofs = 4; ofs < 32; ofs <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, ofs));
    const float mn = fmaxf(m, mx);  // finite: every block has at least one valid key
    const float corr = __expf(m - mn);  // 0 on the first block (m = -inf)
    const float p_lo = __expf(s_lo - mn), p_hi = __expf(s_hi - mn);
    float ps = p_lo + p_hi;
#pragma unroll
    for (int ofs = 4; ofs < 32; ofs <<= 1) ps += __shfl_xor_sync(0xffffffffu, ps, ofs);
    l = l * corr + ps;
    m = mn;
    // B fragments of P^T (column 0: lanes g == 0): b0 = keys 2t, 2t+1; b1 = keys 2t+8, 2t+9
    const float pa = __shfl_sync(0xffffffffu, p_lo, 8 * t), pb = __shfl_sync(0xffffffffu, p_lo, 8 * t + 4);
    const float pc = __shfl_sync(0xffffffffu, p_hi, 8 * t), pd = __shfl_sync(0xffffffffu, p_hi, 8 * t + 4);
    const uint32_t pb0 = g == 0 ? pack16x2(pa, pb, kind) : 0u, pb1 = g == 0 ? pack16x2(pc, pd, kind) : 0u;
    // ---- O^T = O^T * corr + V^T P^T ----
    {
      const uint32_t va = st + (uint32_t)(((lane & 7) + ((lane >> 4) << 3)) * PITCH + DH * 2 + ((lane >> 3) & 1) * 16);
#pragma unroll
      for (int mt = 0; mt < KS; ++mt) {
        uint32_t a0, a1, a2, a3;
        dc_ldsm_x4_trans(va + mt * 32, a0, a1, a2, a3);
        o[mt][0] *= corr; o[mt][1] *= corr; o[mt][2] *= corr; o[mt][3] *= corr;
        mma16816(o[mt], a0, a1, a2, a3, pb0, pb1, kind);
      }
    }
    __syncwarp();  // every lane is done with this buffer before it is refilled
#ifdef D3D_SK_STAMPS
    if (st_on && blk / DC_WARPS < 10) g_sk_stamps[4 + 3 * (blk / DC_WARPS)] = clock64();
#endif
    issue(blk + 2 * DC_WARPS, buf);
  }
#ifdef D3D_SK_STAMPS
  if (st_on) g_sk_stamps[40] = clock64();
#endif
  if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
  if (t == 0) {
#pragma unroll
    for (int mt = 0; mt < KS; ++mt) {
      s_acc[warp][mt * 16 + g] = o[mt][0];
      s_acc[warp][mt * 16 + g + 8] = o[mt][2];
    }
  }
  __syncthreads();
#ifdef D3D_SK_STAMPS
  if (st_on) g_sk_stamps[41] = clock64();
  if (threadIdx.x == 0) g_sk_cta_times[0][(b * gridDim.y + h) % 320][2] = gtime();
#endif
  if (threadIdx.x < DH) {
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < DC_WARPS; ++w) gm = fmaxf(gm, s_m[w]);
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int w = 0; w < DC_WARPS; ++w) {
      const float c = s_m[w] == -INFINITY ? 0.f : __expf(s_m[w] - gm);
      num += c * s_acc[w][threadIdx.x];
      den += c * s_l[w];
    }
    st16(out, (size_t)((long long)b * ldo + (long long)h * DH + threadIdx.x), num / den, kind);
  }
}

__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ x, long long ld, int n, int* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  pdl_trigger();
  pdl_wait();
  const float* r = x + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = r[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
  }
}

// embed the fed tokens (fp32 residual rows) and set their positions: pos[b] = prefill length of b + step
__global__ void decode_prep_kernel(const void* __restrict__ table, int kind, const int* __restrict__ ids, const int* __restrict__ cu, int step, int D,
                                   float* __restrict__ x, int* __restrict__ pos) {
  const int b = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) pos[b] = cu[b + 1] - cu[b] + step;
  const size_t src = (size_t)ids[b] * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) x[(long long)b * D + c] = ld16(table, src + c, kind);
}

// rows of the step's tokens: the same arithmetic (and summation order) as rmsnorm_kernel / rope_table_kernel / rope_apply_kernel of
// nn_kernels.cu, which the prefill uses -- a cache row written by a decode step is bit-identical to the one a prefill would write
__global__ void __launch_bounds__(256) dec_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, float eps, int T, int D,
                                                          uint16_t* __restrict__ out16, int kind) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= T) return;
  warp_rmsnorm_row<false>(x + (long long)row * D, w, eps, D, out16 + (long long)row * D, kind, lane);
}

__global__ void dec_rope_table_kernel(const int* __restrict__ pos, const float* __restrict__ inv_freq, int T, int half, float* __restrict__ tab) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * half) return;
  const int t = idx / half, i = idx - t * half;
  float sn, cs;
  sincosf((float)pos[t] * inv_freq[i], &sn, &cs);
  tab[(size_t)t * 2 * half + i] = cs;
  tab[(size_t)t * 2 * half + half + i] = sn;
}

__global__ void dec_rope_apply_kernel(uint16_t* __restrict__ qkv, long long ld, const float* __restrict__ tab, int T, int H, int Dh, int kind) {
  pdl_trigger();
  pdl_wait();
  const int half = Dh / 2, groups = half / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * 2 * H * groups) return;
  const int g = (int)(idx % groups);
  const int h = (int)((idx / groups) % (2 * H));  // q heads then k heads
  const int t = (int)(idx / ((long long)groups * 2 * H));
  uint16_t* base = qkv + (size_t)t * ld + (size_t)h * Dh + g * 8;
  const float4* cs = reinterpret_cast<const float4*>(tab + (size_t)t * Dh + g * 8);
  const float4* sn = reinterpret_cast<const float4*>(tab + (size_t)t * Dh + half + g * 8);
  const float4 c0 = cs[0], c1 = cs[1], s0 = sn[0], s1 = sn[1];
  const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const uint4 av = *reinterpret_cast<const uint4*>(base), bv = *reinterpret_cast<const uint4*>(base + half);
  const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
  uint32_t ao[4], bo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = unpack16x2(aw[j], kind), b = unpack16x2(bw[j], kind);
    ao[j] = pack16x2(__fsub_rn(__fmul_rn(a.x, c[2 * j]), __fmul_rn(b.x, s[2 * j])), __fsub_rn(__fmul_rn(a.y, c[2 * j + 1]), __fmul_rn(b.y, s[2 * j + 1])), kind);
    bo[j] = pack16x2(__fadd_rn(__fmul_rn(b.x, c[2 * j]), __fmul_rn(a.x, s[2 * j])), __fadd_rn(__fmul_rn(b.y, c[2 * j + 1]), __fmul_rn(a.y, s[2 * j + 1])), kind);
  }
  *reinterpret_cast<uint4*>(base) = make_uint4(ao[0], ao[1], ao[2], ao[3]);
  *reinterpret_cast<uint4*>(base + half) = make_uint4(bo[0], bo[1], bo[2], bo[3]);
}

int g_skinny_cfg = 0;
struct FusedNorm {
  const float* w; uint16_t* out; float eps; int* counter;
};
int g_decode_attn_impl = 1;  // 1: cp.async-staged, 0: register-staged (A/B)

int skinny(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K, int kind, int out_kind,
           const float* bias, int act, const float* residual, long long ldres, cudaStream_t st, NextWeights nx = NextWeights{nullptr, 0, 0, 0},
           const FusedNorm* norm = nullptr) {
  D3D_REQUIRE(M >= 1 && M <= 16, "skinny GEMM handles 1..16 activation rows");
  D3D_REQUIRE(K % 32 == 0 && lda % 8 == 0 && ldw % 8 == 0, "K must be a multiple of 32, rows 16-byte aligned");
  D3D_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "A/W must be 16-byte aligned");
  D3D_REQUIRE(kind == D3D_F16 || kind == D3D_BF16, "16-bit operands");
  D3D_REQUIRE(act != D3D_ACT_SWIGLU || ((N % 2) == 0 && residual == nullptr), "swiglu needs even N, no residual");
  D3D_REQUIRE(!norm || (out_kind == D3D_OUT_F32 && act == D3D_ACT_NONE && N % 128 == 0 && ldc % 4 == 0 && (g_skinny_cfg == 0 || g_skinny_cfg >= 10)), "fused norm: fp32 rows, N % 128, bulk kernel");
  SkinnyEpi ep{C, ldc, bias, residual, ldres, act, out_kind, nullptr, nullptr, 0.f, 0, nullptr};
  if (norm) { ep.norm_w = norm->w; ep.norm_out = norm->out; ep.norm_eps = norm->eps; ep.norm_kind = kind; ep.norm_counter = norm->counter; }
  const uint16_t* a = (const uint16_t*)A;
  const uint16_t* w = (const uint16_t*)W;
#define SK_LAUNCH(NT, U, NA)                                                                                                              \
  do {                                                                                                                                    \
    if (M > 8) skinny_gemm_kernel<true, NT, U, NA><<<d3d_cdiv(N, 8 * NT), SK_WARPS * 32, 0, st>>>(a, lda, w, ldw, M, N, K, kind, ep);      \
    else skinny_gemm_kernel<false, NT, U, NA><<<d3d_cdiv(N, 8 * NT), SK_WARPS * 32, 0, st>>>(a, lda, w, ldw, M, N, K, kind, ep);           \
  } while (0)
  switch (g_skinny_cfg) {  // 0 = default heuristic; 1.. = tuning variants (tools/skinny_bench.py)
    case 1: SK_LAUNCH(1, 2, true); break;
    case 2: SK_LAUNCH(2, 2, true); break;
    case 3: SK_LAUNCH(2, 4, true); break;
    case 4: SK_LAUNCH(4, 2, true); break;
    case 5: SK_LAUNCH(1, 4, true); break;
    case 6: SK_LAUNCH(1, 2, false); break;
    case 7: SK_LAUNCH(2, 2, false); break;
    case 8: SK_LAUNCH(4, 1, true); break;
    case 9: SK_LAUNCH(2, 1, true); break;
    case 10:
    case 11:
    bulk: {  // bulk-copy (cp.async.bulk + mbarrier ring) variant
      static bool attr = false;
      if (!attr) {
        D3D_CHECK_CUDA(cudaFuncSetAttribute(skinny_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM));
        D3D_CHECK_CUDA(cudaFuncSetAttribute(skinny_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM));
        attr = true;
      }
      // two resident CTAs per SM (cfg 11: one, so that the next kernel of a PDL chain is co-resident on every SM)
      const int cap = (g_skinny_cfg == 11 ? 1 : 2) * d3d_num_sms();
      const int tiles = d3d_cdiv(N, 8), grid = tiles < cap ? tiles : cap;  // at least ~8 rows per CTA
#ifdef D3D_SK_STAMPS
      static int launch_no = 0;
      nx.stamp_slot = launch_no++;
#endif
      if (M > 8) D3D_CHECK_CUDA(launch_pdl(skinny_bulk_kernel<true>, dim3(grid), dim3(288), BK_SMEM, st, a, lda, w, ldw, M, N, K, kind, ep, nx));
      else D3D_CHECK_CUDA(launch_pdl(skinny_bulk_kernel<false>, dim3(grid), dim3(288), BK_SMEM, st, a, lda, w, ldw, M, N, K, kind, ep, nx));
      break;
    }
    default: goto bulk;  // tools/skinny_bench.py: the persistent cp.async.bulk kernel streams fastest on every Phi-3 shape
  }
#undef SK_LAUNCH
  D3D_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int d3d_gemm_skinny_set_config(int cfg) { g_skinny_cfg = cfg; return 0; }

#ifdef D3D_SK_STAMPS
extern "C" int d3d_debug_skinny_stamps(long long* host_out) {
  D3D_CHECK_CUDA(cudaDeviceSynchronize());
  D3D_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_sk_stamps, sizeof(long long) * 128));
  return 0;
}
extern "C" int d3d_debug_skinny_cta_times(unsigned long long* host_out) {
  D3D_CHECK_CUDA(cudaDeviceSynchronize());
  D3D_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_sk_cta_times, sizeof(unsigned long long) * 8 * 320 * 4));
  return 0;
}
#endif

extern "C" int d3d_gemm_skinny(const d3d_gemm_args* a, void* stream) {
  D3D_REQUIRE(a != nullptr && a->A && a->W && a->C, "args");
  return skinny(a->A, a->lda, a->W, a->ldw, a->C, a->ldc, a->M, a->N, a->K, a->in_kind, a->out_kind, a->bias, a->act, a->residual, a->ldres,
                (cudaStream_t)stream);
}

extern "C" int d3d_argmax_rows(const float* x, int64_t ld, int rows, int n, int* out, void* stream) {
  if (rows == 0) return 0;
  D3D_REQUIRE(x && out && n > 0, "args");
  argmax_kernel<<<rows, 1024, 0, (cudaStream_t)stream>>>(x, ld, n, out);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_decode_attention(const void* qkv, int64_t ld, const int* cu_seqlens, int n_seq, int t_prefill, int step, int H, int Dh,
                                    int kind, float scale, void* out, int64_t ldo, void* stream) {
  return d3d_decode_attention_rope((void*)qkv, ld, cu_seqlens, n_seq, t_prefill, step, H, Dh, kind, scale, nullptr, out, ldo, stream);
}

// rope_tab != NULL ([n_seq, Dh] fp32: cos | sin of every sequence's position, d3d_rope_table): the step's own rows are still un-rotated; the
// kernel rotates q in registers and the new key in place (cache row), like d3d_rope_apply would have
extern "C" int d3d_decode_attention_rope(void* qkv, int64_t ld, const int* cu_seqlens, int n_seq, int t_prefill, int step, int H, int Dh, int kind,
                                         float scale, const float* rope_tab, void* out, int64_t ldo, void* stream) {
  if (n_seq == 0) return 0;
  D3D_REQUIRE(qkv && cu_seqlens && out && step >= 0, "args");
  D3D_REQUIRE(kind == D3D_F16 || kind == D3D_BF16, "16-bit cache");
  dim3 grid(n_seq, H);
  cudaStream_t st = (cudaStream_t)stream;
  const uint16_t* c = (const uint16_t*)qkv;
  const long long ldl = ld, ldol = ldo;
  if (g_decode_attn_impl == 0) {  // register-staged kernel (A/B)
    if (rope_tab) {
      const long long total = (long long)n_seq * 2 * H * (Dh / 16);
      D3D_CHECK_CUDA(launch_pdl(dec_rope_apply_kernel, dim3(d3d_cdiv(total, 256)), dim3(256), 0, st,
                                (uint16_t*)qkv + ((long long)t_prefill + (long long)step * n_seq) * ld, (long long)ld, rope_tab, n_seq, H, Dh, kind));
    }
    const dim3 blk(DA_WARPS * 32);
    if (Dh == 96) D3D_CHECK_CUDA(launch_pdl(decode_attn_kernel<96>, grid, blk, 0, st, c, ldl, cu_seqlens, n_seq, t_prefill, step, H, kind, scale, (uint16_t*)out, ldol));
    else if (Dh == 64) D3D_CHECK_CUDA(launch_pdl(decode_attn_kernel<64>, grid, blk, 0, st, c, ldl, cu_seqlens, n_seq, t_prefill, step, H, kind, scale, (uint16_t*)out, ldol));
    else if (Dh == 128) D3D_CHECK_CUDA(launch_pdl(decode_attn_kernel<128>, grid, blk, 0, st, c, ldl, cu_seqlens, n_seq, t_prefill, step, H, kind, scale, (uint16_t*)out, ldol));
    else { d3d_set_error("decode attention: head_dim %d not built (64, 96, 128)", Dh); return D3D_EINVAL; }
    D3D_CHECK_LAUNCH();
    return 0;
  }
  D3D_REQUIRE(ld % 8 == 0 && ((uintptr_t)qkv % 16) == 0, "16-byte aligned cache rows (cp.async)");
  const dim3 blk(DC_WARPS * 32);
#define DB_LAUNCH(DHV)                                                                                                                    \
  do {                                                                                                                                    \
    constexpr int SMEM = DC_WARPS * 2 * DC_KEYS * (DHV * 4 + 16);                                                                              \
    static bool attr = false;                                                                                                             \
    if (!attr) {                                                                                                                          \
      D3D_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_staged_kernel<DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));              \
      attr = true;                                                                                                                        \
    }                                                                                                                                     \
    D3D_CHECK_CUDA(launch_pdl(decode_attn_staged_kernel<DHV>, grid, blk, SMEM, st, (uint16_t*)qkv, ldl, cu_seqlens, n_seq, t_prefill, step, H, kind, \
                              scale, (uint16_t*)out, ldol, rope_tab, (g_decode_dbg >> 1) & 1));                                                                                  \
  } while (0)
  if (Dh == 96) DB_LAUNCH(96);
  else if (Dh == 64) DB_LAUNCH(64);
  else if (Dh == 128) DB_LAUNCH(128);
  else { d3d_set_error("decode attention: head_dim %d not built (64, 96, 128)", Dh); return D3D_EINVAL; }
#undef DB_LAUNCH
  D3D_CHECK_LAUNCH();
  return 0;
}

namespace {
// a zeroed int per (device, stream), owned by the library: the arrival counter of the fused norm (self-resetting)
int* stream_counter(cudaStream_t st) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, int*> m;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  int*& p = m[{dev, st}];
  if (!p) {
    if (cudaMalloc(&p, 64) != cudaSuccess) { p = nullptr; return nullptr; }
    if (cudaMemset(p, 0, 64) != cudaSuccess) return nullptr;
  }
  return p;
}
}  // namespace

// One greedy decode step for all sequences of the rank.  Feeds tokens_in (the previous step's arg-max), appends their q/k/v rows to the
// per-layer caches at row t_prefill + step*n_seq + b, and writes the next-token logits and their arg-max.
extern "C" int d3d_lm_decode_step(const d3d_lm_model* m, void* const* qkv_layers_h, int64_t ld_qkv, const int* cu_seqlens, int n_seq,
                                  int t_prefill, int step, const int* tokens_in, const float* inv_freq, float* x32, void* a16, void* att16,
                                  void* h16, float* rope_tab, int* pos, float* logits, int* next_tokens, void* stream) {
  D3D_REQUIRE(m && qkv_layers_h && cu_seqlens && tokens_in && inv_freq && x32 && a16 && att16 && h16 && rope_tab && pos && logits && next_tokens,
              "args");
  D3D_REQUIRE(n_seq >= 1 && n_seq <= 16, "1..16 sequences per decode step");
  cudaStream_t st = (cudaStream_t)stream;
  const int Dm = m->hidden, H = m->n_heads, Dh = m->head_dim, kind = m->kind;
  D3D_REQUIRE(H * Dh == Dm, "hidden = heads * head_dim");
  D3D_REQUIRE(Dm % 128 == 0 && Dh % 16 == 0 && ld_qkv % 8 == 0, "hidden % 128, head_dim % 16, 16-byte cache rows");
  const void* embed = m->embed;
  D3D_CHECK_CUDA(launch_pdl(decode_prep_kernel, dim3(n_seq), dim3(256), 0, st, embed, kind, tokens_in, cu_seqlens, step, Dm, x32, pos));
  D3D_CHECK_CUDA(launch_pdl(dec_rope_table_kernel, dim3(d3d_cdiv(n_seq * (Dh / 2), 256)), dim3(256), 0, st, (const int*)pos, inv_freq, n_seq, Dh / 2, rope_tab));
  const long long row0 = (long long)t_prefill + (long long)step * n_seq;
  const float scale = (float)(1.0 / sqrt((double)Dh));  // same rounding as the prefill (double expression rounded once)
  const dim3 norm_grid(d3d_cdiv(n_seq * 32, 256));
  auto rmsnorm = [&](const float* w) {
    return launch_pdl(dec_rmsnorm_kernel, norm_grid, dim3(256), 0, st, (const float*)x32, w, m->eps, n_seq, Dm, (uint16_t*)a16, kind);
  };
  // what the producers of each GEMM pull into L2 for the kernel after next (capped: two matrices plus a layer's K / V stay below the L2 size)
  const long long es = 2, cap = 48ll << 20;
  auto next = [&](const void* w, long long rows_, long long cols_) {
    const long long b = rows_ * cols_ * es;
    return NextWeights{g_decode_prefetch ? w : nullptr, g_decode_prefetch ? (b < cap ? b : cap) : 0, g_decode_dbg & 1, 0};
  };
  int* counter = stream_counter(st);
  D3D_REQUIRE(counter != nullptr, "arrival counter allocation");
  D3D_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));  // a step that died half-way must not poison the next one
  D3D_CHECK_CUDA(rmsnorm(m->layers[0].rms1));  // every later norm rides on the GEMM that completes the residual row
  for (int l = 0; l < m->n_layers; ++l) {
    const d3d_lm_layer& L = m->layers[l];
    uint16_t* rows = (uint16_t*)qkv_layers_h[l] + row0 * ld_qkv;
    const bool last = l + 1 == m->n_layers;
    const FusedNorm n2{L.rms2, (uint16_t*)a16, m->eps, counter};
    const FusedNorm n1{last ? m->norm : m->layers[last ? l : l + 1].rms1, (uint16_t*)a16, m->eps, counter};
    D3D_TRY(skinny(a16, Dm, L.w_qkv, Dm, rows, ld_qkv, n_seq, 3 * Dm, Dm, kind, kind, nullptr, D3D_ACT_NONE, nullptr, 0, st, next(L.w_o, Dm, Dm)));
    D3D_TRY(d3d_decode_attention_rope(qkv_layers_h[l], ld_qkv, cu_seqlens, n_seq, t_prefill, step, H, Dh, kind, scale, rope_tab, att16, Dm, stream));
    D3D_TRY(skinny(att16, Dm, L.w_o, Dm, x32, Dm, n_seq, Dm, Dm, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, x32, Dm, st, next(L.w_gu, 2 * m->ffn, Dm), &n2));
    D3D_TRY(skinny(a16, Dm, L.w_gu, Dm, h16, m->ffn, n_seq, 2 * m->ffn, Dm, kind, kind, nullptr, D3D_ACT_SWIGLU, nullptr, 0, st, next(L.w_down, Dm, m->ffn)));
    D3D_TRY(skinny(h16, m->ffn, L.w_down, m->ffn, x32, Dm, n_seq, Dm, m->ffn, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, x32, Dm, st,
                   last ? next(m->lm_head, m->vocab, Dm) : next(m->layers[l + 1].w_qkv, 3 * Dm, Dm), &n1));
  }
  D3D_TRY(skinny(a16, Dm, m->lm_head, Dm, logits, m->vocab, n_seq, m->vocab, Dm, kind, D3D_OUT_F32, nullptr, D3D_ACT_NONE, nullptr, 0, st));
  D3D_CHECK_CUDA(launch_pdl(argmax_kernel, dim3(n_seq), dim3(1024), 0, st, (const float*)logits, (long long)m->vocab, m->vocab, next_tokens));
  D3D_CHECK_LAUNCH();
  return 0;
}

// 0: plain stream-ordered launches (A/B of the programmatic-dependent-launch chain; tools/decode_bench.py)
extern "C" int d3d_lm_decode_set_pdl(int on) {
  g_decode_pdl = (on & 1) != 0;
  g_decode_attn_impl = (on & 2) ? 0 : 1;  // bit 1: the register-staged decode attention instead of the cp.async-staged one
  g_decode_prefetch = (on & 4) == 0;      // bit 2: no L2 prefetch of the next kernel's weights
  g_decode_dbg = (on >> 3) & 3;           // bits 3, 4: timing experiments (results are wrong)
  return 0;
}
