// Tensor-core flash attention (online softmax) over packed variable-length sequences: mma.sync m16n8k16, fp32 accumulate.
// One CTA = 128 queries of one (sequence, head): 8 warps x 16 query rows; K/V streamed in 64-key tiles through a
// double-buffered cp.async pipeline in shared memory (ldmatrix / ldmatrix.trans fragments); S = QK^T and O += PV stay in registers (P re-used as the A fragment).
// Used for CLIP ViT (CLIPM:181-183: 577 tokens, 16 heads x 64) and the Phi-3 prefill (causal, 32 heads x 96).
// TODO(next round): tcgen05 version (S and O in TMEM); this legacy-HMMA kernel is ~8 % of the step's FLOPs.
#include "common.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

namespace {

constexpr int NWARPS = 8;
constexpr int BQ = 16 * NWARPS;  // 128 query rows per CTA
constexpr int BKV = 64;
constexpr int NTHREADS = 32 * NWARPS;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  // 16-byte async copy; src_bytes = 0 zero-fills (rows past the end of the sequence)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <int KIND>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (KIND == D3D_BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

extern __shared__ __align__(16) uint16_t smem_attn[];

// one (sequence, head, 128-query tile); every branch on (seq, qt) is CTA-uniform
template <int D, int KIND>
__device__ __forceinline__ void attn_mma_tile(const uint16_t* __restrict__ qkv, long long ld, uint16_t* __restrict__ out, long long ldo,
                                              const int* __restrict__ cu, int H, int causal, float scale_log2, int seq, int h, int qt) {
  constexpr int LDS = D + 8;  // padded row (halves): 16-byte aligned rows, conflict-free ldmatrix
  constexpr int KS = D / 16;  // k-steps over head dim
  constexpr int DT = D / 8;   // output n-tiles
  uint16_t* sQ = smem_attn;                       // [BQ][LDS]
  uint16_t* sKV = smem_attn + BQ * LDS;           // 2 stages x {K [BKV][LDS], V [BKV][LDS]}

  const int b = cu[seq], len = cu[seq + 1] - b;
  const int q0 = qt * BQ;
  if (q0 >= len) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t qoff = (size_t)h * D, koff = (size_t)(H + h) * D, voff = (size_t)(2 * H + h) * D;
  constexpr int CPR = D / 8;  // 16-byte chunks per row

  for (int i = threadIdx.x; i < BQ * CPR; i += NTHREADS) {
    const int r = i / CPR, c = i % CPR;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q0 + r < len) v = *reinterpret_cast<const uint4*>(qkv + (size_t)(b + q0 + r) * ld + qoff + c * 8);
    *reinterpret_cast<uint4*>(&sQ[r * LDS + c * 8]) = v;
  }
  __syncthreads();
  // Q fragments of this warp's 16 rows
  uint32_t qf[KS][4];
  {
    const int row = warp * 16 + (lane & 15);
    const int colsel = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) ldsm_x4(smem_u32(&sQ[row * LDS + ks * 16 + colsel]), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }
  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;

  const int kmax = causal ? min(len, q0 + BQ) : len;
  const int n_tiles = (kmax + BKV - 1) / BKV;
  auto load_tile = [&](int tile, int stage) {
    const int c0 = tile * BKV;
    uint16_t* dK = sKV + stage * 2 * BKV * LDS;
    uint16_t* dV = dK + BKV * LDS;
    for (int i = threadIdx.x; i < BKV * CPR; i += NTHREADS) {
      const int r = i / CPR, c = i % CPR;
      const bool ok = c0 + r < len;
      const uint16_t* rowp = qkv + (size_t)(b + (ok ? c0 + r : 0)) * ld;
      cp_async16(smem_u32(&dK[r * LDS + c * 8]), rowp + koff + c * 8, ok ? 16 : 0);
      cp_async16(smem_u32(&dV[r * LDS + c * 8]), rowp + voff + c * 8, ok ? 16 : 0);
    }
    cp_async_commit();
  };
  load_tile(0, 0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int c0 = tile * BKV;
    if (tile + 1 < n_tiles) {
      load_tile(tile + 1, (tile + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint16_t* sK = sKV + (tile & 1) * 2 * BKV * LDS;
    const uint16_t* sV = sK + BKV * LDS;
    const bool active = !(causal && c0 > q0 + warp * 16 + 15);  // whole tile above the diagonal for this warp (warp-uniform)
    if (active) {
    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
        // x4: matrices (keys np*16+0..7, d ks*16+0..7), (same keys, d +8), (keys +8, d +0), (keys +8, d +8)
        const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int dcol = ks * 16 + (((lane >> 3) & 1) << 3);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(&sK[key * LDS + dcol]), b0, b1, b2, b3);
        mma16816<KIND>(s[2 * np], qf[ks], b0, b1);
        mma16816<KIND>(s[2 * np + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask + online softmax (base-2 exponent: scale_log2 = scale * log2(e)) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = c0 + nt * 8 + 2 * t + (e & 1);
        const int row = (e < 2) ? row0 : row1;
        const bool ok = key < len && (!causal || key <= row);
        const float v = ok ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float base0 = (mn0 == -INFINITY) ? 0.f : mn0, base1 = (mn1 == -INFINITY) ? 0.f : mn1;
    const float corr0 = exp2f(m0 - base0), corr1 = exp2f(m1 - base1);  // m = -inf -> 0
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - base0), p1 = exp2f(s[nt][1] - base0);
      const float p2 = exp2f(s[nt][2] - base1), p3 = exp2f(s[nt][3] - base1);
      rs0 += p0 + p1; rs1 += p2 + p3;
      const int kk = nt >> 1, hi = nt & 1;
      pf[kk][hi * 2 + 0] = pack16x2(p0, p1, KIND);
      pf[kk][hi * 2 + 1] = pack16x2(p2, p3, KIND);
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * corr0 + rs0; l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {  // pairs of 8-wide d tiles
        // trans x4: matrices (keys kk*16+0..7, d dp*16+0..7), (keys +8, same d), (keys +0, d +8), (keys +8, d +8)
        const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int dcol = dp * 16 + ((lane >> 4) << 3);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(smem_u32(&sV[key * LDS + dcol]), b0, b1, b2, b3);
        mma16816<KIND>(o[2 * dp], pf[kk], b0, b1);
        mma16816<KIND>(o[2 * dp + 1], pf[kk], b2, b3);
      }
    }
    }  // active
    __syncthreads();  // everyone is done with this stage before tile+2 overwrites it
  }
  const float inv0 = l0 > 0.f ? 1.0f / l0 : 0.f, inv1 = l1 > 0.f ? 1.0f / l1 : 0.f;
#pragma unroll
  for (int dt = 0; dt < DT; ++dt) {
    const int col = dt * 8 + 2 * t;
    if (row0 < len) *reinterpret_cast<uint32_t*>(out + (size_t)(b + row0) * ldo + (size_t)h * D + col) = pack16x2(o[dt][0] * inv0, o[dt][1] * inv0, KIND);
    if (row1 < len) *reinterpret_cast<uint32_t*>(out + (size_t)(b + row1) * ldo + (size_t)h * D + col) = pack16x2(o[dt][2] * inv1, o[dt][3] * inv1, KIND);
  }
}

template <int D, int KIND>
__global__ void __launch_bounds__(NTHREADS) attn_mma_kernel(const uint16_t* __restrict__ qkv, long long ld, uint16_t* __restrict__ out,
                                                            long long ldo, const int* __restrict__ cu, int H, int causal, float scale_log2) {
  // heavier (later) causal tiles first
  const int qt = causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  attn_mma_tile<D, KIND>(qkv, ld, out, ldo, cu, H, causal, scale_log2, blockIdx.z, blockIdx.y, qt);
}

// The same tiles over a device-side LIST of sequences (list[0] = count, list[1..] = sequence ids, written by attn_small_kernel for the
// sequences it leaves alone): gridDim.z CTAs stride over the list, so a batch of 1 500 short sequences and 30 long ones does not launch
// 1 500 x heads x tiles CTAs that exit at once (measured: 76 us of the 180 us call).
template <int D, int KIND>
__global__ void __launch_bounds__(NTHREADS) attn_mma_list_kernel(const uint16_t* __restrict__ qkv, long long ld, uint16_t* __restrict__ out,
                                                                 long long ldo, const int* __restrict__ cu, int H, float scale_log2,
                                                                 const int* __restrict__ list) {
  const int n = list[0];
  for (int i = blockIdx.z; i < n; i += gridDim.z) {
    attn_mma_tile<D, KIND>(qkv, ld, out, ldo, cu, H, 0, scale_log2, list[1 + i], blockIdx.y, blockIdx.x);
    __syncthreads();  // the next tile re-fills the Q / K / V stages
  }
}

// Short sequences (<= 64 tokens: the patch -> instance pooling of a step packs ~1 500 sequences of ~37 tokens): ONE WARP per (sequence, head),
// four independent warps per CTA.  The 128-query-row CTAs of the kernel above spend their time in fixed overhead on such sequences (measured:
// 353 us per layer for 0.3 GFLOP); here a warp stages its Q / K / V head slices (cp.async, zero-filled past the end), runs S = QK^T for 16 query
// rows at a time against the single 64-key tile, a plain (not online) softmax, and O = PV -- the same fragment layouts and 16-bit P as above.
constexpr int SM_WARPS = 4;
template <int KIND>
__global__ void __launch_bounds__(SM_WARPS * 32) attn_small_kernel(const uint16_t* __restrict__ qkv, long long ld, uint16_t* __restrict__ out, long long ldo,
                                                                   const int* __restrict__ cu, int n_pairs, int H, float scale_log2,
                                                                   int* __restrict__ long_list) {
  constexpr int D = 64, LDS = D + 8, KS = D / 16, DT = D / 8, CPR = D / 8, ROWS = 64;
  extern __shared__ __align__(16) uint16_t smem_small[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * SM_WARPS + warp;
  if (pair >= n_pairs) return;
  const int seq = pair / H, h = pair - seq * H;
  const int b = cu[seq], len = cu[seq + 1] - b;
  if (len > ROWS) {  // longer sequences: attn_mma_list_kernel
    if (h == 0 && lane == 0) long_list[1 + atomicAdd(long_list, 1)] = seq;
    return;
  }
  if (len <= 0) return;
  uint16_t* sK = smem_small + (size_t)warp * 2 * ROWS * LDS;
  uint16_t* sV = sK + ROWS * LDS;
  const size_t qoff = (size_t)h * D, koff = (size_t)(H + h) * D, voff = (size_t)(2 * H + h) * D;
  const int rows16 = (len + 15) & ~15;  // key rows with data or cp.async zero fill
  for (int i = lane; i < rows16 * CPR; i += 32) {
    const int r = i / CPR, c = i % CPR;
    const bool ok = r < len;
    const uint16_t* rowp = qkv + (size_t)(b + (ok ? r : 0)) * ld + c * 8;
    const int nb = ok ? 16 : 0;
    cp_async16(smem_u32(&sK[r * LDS + c * 8]), rowp + koff, nb);
    cp_async16(smem_u32(&sV[r * LDS + c * 8]), rowp + voff, nb);
  }
  cp_async_commit();
  // key rows [rows16, 64) of K / V are read by the 64-key fragments below: zero them (P is zero there, but 0 * garbage must stay finite)
  for (int i = lane; i < (ROWS - rows16) * CPR; i += 32) {
    const int r = rows16 + i / CPR, c = i % CPR;
    *reinterpret_cast<uint4*>(&sK[r * LDS + c * 8]) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(&sV[r * LDS + c * 8]) = make_uint4(0, 0, 0, 0);
  }
  const int g = lane >> 2, t = lane & 3;
  // Q as A fragments straight from global memory (each 128-byte head row is consumed whole by the 4 lanes x 4 k-steps that share it)
  auto load_q = [&](int q0, uint32_t (&q)[KS][4]) {
    const int r0 = q0 + g, r1 = r0 + 8;
    const uint16_t* p0 = qkv + (size_t)(b + min(r0, len - 1)) * ld + qoff + 2 * t;
    const uint16_t* p1 = qkv + (size_t)(b + min(r1, len - 1)) * ld + qoff + 2 * t;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      q[ks][0] = r0 < len ? __ldg(reinterpret_cast<const uint32_t*>(p0 + ks * 16)) : 0u;
      q[ks][1] = r1 < len ? __ldg(reinterpret_cast<const uint32_t*>(p1 + ks * 16)) : 0u;
      q[ks][2] = r0 < len ? __ldg(reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8)) : 0u;
      q[ks][3] = r1 < len ? __ldg(reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8)) : 0u;
    }
  };
  uint32_t qn[KS][4];
  load_q(0, qn);
  cp_async_wait<0>();
  __syncwarp();
  for (int q0 = 0; q0 < len; q0 += 16) {
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) { qf[ks][0] = qn[ks][0]; qf[ks][1] = qn[ks][1]; qf[ks][2] = qn[ks][2]; qf[ks][3] = qn[ks][3]; }
    if (q0 + 16 < len) load_q(q0 + 16, qn);  // next tile's Q in flight under this tile's math
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int dcol = ks * 16 + (((lane >> 3) & 1) << 3);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(&sK[key * LDS + dcol]), b0, b1, b2, b3);
        mma16816<KIND>(s[2 * np], qf[ks], b0, b1);
        mma16816<KIND>(s[2 * np + 1], qf[ks], b2, b3);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + 2 * t + (e & 1);
        const float v = key < len ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - mx0), p1 = exp2f(s[nt][1] - mx0);
      const float p2 = exp2f(s[nt][2] - mx1), p3 = exp2f(s[nt][3] - mx1);
      rs0 += p0 + p1; rs1 += p2 + p3;
      const int kk = nt >> 1, hi = nt & 1;
      pf[kk][hi * 2 + 0] = pack16x2(p0, p1, KIND);
      pf[kk][hi * 2 + 1] = pack16x2(p2, p3, KIND);
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {
        const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int dcol = dp * 16 + ((lane >> 4) << 3);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(smem_u32(&sV[key * LDS + dcol]), b0, b1, b2, b3);
        mma16816<KIND>(o[2 * dp], pf[kk], b0, b1);
        mma16816<KIND>(o[2 * dp + 1], pf[kk], b2, b3);
      }
    }
    const float inv0 = rs0 > 0.f ? 1.0f / rs0 : 0.f, inv1 = rs1 > 0.f ? 1.0f / rs1 : 0.f;
    const int row0 = q0 + g, row1 = row0 + 8;
#pragma unroll
    for (int dt = 0; dt < DT; ++dt) {
      const int col = dt * 8 + 2 * t;
      if (row0 < len) *reinterpret_cast<uint32_t*>(out + (size_t)(b + row0) * ldo + (size_t)h * D + col) = pack16x2(o[dt][0] * inv0, o[dt][1] * inv0, KIND);
      if (row1 < len) *reinterpret_cast<uint32_t*>(out + (size_t)(b + row1) * ldo + (size_t)h * D + col) = pack16x2(o[dt][2] * inv1, o[dt][3] * inv1, KIND);
    }
  }
}

template <int D>
int launch(const void* qkv, long long ld, void* out, long long ldo, const int* cu, int n_seq, int max_len, int H, int causal, int kind,
           float scale, cudaStream_t st) {
  dim3 grid(d3d_cdiv(max_len, BQ), H, n_seq);
  const float sl2 = scale * 1.4426950408889634f;
  constexpr int SMEM = (BQ + 4 * BKV) * (D + 8) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_kernel<D, D3D_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_kernel<D, D3D_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  if (kind == D3D_BF16)
    attn_mma_kernel<D, D3D_BF16><<<grid, NTHREADS, SMEM, st>>>((const uint16_t*)qkv, ld, (uint16_t*)out, ldo, cu, H, causal, sl2);
  else
    attn_mma_kernel<D, D3D_F16><<<grid, NTHREADS, SMEM, st>>>((const uint16_t*)qkv, ld, (uint16_t*)out, ldo, cu, H, causal, sl2);
  D3D_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int d3d_attention_mma(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                                 int Dh, int causal, int kind, float scale, void* stream) {
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(qkv && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == 64 || Dh == 96, "head_dim 64 or 96");
  D3D_REQUIRE(ld % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)qkv % 16) == 0, "16-byte aligned rows");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
  cudaStream_t st = (cudaStream_t)stream;
  if (Dh == 64) return launch<64>(qkv, ld, out, ldo, cu_seqlens, n_seq, max_len, H, causal, kind, scale, st);
  return launch<96>(qkv, ld, out, ldo, cu_seqlens, n_seq, max_len, H, causal, kind, scale, st);
}

// Packed batches that mix many short sequences with a few long ones (non-causal, head_dim 64): sequences of <= 64 tokens on the
// warp-per-(sequence, head) kernel, the others on the 128-row kernel (which skips the short ones).  Same contract as d3d_attention_mma.
extern "C" int d3d_attention_mixed(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                                   int Dh, int kind, float scale, void* stream) {
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(qkv && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == 64, "head_dim 64");
  D3D_REQUIRE(ld % 8 == 0 && ldo % 2 == 0 && ((uintptr_t)qkv % 16) == 0, "16-byte aligned rows");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int SMEM = SM_WARPS * 2 * 64 * (64 + 8) * 2;
  constexpr int SMEM_L = (BQ + 4 * BKV) * (64 + 8) * (int)sizeof(uint16_t);
  static bool attr_set = false;
  if (!attr_set) {
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<D3D_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<D3D_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_list_kernel<64, D3D_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_L));
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_list_kernel<64, D3D_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_L));
    attr_set = true;
  }
  // the list of long sequences lives in a per-stream device buffer owned by the library (two streams may run this entry concurrently)
  int* list = nullptr;
  {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, std::pair<int*, int>> lists;
    int dev = 0;
    D3D_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto& e = lists[{dev, st}];
    if (e.second < n_seq + 1) {
      if (e.first) D3D_CHECK_CUDA(cudaFree(e.first));  // synchronises: nothing in flight reads the old list
      e.second = std::max(n_seq + 1, 4096);
      D3D_CHECK_CUDA(cudaMalloc(&e.first, (size_t)e.second * sizeof(int)));
    }
    list = e.first;
  }
  D3D_CHECK_CUDA(cudaMemsetAsync(list, 0, sizeof(int), st));
  const int n_pairs = n_seq * H;
  const float sl2 = scale * 1.4426950408889634f;
  if (kind == D3D_BF16)
    attn_small_kernel<D3D_BF16><<<d3d_cdiv(n_pairs, SM_WARPS), SM_WARPS * 32, SMEM, st>>>((const uint16_t*)qkv, ld, (uint16_t*)out, ldo, cu_seqlens, n_pairs, H, sl2, list);
  else
    attn_small_kernel<D3D_F16><<<d3d_cdiv(n_pairs, SM_WARPS), SM_WARPS * 32, SMEM, st>>>((const uint16_t*)qkv, ld, (uint16_t*)out, ldo, cu_seqlens, n_pairs, H, sl2, list);
  D3D_CHECK_LAUNCH();
  if (max_len > 64) {
    // CTAs per (q tile, head) column striding over the list: enough to fill the 148 SMs twice at 2 CTAs / SM
    const int nq = d3d_cdiv(max_len, BQ);
    const int nz = std::max(1, std::min(n_seq, d3d_cdiv(4 * 148, nq * H)));
    dim3 grid(nq, H, nz);
    if (kind == D3D_BF16)
      attn_mma_list_kernel<64, D3D_BF16><<<grid, NTHREADS, SMEM_L, st>>>((const uint16_t*)qkv, ld, (uint16_t*)out, ldo, cu_seqlens, H, sl2, list);
    else
      attn_mma_list_kernel<64, D3D_F16><<<grid, NTHREADS, SMEM_L, st>>>((const uint16_t*)qkv, ld, (uint16_t*)out, ldo, cu_seqlens, H, sl2, list);
    D3D_CHECK_LAUNCH();
  }
  return 0;
}
