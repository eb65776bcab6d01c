// Precise-mode attention on the tensor cores: fp32 Q/K/V in, fp32 out, every product computed with SPLIT fp16x2 operands
//   x = hi + lo,  hi = fp16(x), lo = fp16(x - hi)   (~22 significant bits, the same split d3d_split16 gives the precise GEMMs)
//   S = Qh Kh^T + Ql Kh^T + Qh Kl^T,   O += Ph Vh + Pl Vh + Ph Vl      (mma.sync m16n8k16, fp32 accumulate; the lo x lo terms are ~2^-22)
// with the online softmax in fp32.  Same packed variable-length contract as d3d_attention_f32 (which it replaces in precise.py: the
// CUDA-core kernel was ~70 % of a precise step); parity target: <= 1e-3 on the action logits vs the pure-fp32 oracle at full depth.
// One CTA = 128 queries of one (sequence, head), 8 warps x 16 rows, 64-key tiles.  The operands arrive PRE-SPLIT: d3d_split16 turns the fp32
// QKV matrix of a layer into [hi | lo] fp16 halves once (instead of once per query tile), and the K / V hi / lo tiles stream through a
// double-buffered cp.async pipeline like the production mma kernel's.
#include "common.cuh"

namespace {

constexpr int NWARPS = 8;
constexpr int BQ = 16 * NWARPS;
constexpr int BKV = 64;
constexpr int NTHREADS = 32 * NWARPS;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half ah = __float2half_rn(a), bh = __float2half_rn(b);
  const __half al = __float2half_rn(a - __half2float(ah)), bl = __float2half_rn(b - __half2float(bh));
  __half2 h = __halves2half2(ah, bh), l = __halves2half2(al, bl);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// qkv: [T, ld] fp16 rows = [q | k | v hi halves (3*H*D) ... | lo halves at column lo_off ...]
template <int D>
__global__ void __launch_bounds__(NTHREADS) attn_split_kernel(const uint16_t* __restrict__ qkv, long long ld, long long lo_off, float* __restrict__ out,
                                                              long long ldo, const int* __restrict__ cu, int H, int causal, float scale_log2) {
  constexpr int LDS = D + 8, KS = D / 16, DT = D / 8, CPR = D / 8;
  constexpr int STAGE = 4 * BKV * LDS;  // halves per stage: K hi, K lo, V hi, V lo
  extern __shared__ __align__(16) uint16_t smem_split[];
  uint16_t* sQh = smem_split;
  uint16_t* sQl = sQh + BQ * LDS;
  uint16_t* sKV = sQl + BQ * LDS;  // 2 stages

  const int seq = blockIdx.z, h = blockIdx.y;
  const int b = cu[seq], len = cu[seq + 1] - b;
  const int qt = causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int q0 = qt * BQ;
  if (q0 >= len) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t qoff = (size_t)h * D, koff = (size_t)(H + h) * D, voff = (size_t)(2 * H + h) * D;

  for (int i = threadIdx.x; i < BQ * CPR; i += NTHREADS) {
    const int r = i / CPR, c = i % CPR;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = vh;
    if (q0 + r < len) {
      const uint16_t* rowp = qkv + (size_t)(b + q0 + r) * ld + qoff + c * 8;
      vh = *reinterpret_cast<const uint4*>(rowp);
      vl = *reinterpret_cast<const uint4*>(rowp + lo_off);
    }
    *reinterpret_cast<uint4*>(&sQh[r * LDS + c * 8]) = vh;
    *reinterpret_cast<uint4*>(&sQl[r * LDS + c * 8]) = vl;
  }
  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int row0 = q0 + warp * 16 + g, row1 = row0 + 8;
  const int kmax = causal ? min(len, q0 + BQ) : len;
  const int n_tiles = (kmax + BKV - 1) / BKV;
  const int qrow = warp * 16 + (lane & 15), qcolsel = (lane >> 4) * 8;

  auto load_tile = [&](int tile, int stage) {
    const int c0 = tile * BKV;
    uint16_t* dKh = sKV + stage * STAGE;
    uint16_t* dKl = dKh + BKV * LDS;
    uint16_t* dVh = dKl + BKV * LDS;
    uint16_t* dVl = dVh + BKV * LDS;
    for (int i = threadIdx.x; i < BKV * CPR; i += NTHREADS) {
      const int r = i / CPR, c = i % CPR;
      const bool ok = c0 + r < len;
      const uint16_t* rowp = qkv + (size_t)(b + (ok ? c0 + r : 0)) * ld + c * 8;
      const int nb = ok ? 16 : 0;
      cp_async16(smem_u32(&dKh[r * LDS + c * 8]), rowp + koff, nb);
      cp_async16(smem_u32(&dKl[r * LDS + c * 8]), rowp + koff + lo_off, nb);
      cp_async16(smem_u32(&dVh[r * LDS + c * 8]), rowp + voff, nb);
      cp_async16(smem_u32(&dVl[r * LDS + c * 8]), rowp + voff + lo_off, nb);
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
    __syncthreads();  // tile's K / V (and, first time, Q) are visible to every warp
    const uint16_t* sKh = sKV + (tile & 1) * STAGE;
    const uint16_t* sKl = sKh + BKV * LDS;
    const uint16_t* sVh = sKl + BKV * LDS;
    const uint16_t* sVl = sVh + BKV * LDS;
    const bool active = !(causal && c0 > q0 + warp * 16 + 15);  // whole tile above the diagonal for this warp (warp-uniform)
    if (active) {
    // ---- S = Qh Kh^T + Ql Kh^T + Qh Kl^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t qh[4], ql[4];
      ldsm_x4(smem_u32(&sQh[qrow * LDS + ks * 16 + qcolsel]), qh[0], qh[1], qh[2], qh[3]);
      ldsm_x4(smem_u32(&sQl[qrow * LDS + ks * 16 + qcolsel]), ql[0], ql[1], ql[2], ql[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int dcol = ks * 16 + (((lane >> 3) & 1) << 3);
        uint32_t b0, b1, b2, b3, c0_, c1_, c2_, c3_;
        ldsm_x4(smem_u32(&sKh[key * LDS + dcol]), b0, b1, b2, b3);
        ldsm_x4(smem_u32(&sKl[key * LDS + dcol]), c0_, c1_, c2_, c3_);
        mma16816(s[2 * np], ql, b0, b1);       // small terms first, the leading term last
        mma16816(s[2 * np], qh, c0_, c1_);
        mma16816(s[2 * np], qh, b0, b1);
        mma16816(s[2 * np + 1], ql, b2, b3);
        mma16816(s[2 * np + 1], qh, c2_, c3_);
        mma16816(s[2 * np + 1], qh, b2, b3);
      }
    }
    // ---- mask + online softmax in fp32 (base-2 exponent) ----
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
    const float corr0 = exp2f(m0 - base0), corr1 = exp2f(m1 - base1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t ph[4][4], pl[4][4];  // P (hi / lo) as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - base0), p1 = exp2f(s[nt][1] - base0);
      const float p2 = exp2f(s[nt][2] - base1), p3 = exp2f(s[nt][3] - base1);
      rs0 += p0 + p1; rs1 += p2 + p3;
      const int kk = nt >> 1, hi = nt & 1;
      split2(p0, p1, ph[kk][hi * 2 + 0], pl[kk][hi * 2 + 0]);
      split2(p2, p3, ph[kk][hi * 2 + 1], pl[kk][hi * 2 + 1]);
    }
    rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
    rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
    l0 = l0 * corr0 + rs0; l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
    // ---- O += Ph Vh + Pl Vh + Ph Vl ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {
        const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int dcol = dp * 16 + ((lane >> 4) << 3);
        uint32_t b0, b1, b2, b3, c0_, c1_, c2_, c3_;
        ldsm_x4_trans(smem_u32(&sVh[key * LDS + dcol]), b0, b1, b2, b3);
        ldsm_x4_trans(smem_u32(&sVl[key * LDS + dcol]), c0_, c1_, c2_, c3_);
        mma16816(o[2 * dp], pl[kk], b0, b1);
        mma16816(o[2 * dp], ph[kk], c0_, c1_);
        mma16816(o[2 * dp], ph[kk], b0, b1);
        mma16816(o[2 * dp + 1], pl[kk], b2, b3);
        mma16816(o[2 * dp + 1], ph[kk], c2_, c3_);
        mma16816(o[2 * dp + 1], ph[kk], b2, b3);
      }
    }
    }  // active
    __syncthreads();  // everyone is done with this stage before tile + 2 overwrites it
  }
  const float inv0 = l0 > 0.f ? 1.0f / l0 : 0.f, inv1 = l1 > 0.f ? 1.0f / l1 : 0.f;
#pragma unroll
  for (int dt = 0; dt < DT; ++dt) {
    const int col = dt * 8 + 2 * t;
    if (row0 < len) *reinterpret_cast<float2*>(out + (size_t)(b + row0) * ldo + (size_t)h * D + col) = make_float2(o[dt][0] * inv0, o[dt][1] * inv0);
    if (row1 < len) *reinterpret_cast<float2*>(out + (size_t)(b + row1) * ldo + (size_t)h * D + col) = make_float2(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

template <int D>
int launch(const uint16_t* qkv, long long ld, long long lo_off, float* out, long long ldo, const int* cu, int n_seq, int max_len, int H, int causal,
           float scale, cudaStream_t st) {
  dim3 grid(d3d_cdiv(max_len, BQ), H, n_seq);
  constexpr int SMEM = (2 * BQ + 8 * BKV) * (D + 8) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_split_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  attn_split_kernel<D><<<grid, NTHREADS, SMEM, st>>>(qkv, ld, lo_off, out, ldo, cu, H, causal, scale * 1.4426950408889634f);
  D3D_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int d3d_attention_split(const void* qkv_hl, int64_t ld, int64_t lo_off, float* out, int64_t ldo, const int* cu_seqlens, int n_seq,
                                   int max_len, int H, int Dh, int causal, float scale, void* stream) {
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(qkv_hl && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == 64 || Dh == 96, "head_dim 64 or 96");
  D3D_REQUIRE(ld % 8 == 0 && lo_off % 8 == 0 && lo_off >= 3LL * H * Dh && ldo % 2 == 0 && ((uintptr_t)qkv_hl % 16) == 0 && ((uintptr_t)out % 8) == 0,
              "16-byte aligned fp16 rows [hi | lo], 8-byte aligned output rows");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
  cudaStream_t st = (cudaStream_t)stream;
  if (Dh == 64) return launch<64>((const uint16_t*)qkv_hl, ld, lo_off, out, ldo, cu_seqlens, n_seq, max_len, H, causal, scale, st);
  return launch<96>((const uint16_t*)qkv_hl, ld, lo_off, out, ldo, cu_seqlens, n_seq, max_len, H, causal, scale, st);
}
