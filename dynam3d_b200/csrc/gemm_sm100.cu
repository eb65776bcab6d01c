// tcgen05 / TMEM / TMA persistent GEMM for sm_100a:  C[M,N] = epi(A[M,K] @ W[N,K]^T), fp32 accumulate.
//
// Structure (one CTA per SM, 384 threads, static persistent tile schedule):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor.2d of a 128x64 A tile and a BNx64 W tile per stage
//                                  (SWIZZLE_128B, K-major), mbarrier complete_tx
//   warp 1      : MMA issuer    -- one elected lane issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16)
//                                  per stage into a TMEM accumulator; tcgen05.commit frees the smem stage and, after the
//                                  last K block, publishes the accumulator
//   warp 2      : TMEM allocator (2*BN columns: double-buffered accumulator so the epilogue of tile i overlaps
//                                  the main loop of tile i+1)
//   warps 4..11 : epilogue      -- tcgen05.ld 32x32b (one accumulator row per thread, two warps per TMEM lane quadrant split the
//                                  columns), bias / activation / residual in registers, vectorised global stores
// CTAS == 2 (large problems): the two CTAs of a cluster (one TPC) work as a pair on a 256 x BN tile with
//   tcgen05.mma.cta_group::2 (UMMA_M = 256): each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows), so the
//   shared-memory traffic per MMA halves; TMA loads of both CTAs complete on the leader's full barrier, the leader's MMA warp
//   issues for the pair and multicasts its commits to both CTAs' empty / tmem-full barriers; every CTA drains its own 128 TMEM
//   lanes, and the peer's epilogue warps release the accumulator with a remote mbarrier arrive on the leader.
// Replaces the cuBLAS calls behind every nn.Linear on the reference path (see include/dynam3d_b200.h).
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = 128 B = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;  // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue (2 per TMEM lane quadrant)
constexpr int NUM_EPI_WARPS = 8;

template <int BN, int CTAS>
struct Cfg {
  static constexpr int STAGES = (BN == 256 && CTAS == 1) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = BN / CTAS;  // rows of the W tile this CTA stages
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512: power of two >= 32
  static constexpr int EPI_STAGING = NUM_EPI_WARPS * 32 * 32 * 4;  // one 32x32 fp32 transposition tile per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING + 1024 /*align*/ + 256 /*barriers*/;
};

struct Epilogue {
  void* C;
  long long ldc;
  const float* bias;
  const float* residual;
  long long ldres;
  int act;
  int out_kind;
};

// ---------------- PTX wrappers ----------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- cta_group::2 (CTA pair) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes complete on `bar_cluster` (the LEADER's full barrier)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs: arrives once on the barrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait that also names the destination registers of the outstanding tcgen05.ld, so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// x * sigmoid(a x) = 0.5 x (1 + tanh(0.5 a x)): ONE MUFU op (tanh.approx, rel. error 2^-11) instead of ex2 + rcp.  Used where the
// result is stored as a 16-bit value anyway (same 2^-11 rounding); fp32 outputs (precise mode) take the exact path.
__device__ __forceinline__ float x_sigmoid_fast(float x, float half_a) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * half_a));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor):
// [0,14) start>>4 | [16,30) LBO>>4 (=1, ignored for swizzled K-major) | [32,46) SBO>>4 (=64) | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b format (0=f16,1=bf16) at 7/10,
// K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc(int kind, int n, int m = BM) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)kind << 7;
  d |= (uint32_t)kind << 10;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case D3D_ACT_QUICK_GELU: return quick_gelu(v);
    case D3D_ACT_GELU: return gelu_erf(v);
    case D3D_ACT_SILU: return silu(v);
    case D3D_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;
    default: return v;
  }
}

// Epilogue of one 32-row x (BN/2)-column slab of an accumulator tile, executed by ONE warp (TMEM lane quadrant = rows row0..row0+31).
// Per 32-column chunk: tcgen05.ld (one accumulator row per thread) -> the raw 32x32 fp32 chunk is transposed through the warp's
// XOR-swizzled smem tile (the next chunk's TMEM load is issued right behind it) -> every lane then owns 8 consecutive columns of one row
// per pass (4 passes): + bias (8 registers, loaded before the TMEM wait so the L2 latency is hidden), activation, + fp32 residual and
// a row-contiguous vector store (32 B .. 128 B per row and instruction instead of 32 rows x 16 B).  All addressing that does not
// depend on the chunk is hoisted out of the chunk loop.
template <int BN, bool SWIGLU>
__device__ __forceinline__ void epilogue_tile(const Epilogue& ep, uint32_t taddr, uint32_t stg, int lane, int chalf, int row0, int n0, int M, int N) {
  constexpr int PASSES = 4;
  constexpr int OW = SWIGLU ? 4 : 8;  // output columns per lane and pass (SWIGLU: acc cols (2j, 2j+1) -> out col j)
  const int n_out = SWIGLU ? (N >> 1) : N;
  const int rr0 = lane >> 2;        // row of this lane in pass 0 (pass p: + 8p)
  const int g0 = (lane & 3) * 2;    // first of the lane's two 16-byte groups = accumulator columns g0*4 .. g0*4+7 of the chunk
  const int esz = ep.out_kind == D3D_OUT_F32 ? 4 : 2;
  const bool fast = ep.out_kind != D3D_OUT_F32;
  const uint32_t st_row = stg + (uint32_t)lane * 128u;
  const uint32_t sw = (uint32_t)(lane & 7);
  uint32_t ld_a[PASSES], ld_b[PASSES];
  char* cptr[PASSES];
  const float* rptr[PASSES];
  bool ok[PASSES];
  const int oc_lane = SWIGLU ? ((n0 + g0 * 4) >> 1) : (n0 + g0 * 4);  // output column of the lane's first element in chunk 0 of the tile
#pragma unroll
  for (int p = 0; p < PASSES; ++p) {
    const int rr = rr0 + p * 8;
    const uint32_t rsw = (uint32_t)(rr & 7);
    ld_a[p] = stg + (uint32_t)rr * 128u + (((uint32_t)g0 ^ rsw) << 4);
    ld_b[p] = stg + (uint32_t)rr * 128u + (((uint32_t)(g0 + 1) ^ rsw) << 4);
    const long long grow = row0 + rr;
    ok[p] = grow < M;
    cptr[p] = (char*)ep.C + (grow * ep.ldc + oc_lane) * esz;
    rptr[p] = ep.residual ? ep.residual + grow * ep.ldres + oc_lane : nullptr;
  }
  const int c_begin = chalf * (BN / 2), c_end = (chalf + 1) * (BN / 2);
  uint32_t r[32];
  tmem_ld32(taddr + (uint32_t)c_begin, r);
#pragma unroll 1
  for (int c = c_begin; c < c_end; c += 32) {
    const int col0 = n0 + c;                 // first accumulator column of the chunk
    const int acol = col0 + g0 * 4;          // first accumulator column of this lane
    const bool vec = col0 + 32 <= N;         // warp-uniform: the whole chunk is inside the matrix
    float b8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b8[i] = 0.f;
    if (ep.bias && col0 < N) {
      if (vec) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(ep.bias + acol)), t1 = __ldg(reinterpret_cast<const float4*>(ep.bias + acol) + 1);
        b8[0] = t0.x; b8[1] = t0.y; b8[2] = t0.z; b8[3] = t0.w; b8[4] = t1.x; b8[5] = t1.y; b8[6] = t1.z; b8[7] = t1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (acol + i < N) b8[i] = __ldg(ep.bias + acol + i);
      }
    }
    tmem_ld_wait32(r);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_row + (((uint32_t)j ^ sw) << 4)), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                   "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                   : "memory");
    if (c + 32 < c_end) tmem_ld32(taddr + (uint32_t)(c + 32), r);  // lands while this chunk is processed
    __syncwarp();
    if (col0 < N) {  // warp-uniform
      const int oc_chunk = SWIGLU ? (c >> 1) : c;  // output-column offset of this chunk inside the tile
      // straight-line over the lane's 4 x 8 values: all smem loads first, then the arithmetic with 32-way ILP, then the memory operations
      float f[PASSES][8];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[p][0]), "=f"(f[p][1]), "=f"(f[p][2]), "=f"(f[p][3]) : "r"(ld_a[p]));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[p][4]), "=f"(f[p][5]), "=f"(f[p][6]), "=f"(f[p][7]) : "r"(ld_b[p]));
      }
      if (ep.bias) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
#pragma unroll
          for (int i = 0; i < 8; ++i) f[p][i] += b8[i];
      }
      if (SWIGLU) {
        if (fast) {
#pragma unroll
          for (int p = 0; p < PASSES; ++p)
#pragma unroll
            for (int i = 0; i < 4; ++i) f[p][i] = x_sigmoid_fast(f[p][2 * i], 0.5f) * f[p][2 * i + 1];
        } else {
#pragma unroll
          for (int p = 0; p < PASSES; ++p)
#pragma unroll
            for (int i = 0; i < 4; ++i) f[p][i] = __fdividef(f[p][2 * i], 1.0f + __expf(-f[p][2 * i])) * f[p][2 * i + 1];
        }
      } else if (ep.act == D3D_ACT_QUICK_GELU) {
        if (fast) {
#pragma unroll
          for (int p = 0; p < PASSES; ++p)
#pragma unroll
            for (int i = 0; i < 8; ++i) f[p][i] = x_sigmoid_fast(f[p][i], 0.851f);
        } else {
#pragma unroll
          for (int p = 0; p < PASSES; ++p)
#pragma unroll
            for (int i = 0; i < 8; ++i) f[p][i] = __fdividef(f[p][i], 1.0f + __expf(-1.702f * f[p][i]));
        }
      } else if (ep.act == D3D_ACT_GELU) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
#pragma unroll
          for (int i = 0; i < 8; ++i) f[p][i] = gelu_erf(f[p][i]);
      } else if (ep.act == D3D_ACT_SILU) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
#pragma unroll
          for (int i = 0; i < 8; ++i) f[p][i] = __fdividef(f[p][i], 1.0f + __expf(-f[p][i]));
      } else if (ep.act == D3D_ACT_LEAKY_RELU) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p)
#pragma unroll
          for (int i = 0; i < 8; ++i) f[p][i] = f[p][i] > 0.f ? f[p][i] : 0.01f * f[p][i];
      }
      if (vec) {
        if (ep.residual) {
          float4 t[PASSES][2];
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
            t[p][0] = t[p][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[p]) {
              const float4* res = reinterpret_cast<const float4*>(rptr[p] + oc_chunk);
              t[p][0] = res[0];
              if (!SWIGLU) t[p][1] = res[1];
            }
          }
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
            f[p][0] += t[p][0].x; f[p][1] += t[p][0].y; f[p][2] += t[p][0].z; f[p][3] += t[p][0].w;
            f[p][4] += t[p][1].x; f[p][5] += t[p][1].y; f[p][6] += t[p][1].z; f[p][7] += t[p][1].w;
          }
        }
        if (ep.out_kind == D3D_OUT_F32) {
#pragma unroll
          for (int p = 0; p < PASSES; ++p)
            if (ok[p]) {
              float4* dst = reinterpret_cast<float4*>(cptr[p] + oc_chunk * 4);
              dst[0] = make_float4(f[p][0], f[p][1], f[p][2], f[p][3]);
              if (!SWIGLU) dst[1] = make_float4(f[p][4], f[p][5], f[p][6], f[p][7]);
            }
        } else if (ep.out_kind == D3D_BF16) {
#pragma unroll
          for (int p = 0; p < PASSES; ++p)
            if (ok[p]) {
              if (SWIGLU) *reinterpret_cast<uint2*>(cptr[p] + oc_chunk * 2) = make_uint2(pack16x2(f[p][0], f[p][1], D3D_BF16), pack16x2(f[p][2], f[p][3], D3D_BF16));
              else *reinterpret_cast<uint4*>(cptr[p] + oc_chunk * 2) = make_uint4(pack16x2(f[p][0], f[p][1], D3D_BF16), pack16x2(f[p][2], f[p][3], D3D_BF16),
                                                                                 pack16x2(f[p][4], f[p][5], D3D_BF16), pack16x2(f[p][6], f[p][7], D3D_BF16));
            }
        } else {
#pragma unroll
          for (int p = 0; p < PASSES; ++p)
            if (ok[p]) {
              if (SWIGLU) *reinterpret_cast<uint2*>(cptr[p] + oc_chunk * 2) = make_uint2(pack16x2(f[p][0], f[p][1], D3D_F16), pack16x2(f[p][2], f[p][3], D3D_F16));
              else *reinterpret_cast<uint4*>(cptr[p] + oc_chunk * 2) = make_uint4(pack16x2(f[p][0], f[p][1], D3D_F16), pack16x2(f[p][2], f[p][3], D3D_F16),
                                                                                 pack16x2(f[p][4], f[p][5], D3D_F16), pack16x2(f[p][6], f[p][7], D3D_F16));
            }
        }
      } else {  // ragged last columns: scalar
        const int ocol = oc_lane + oc_chunk;
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
#pragma unroll
          for (int i = 0; i < OW; ++i) {
            if (ok[p] && ocol + i < n_out) {
              float x = f[p][i];
              if (rptr[p]) x += rptr[p][oc_chunk + i];
              if (ep.out_kind == D3D_OUT_F32) reinterpret_cast<float*>(cptr[p])[oc_chunk + i] = x;
              else st16(cptr[p], (size_t)(oc_chunk + i), x, ep.out_kind);
            }
          }
        }
      }
    }
    __syncwarp();  // the smem tile is rewritten by the next chunk
  }
}

// Tile order: column groups of RASTER_GW n-tiles are swept over all m-tiles before the next group starts, so the CTAs running at the
// same time touch ~GW W-tiles and ~#CTAs/GW A-tiles (instead of one A tile and every W tile): the W slice of a group stays in L2
// while it is reused by every m-tile (Phi-3 gate/up: 100 MB of weights no longer stream through L2 once per m-tile).
constexpr int RASTER_GW = 8;
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int group_tiles = RASTER_GW * tiles_m;
  const int grp = tile / group_tiles, rem = tile - grp * group_tiles;
  const int left = tiles_n - grp * RASTER_GW;
  const int gw = left < RASTER_GW ? left : RASTER_GW;
  tm = rem / gw;
  tn = grp * RASTER_GW + (rem - tm * gw);
}

template <int BN, int CTAS>
__device__ __forceinline__ void gemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const Epilogue& ep, int M, int N, int K, int in_kind) {
  using C = Cfg<BN, CTAS>;
  constexpr int TM = BM * CTAS;  // rows of the tile a CTA (pair) owns
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = epi_base + C::EPI_STAGING;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr (4 B)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (M + TM - 1) / TM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int first_tile = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS * CTAS);  // one arrive per epilogue warp (of both CTAs of a pair, on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync(); else __syncthreads();  // the peer's barriers must be initialised before any remote arrive / complete_tx
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        int tm, tn;
        tile_coords(tile, tiles_m, tiles_n, tm, tn);
        const int m0 = tm * TM + (int)cta_rank * BM;
        const int n0 = tn * BN + (int)cta_rank * C::B_ROWS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          if (CTAS == 2) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the pair
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            const uint32_t lead_full = mapa(full_bar(stage), 0);
            tma_load_2d_2sm(sa, &tmA, lead_full, kb * BK, m0);
            tma_load_2d_2sm(sa + C::A_BYTES, &tmB, lead_full, kb * BK, n0);
          } else {
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
            tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m0);
            tma_load_2d(sa + C::A_BYTES, &tmB, full_bar(stage), kb * BK, n0);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {  // in a pair only the leader issues
      const uint32_t idesc = make_idesc(in_kind, BN, TM);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint64_t da = make_smem_desc(sa);
          const uint64_t db = make_smem_desc(sa + C::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle atom: +2 in the (addr >> 4) field
            if (CTAS == 2) umma_f16_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            else umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          if (CTAS == 2) {
            umma_commit_2sm(empty_bar(stage));  // frees this smem stage in BOTH CTAs when the MMAs above retire
            if (kb == num_kb - 1) umma_commit_2sm(tfull_bar(acc));
          } else {
            umma_commit(empty_bar(stage));  // frees this smem stage when the MMAs above retire
            if (kb == num_kb - 1) umma_commit(tfull_bar(acc));
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warps 4..11; warp w drains TMEM lane quadrant w%4, column half (w-4)/4 =====================
    const int q = warp & 3;
    const int chalf = (warp - 4) >> 2;
    const uint32_t stg = epi_base + (uint32_t)(warp - 4) * 4096u;
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      int tm, tn;
      tile_coords(tile, tiles_m, tiles_n, tm, tn);
      const int m0 = tm * TM + (int)cta_rank * BM;
      const int n0 = tn * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if (ep.act == D3D_ACT_SWIGLU) epilogue_tile<BN, true>(ep, taddr, stg, lane, chalf, m0 + q * 32, n0, M, N);
      else epilogue_tile<BN, false>(ep, taddr, stg, lane, chalf, m0 + q * 32, n0, M, N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(mapa(tempty_bar(acc), 0));  // the leader's MMA warp waits for both CTAs' epilogues
        else mbar_arrive(tempty_bar(acc));
      }
    }
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync(); else __syncthreads();  // the leader's MMAs read the peer's smem: nobody leaves early
  if (warp == 2) {
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Epilogue ep, int M, int N,
                    int K, int in_kind) {
  gemm_body<BN, 1>(tmA, tmB, ep, M, N, K, in_kind);
}

// CTA-pair variant: cluster of 2 (same TPC), 256 x BN tile per pair
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Epilogue ep, int M, int N,
                         int K, int in_kind) {
  gemm_body<BN, 2>(tmA, tmB, ep, M, N, K, in_kind);
}

// ---------------- host side ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap(CUtensorMap* map, const void* base, int kind, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    d3d_set_error("cuTensorMapEncodeTiled entry point not available");
    return D3D_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, kind == D3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    d3d_set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%lld cols=%lld ld=%lld", (int)r, base, rows, cols, ld);
    return D3D_ECUDA;
  }
  return 0;
}

int g_gemm_sm_limit = 0;  // 0 = all SMs; otherwise persistent grids are capped (d3d_gemm_set_sm_limit)
inline int usable_sms() { return (g_gemm_sm_limit > 0 && g_gemm_sm_limit < d3d_num_sms()) ? g_gemm_sm_limit : d3d_num_sms(); }

template <int BN>
int launch(const d3d_gemm_args& a, cudaStream_t st) {
  using C = Cfg<BN, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    D3D_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  D3D_TRY(make_tmap(&tmA, a.A, a.in_kind, a.M, a.K, a.lda, BM));
  D3D_TRY(make_tmap(&tmB, a.W, a.in_kind, a.N, a.K, a.ldw, BN));
  Epilogue ep{a.C, a.ldc, a.bias, a.residual, a.ldres, a.act, a.out_kind};
  const int tiles = d3d_cdiv(a.M, BM) * d3d_cdiv(a.N, BN);
  const int grid = tiles < usable_sms() ? tiles : usable_sms();
  gemm_tcgen05_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(tmA, tmB, ep, a.M, a.N, a.K, a.in_kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

__global__ void gemm_simt_kernel(d3d_gemm_args a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool swiglu = a.act == D3D_ACT_SWIGLU;
  const int nout = swiglu ? a.N / 2 : a.N;
  if (idx >= (long long)a.M * nout) return;
  const int m = (int)(idx / nout), j = (int)(idx % nout);
  auto dot = [&](int n) {
    float acc = 0.f;
    for (int k = 0; k < a.K; ++k) acc += ld16(a.A, (size_t)m * a.lda + k, a.in_kind) * ld16(a.W, (size_t)n * a.ldw + k, a.in_kind);
    if (a.bias) acc += a.bias[n];
    return acc;
  };
  float v;
  if (swiglu) {
    v = silu(dot(2 * j)) * dot(2 * j + 1);
  } else {
    v = apply_act(dot(j), a.act);
    if (a.residual) v += a.residual[(size_t)m * a.ldres + j];
  }
  if (a.out_kind == D3D_OUT_F32) ((float*)a.C)[(size_t)m * a.ldc + j] = v;
  else st16(a.C, (size_t)m * a.ldc + j, v, a.out_kind);
}

int validate(const d3d_gemm_args& a) {
  D3D_REQUIRE(a.A && a.W && a.C, "null operand");
  D3D_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "empty problem");
  D3D_REQUIRE(a.in_kind == D3D_F16 || a.in_kind == D3D_BF16, "in_kind");
  D3D_REQUIRE(a.out_kind == D3D_F16 || a.out_kind == D3D_BF16 || a.out_kind == D3D_OUT_F32, "out_kind");
  D3D_REQUIRE(a.lda >= a.K && a.ldw >= a.K, "leading dimension < K");
  D3D_REQUIRE((a.lda % 8) == 0 && (a.ldw % 8) == 0, "lda/ldw must be multiples of 8 elements (16 B TMA stride)");
  D3D_REQUIRE(((uintptr_t)a.A % 16) == 0 && ((uintptr_t)a.W % 16) == 0, "A/W must be 16-byte aligned");
  D3D_REQUIRE(a.act != D3D_ACT_SWIGLU || ((a.N % 2) == 0 && a.residual == nullptr), "swiglu needs even N, no residual");
  const int esz = a.out_kind == D3D_OUT_F32 ? 4 : 2;
  D3D_REQUIRE(((uintptr_t)a.C % 16) == 0 && ((a.ldc * esz) % 16) == 0, "C rows must be 16-byte aligned");
  D3D_REQUIRE(!a.residual || (((uintptr_t)a.residual % 16) == 0 && (a.ldres % 4) == 0), "residual rows must be 16-byte aligned");
  D3D_REQUIRE(!a.bias || ((uintptr_t)a.bias % 16) == 0, "bias must be 16-byte aligned");
  return 0;
}

}  // namespace

template <int BN>
int launch_pair(const d3d_gemm_args& a, cudaStream_t st) {
  using C = Cfg<BN, 2>;
  static int max_clusters = -1;
  if (max_clusters < 0) {
    D3D_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(d3d_num_sms() & ~1)); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_pair_kernel<BN>, &cfg) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = d3d_num_sms() / 2; }
    max_clusters = n < d3d_num_sms() / 2 ? n : d3d_num_sms() / 2;
  }
  CUtensorMap tmA, tmB;
  D3D_TRY(make_tmap(&tmA, a.A, a.in_kind, a.M, a.K, a.lda, BM));
  D3D_TRY(make_tmap(&tmB, a.W, a.in_kind, a.N, a.K, a.ldw, C::B_ROWS));
  Epilogue ep{a.C, a.ldc, a.bias, a.residual, a.ldres, a.act, a.out_kind};
  const int tiles = d3d_cdiv(a.M, 2 * BM) * d3d_cdiv(a.N, BN);
  const int cap = max_clusters < usable_sms() / 2 ? max_clusters : usable_sms() / 2;
  const int clusters = tiles < cap ? tiles : cap;
  gemm_tcgen05_pair_kernel<BN><<<2 * clusters, NUM_THREADS, C::SMEM_BYTES, st>>>(tmA, tmB, ep, a.M, a.N, a.K, a.in_kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

static int g_gemm_pair_mode = -1;  // -1: read D3D_GEMM_PAIR from the environment (default on), 0 off, 1 on
extern "C" int d3d_gemm_set_pair_mode(int mode) { g_gemm_pair_mode = mode; return 0; }
extern "C" int d3d_gemm_set_sm_limit(int n) { g_gemm_sm_limit = n < 0 ? 0 : n; return 0; }

namespace {
struct GemmRec { cudaEvent_t e0, e1; double flops; int variant; };
struct GemmProf { bool on = false; std::vector<GemmRec> recs; } g_gprof;
int gemm_dispatch(const d3d_gemm_args& a, cudaStream_t st, int* variant);
}  // namespace

extern "C" int d3d_gemm_profile_begin(void) {
  for (auto& r : g_gprof.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_gprof.recs.clear();
  g_gprof.on = true;
  return 0;
}
extern "C" int d3d_gemm_profile_end(double* flops, float* ms, int* launches) {
  g_gprof.on = false;
  D3D_CHECK_CUDA(cudaDeviceSynchronize());
  for (int v = 0; v < 3; ++v) { flops[v] = 0.0; ms[v] = 0.f; launches[v] = 0; }
  for (auto& r : g_gprof.recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms[r.variant] += t; flops[r.variant] += r.flops; ++launches[r.variant]; }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_gprof.recs.clear();
  return 0;
}

extern "C" int d3d_gemm(const d3d_gemm_args* args_h, void* stream) {
  D3D_REQUIRE(args_h != nullptr, "args");
  const d3d_gemm_args& a = *args_h;
  D3D_TRY(validate(a));
  cudaStream_t st = (cudaStream_t)stream;
  int variant = 0;
  if (!g_gprof.on) return gemm_dispatch(a, st, &variant);
  GemmRec r{nullptr, nullptr, 2.0 * a.M * a.N * a.K, 0};
  D3D_CHECK_CUDA(cudaEventCreate(&r.e0));
  D3D_CHECK_CUDA(cudaEventCreate(&r.e1));
  D3D_CHECK_CUDA(cudaEventRecord(r.e0, st));
  const int rc = gemm_dispatch(a, st, &r.variant);
  D3D_CHECK_CUDA(cudaEventRecord(r.e1, st));
  g_gprof.recs.push_back(r);
  return rc;
}

namespace {
int gemm_dispatch(const d3d_gemm_args& a, cudaStream_t st, int* variant) {
  // wide tiles once there is enough work to fill the machine with them; 128-wide otherwise
  const long long tiles256 = (long long)d3d_cdiv(a.M, BM) * d3d_cdiv(a.N, 256);
  if (g_gemm_pair_mode < 0) {
    const char* e = getenv("D3D_GEMM_PAIR");
    g_gemm_pair_mode = (e && e[0] == '0') ? 0 : 1;
  }
  // CTA pairs (cta_group::2, 256 x 256 tiles) once every pair has at least ~2 tiles
  if (g_gemm_pair_mode == 1 && a.N >= 256 && (long long)d3d_cdiv(a.M, 2 * BM) * d3d_cdiv(a.N, 256) >= (long long)d3d_num_sms()) {
    *variant = 0;
    return launch_pair<256>(a, st);
  }
  if (a.N >= 256 && tiles256 >= 2LL * d3d_num_sms()) { *variant = 1; return launch<256>(a, st); }
  *variant = 2;
  return launch<128>(a, st);
}
}  // namespace

extern "C" int d3d_gemm_simt(const d3d_gemm_args* args_h, void* stream) {
  D3D_REQUIRE(args_h != nullptr, "args");
  const d3d_gemm_args& a = *args_h;
  D3D_TRY(validate(a));
  const int nout = a.act == D3D_ACT_SWIGLU ? a.N / 2 : a.N;
  const long long total = (long long)a.M * nout;
  gemm_simt_kernel<<<d3d_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
  D3D_CHECK_LAUNCH();
  return 0;
}
