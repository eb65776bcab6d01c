// tcgen05 flash attention (forward, non-causal or causal, head_dim 64) over packed variable-length sequences.
//
//   S = Q K^T      : tcgen05.mma kind::f16, M=128 (queries) x N=128 (keys) x K=64, Q and K tiles TMA-loaded (SWIZZLE_128B,
//                    K-major) straight out of the packed [T, 3*H*64] QKV matrix, accumulator in TMEM (128 columns)
//   softmax        : 8 warps, two threads per query row (TMEM lane == row; each owns 64 of the tile's 128 keys): tcgen05.ld of S,
//                    running max / sum in registers, ex2.approx, P written to shared memory in the UMMA K-major SWIZZLE_128B layout.
//                    The two key halves of a row are INDEPENDENT online softmaxes (own base, own sum, own O accumulator), merged
//                    once in the epilogue -- no per-tile exchange or block barrier between the halves.
//   O_j = P V      : per key half, tcgen05.mma M=128 x N=64 x K=64 with V as an MN-major B operand (V rows are keys, d contiguous),
//                    two accumulators in TMEM (2 x 64 columns) that stay there over all key tiles (lazy rescale)
// One CTA = 128 queries of one (sequence, head); K/V double-buffered; 2 CTAs per SM (112 KB smem, 256 TMEM columns each) so one
// CTA's softmax overlaps the other's MMAs.  Replaces nn.MultiheadAttention's attention in CLIPM:181-183 and the LLaVA tower.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int ATOM64 = 128 * 64 * 2;                     // [128 rows x 64 columns] 16-bit, SWIZZLE_128B: 16 KB
constexpr int ATOM32 = 128 * 32 * 2;                     // [128 rows x 32 columns] 16-bit, SWIZZLE_64B: 8 KB (head_dim 96 = 64 + 32)
// Per (head dim, key halves).  HV = key halves per tile = softmax threads per query row: the key tile has 64 * HV rows.
//   HV = 2 (round 1/2a): 128-key tiles, two threads per row with independent online softmaxes, 2 CTAs per SM at D = 64.
//   HV = 1: 64-key tiles, one thread per row, half the shared memory and TMEM (S 64 + O D columns): three CTAs per SM at D = 64, TWO at
//           D = 96 (where HV = 2 fits only one).  Measured (tools/attn_halves.py): Phi-3 prefill shape 0.113 -> 0.088 ms (default there);
//           ViT shape 0.353 vs 0.357 ms -- a third resident CTA buys nothing, the per-tile hand-offs double (default stays HV = 2).
// D = 96 (Phi-3): every tile is a 64-column atom followed by a 32-column atom.
// SPLIT = 1 (the <= 1e-3 "precise" mode): every operand is a pair of fp16 matrices (hi | lo, x ~ hi + lo to ~2^-22), so that
//   S = Qh Kh^T + Ql Kh^T + Qh Kl^T   and   O += Ph Vh + Pl Vh + Ph Vl
// accumulate in fp32 in the SAME TMEM columns: three MMA groups instead of one, twice the operand tiles in shared memory, P written as
// (hi, lo) by the softmax warps, fp32 output.  HV = 1 only.  Replaces the mma.sync kernel of attention_split.cu for sequences >= 256.
template <int D, int HV, int SPLIT = 0>
struct AC {
  static_assert(D == 64 || D == 96, "head_dim 64 or 96");
  static_assert(HV == 1 || HV == 2, "key halves");
  static_assert(SPLIT == 0 || HV == 1, "split operands: 64-key tiles");
  static constexpr int NP = SPLIT ? 2 : 1;               // operand parts (hi, lo)
  static constexpr bool HAS32 = D == 96;
  static constexpr int BKV = 64 * HV;
  static constexpr int NSOFT = 4 * HV;
  static constexpr int NTHREADS = 64 + 32 * NSOFT;       // warp 0: TMEM alloc + TMA; warp 1: MMA issue; then the softmax warps
  static constexpr int Q_BYTES = ATOM64 + (HAS32 ? ATOM32 : 0);
  static constexpr int KV_ATOM64 = BKV * 64 * 2;         // [BKV rows x 64 columns]
  static constexpr int KV_ATOM32 = BKV * 32 * 2;
  static constexpr int KV_BYTES = KV_ATOM64 + (HAS32 ? KV_ATOM32 : 0);
  static constexpr int SMEM_Q = 0;                       // NP parts
  static constexpr int SMEM_K = NP * Q_BYTES;            // 2 stages x NP parts
  static constexpr int SMEM_V = SMEM_K + 2 * NP * KV_BYTES;
  static constexpr int SMEM_P = SMEM_V + 2 * NP * KV_BYTES;  // [128 x BKV] = HV atoms of [128 x 64], x NP parts
  static constexpr int SMEM_BAR = SMEM_P + NP * HV * ATOM64;
  static constexpr int SMEM_BYTES = SMEM_BAR + 640;
  static constexpr int TMEM_O = 64 * HV;                 // S: [0, 64 HV)   O of key half h: [64 HV + h D, 64 HV + (h + 1) D)
  static constexpr int TMEM_NEED = 64 * HV + HV * D;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static constexpr int CTAS_PER_SM = 512 / TMEM_COLS < (233472 / (SMEM_BYTES + 1024)) ? 512 / TMEM_COLS : (233472 / (SMEM_BYTES + 1024));
};
static_assert(AC<64, 2>::CTAS_PER_SM == 2 && AC<64, 1>::CTAS_PER_SM == 3, "CTAs per SM at head_dim 64");
static_assert(AC<96, 2>::CTAS_PER_SM == 1 && AC<96, 1>::CTAS_PER_SM == 2, "CTAs per SM at head_dim 96");
static_assert(AC<96, 1, 1>::CTAS_PER_SM == 1 && AC<64, 1, 1>::CTAS_PER_SM == 1, "split operands: one CTA per SM");

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
// same wait with a suspend-time hint: the polling warps (8 softmax warps per CTA) give their issue slots to the other CTA
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_S:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE_S;\n\t"
      "bra WAIT_LOOP_S;\n\t"
      "DONE_S:\n\t"
      "}" ::"r"(bar),
      "r"(parity), "r"(0x989680)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map),
               "r"(bar), "r"(c0), "r"(c1)
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
#ifdef D3D_ATTN_STAMPS
// debug build only (make EXTRA=-DD3D_ATTN_STAMPS): per-phase clock64 stamps of one CTA, read back by tools/attn_stamps.py
__device__ long long g_stamps[256];
#define STAMP(cond, slot) do { if (cond) g_stamps[(slot)] = clock64(); } while (0)
#else
#define STAMP(cond, slot) do { } while (0)
#endif

// packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2): halves the issue slots of the softmax inner loops
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};" ::"r"(r[0]),
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
      "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B operand descriptor (8-row groups 1024 B apart) -- same as the GEMM's
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand descriptor for a [K rows x 64 (MN, contiguous)] tile with 128-byte rows:
// SBO = 1024 B between 8-row K groups; LBO = byte distance between 64-element MN atoms (only one atom here)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(ATOM64 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// the same two descriptors for the 32-column atom (64-byte rows, SWIZZLE_64B = layout type 4, 8-row groups 512 B apart)
__device__ __forceinline__ uint64_t desc_kmajor_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(ATOM32 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int kind, int m, int n, int b_mn_major) {
  return (1u << 4) | ((uint32_t)kind << 7) | ((uint32_t)kind << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// KIND (D3D_F16 / D3D_BF16) is a template parameter: as a run-time argument every P / output pack was emitted for both types and predicated
// (ncu, round 2: F2FP = 8 % of the issued instructions, half of them predicated off)
template <int D, int KIND, int HV, int SPLIT>
__global__ void __launch_bounds__((AC<D, HV, SPLIT>::NTHREADS), (AC<D, HV, SPLIT>::CTAS_PER_SM))
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_qkv32, const __grid_constant__ CUtensorMap tm_kv,
               const __grid_constant__ CUtensorMap tm_kv32, uint16_t* __restrict__ out, long long ldo, const int* __restrict__ cu,
               const int* __restrict__ lens, int q_tile_begin, int H, int causal, float scale_log2, int lo_off, float* __restrict__ out32) {
  constexpr int kind = KIND;
  using C = AC<D, HV, SPLIT>;
  constexpr int BKV = C::BKV, NSOFT = C::NSOFT, NP = C::NP;
  constexpr int TILE_BYTES = C::KV_BYTES, Q_BYTES = C::Q_BYTES, KV_ATOM64 = C::KV_ATOM64;
  constexpr int SMEM_Q = C::SMEM_Q, SMEM_K = C::SMEM_K, SMEM_V = C::SMEM_V, SMEM_P = C::SMEM_P, SMEM_BAR = C::SMEM_BAR;
  constexpr int TMEM_COLS = C::TMEM_COLS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bar0 = sbase + SMEM_BAR;
  // barriers: 0 q_full | 1,2 kv_full | 3,4 kv_empty | 5 s_full | 6 s_free | 7 p_ready (and O rescaled) | 8 o_full ; tmem ptr at +96
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_ptr_addr = bar0 + 96;

  const int seq = blockIdx.z, h = blockIdx.y;
  // sequences are [cu[seq], cu[seq] + len): packed back to back (lens == NULL: len = cu[seq+1] - cu[seq]) or at arbitrary starts with explicit
  // lengths (the strided KV cache of the chunked prefill); this launch computes query tiles q_tile_begin + [0, gridDim.x) of every sequence
  const int b = cu[seq], len = lens ? lens[seq] : cu[seq + 1] - b;
  const int qt = (causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) + q_tile_begin;
  const int q0 = qt * BQ;
  if (q0 >= len) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kmax = causal ? min(len, q0 + BQ) : len;
  const int n_tiles = (kmax + BKV - 1) / BKV;

  if (warp == 1 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_qkv) : "memory");
    mbar_init(bar(0), 1);
    mbar_init(bar(1), 1); mbar_init(bar(2), 1);
    mbar_init(bar(3), 1); mbar_init(bar(4), 1);
    mbar_init(bar(5), 1);
    mbar_init(bar(6), NSOFT);
    mbar_init(bar(7), NSOFT);
    mbar_init(bar(8), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + C::TMEM_O;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // operand part p (0 = hi, 1 = lo of the split mode) lives lo_off columns to the right in the packed matrix
      mbar_expect_tx(bar(0), NP * Q_BYTES);
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        tma_load_2d(sbase + SMEM_Q + p * Q_BYTES, &tm_qkv, bar(0), p * lo_off + h * D, b + q0);
        if (C::HAS32) tma_load_2d(sbase + SMEM_Q + p * Q_BYTES + ATOM64, &tm_qkv32, bar(0), p * lo_off + h * D + 64, b + q0);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar(3 + st), (((uint32_t)j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(bar(1 + st), 2 * NP * TILE_BYTES);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t ko = sbase + SMEM_K + (st * NP + p) * TILE_BYTES, vo = sbase + SMEM_V + (st * NP + p) * TILE_BYTES;
          tma_load_2d(ko, &tm_kv, bar(1 + st), p * lo_off + (H + h) * D, b + j * BKV);
          tma_load_2d(vo, &tm_kv, bar(1 + st), p * lo_off + (2 * H + h) * D, b + j * BKV);
          if (C::HAS32) {
            tma_load_2d(ko + KV_ATOM64, &tm_kv32, bar(1 + st), p * lo_off + (H + h) * D + 64, b + j * BKV);
            tma_load_2d(vo + KV_ATOM64, &tm_kv32, bar(1 + st), p * lo_off + (2 * H + h) * D + 64, b + j * BKV);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
#ifdef D3D_ATTN_STAMPS
      const bool mst_on = blockIdx.x == 1 && blockIdx.y == 5 && blockIdx.z == 40;
#endif
      const uint32_t idesc_s = make_idesc(kind, BQ, BKV, 0);  // N = 64 or 128 keys
      const uint32_t idesc_o = make_idesc(kind, BQ, 64, 1);
      const uint32_t idesc_o32 = make_idesc(kind, BQ, 32, 1);
      // operand-part pairs (A part, B part): plain mode (hi, hi); split mode (hi, hi), (lo, hi), (hi, lo) -- for S the parts of (Q, K),
      // for O the parts of (P, V); all pairs accumulate into the same TMEM columns
      constexpr int NPAIR = SPLIT ? 3 : 1;
      constexpr int PA[3] = {0, 1, 0}, PB[3] = {0, 0, 1};
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(bar(1 + st), ((uint32_t)j >> 1) & 1u);  // K_j, V_j landed
        tc_fence_after();
#pragma unroll
        for (int pr = 0; pr < NPAIR; ++pr) {
          const uint32_t qs = sbase + SMEM_Q + PA[pr] * Q_BYTES, ks = sbase + SMEM_K + (st * NP + PB[pr]) * TILE_BYTES;
          const uint64_t dq = desc_kmajor(qs), dk = desc_kmajor(ks);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_s, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, (pr | k) ? 1u : 0u);
          if (C::HAS32) {  // head dims 64..95: the 32-column atom (two more K-steps)
            const uint64_t dq32 = desc_kmajor_sw64(qs + ATOM64), dk32 = desc_kmajor_sw64(ks + KV_ATOM64);
#pragma unroll
            for (int k = 0; k < 2; ++k) umma_f16(tmem_s, dq32 + (uint64_t)(2 * k), dk32 + (uint64_t)(2 * k), idesc_s, 1u);
          }
        }
        umma_commit(bar(5));
      };
      mbar_wait(bar(0), 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j + 1 < n_tiles) {
          mbar_wait(bar(6), (uint32_t)j & 1u);  // the softmax warps hold S_j in registers
          tc_fence_after();
          STAMP(mst_on && j < 8, 64 + j * 8 + 0);
          issue_s(j + 1);
          STAMP(mst_on && j < 8, 64 + j * 8 + 1);
        }
        mbar_wait(bar(7), (uint32_t)j & 1u);  // P_j is in shared memory and O has been rescaled where the running max moved
        STAMP(mst_on && j < 8, 64 + j * 8 + 2);
        tc_fence_after();
#pragma unroll
        for (int pr = 0; pr < NPAIR; ++pr) {
          const uint32_t vs = sbase + SMEM_V + (st * NP + PB[pr]) * TILE_BYTES;
          const uint64_t dp = desc_kmajor(sbase + SMEM_P + PA[pr] * HV * ATOM64);
          const uint64_t dv = desc_mnmajor(vs);
          const uint64_t dv32 = desc_mnmajor_sw64(vs + KV_ATOM64);
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            // A = P: 16 keys = 32 B inside a 64-key atom (+2), next atom +16 KB; B = V: 16 key rows = 2048 B (+128).
            // O accumulates in TMEM over all key tiles (the tensor pipe executes the products in issue order).
            const uint64_t a = dp + (uint64_t)((k & 3) * 2 + (k >> 2) * (ATOM64 >> 4));
            const uint32_t acc = (j | pr | (k & 3)) ? 1u : 0u;
            // keys 0..63 of the tile (k < 4) accumulate into O half 0, keys 64..127 into O half 1 (independent softmax bases)
            umma_f16(tmem_o + (uint32_t)((k >> 2) * D), a, dv + (uint64_t)(k * 128), idesc_o, acc);
            // channels 64..95: 16 key rows of the 32-column atom = 1024 B (+64)
            if (C::HAS32) umma_f16(tmem_o + (uint32_t)((k >> 2) * D + 64), a, dv32 + (uint64_t)(k * 64), idesc_o32, acc);
          }
        }
        umma_commit(bar(3 + st));  // K_j / V_j stage free
        umma_commit(bar(8));       // O += P_j V_j done: P buffer free, O readable
        STAMP(mst_on && j < 8, 64 + j * 8 + 3);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    // 4 HV warps: warp w owns TMEM lane quadrant w&3 (rows) and key half (w-2)>>2 (HV = 2: 64 of the 128 keys of a tile and half of the
    // output channels; HV = 1: the whole 64-key tile and all channels).  A thread holds its 64 scores in registers (one TMEM read per
    // tile); O stays in TMEM and is only rescaled when a row's running max moves by more than 2^8 (P stays within 16-bit range).
    const int qd = warp & 3;
    const int hf = HV == 2 ? (warp - 2) >> 2 : 0;
    const int r = qd * 32 + lane;        // row inside the tile == TMEM lane
    const int row = q0 + r;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    constexpr int DH = D / HV;  // output channels written by this thread
    float m = -INFINITY, l = 0.f;
    const uint32_t p_row = sbase + SMEM_P + (uint32_t)hf * (uint32_t)ATOM64 + (uint32_t)r * 128u;  // this half's 64-key atom
    const uint32_t sw = (uint32_t)(r & 7);
    // (base, sum) of the two key halves are exchanged (fp32) through the P buffer once the last P.V product has completed
    const uint32_t xsum_mine = sbase + SMEM_P + (uint32_t)(hf * 128 + r) * 8u;
    const uint32_t xsum_other = sbase + SMEM_P + (uint32_t)((hf ^ 1) * 128 + r) * 8u;
#ifdef D3D_ATTN_STAMPS
    const bool st_on = blockIdx.x == 1 && blockIdx.y == 5 && blockIdx.z == 40 && warp == 2 && lane == 0;
#endif
    const int lim_row = causal ? min(len, row + 1) : len;  // keys >= lim_row are masked for this row
    // one key tile; EDGE tiles (crossing the sequence end or the causal diagonal) pay for the per-element compare
    auto tile = [&](int j, auto edge_tag) {
      constexpr bool EDGE = decltype(edge_tag)::value;
      const int nv = lim_row - (j * BKV + hf * 64);  // this thread's keys [0, 64) of the tile half are valid below nv
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 64; c += 16) tmem_ld16(tmem_s + lane_off + (uint32_t)(hf * 64 + c), v + c);
      tmem_ld_wait();
      // S_j is in registers: the MMA warp may overwrite it with S_{j+1} once all 8 warps arrive
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(6));
      STAMP(st_on && j < 8, j * 8 + 2);
      // rows past the end of the sequence (last query tile) and keys past it (last key tile) take no part in the arithmetic below:
      // their exponentials are skipped (the kernel is MUFU-bound), their P entries are zero
      const int nvv = EDGE ? (row < len ? nv : 0) : 64;
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) mx = fmaxf(mx, (!EDGE || i < nvv) ? __uint_as_float(v[i]) : -INFINITY);
      STAMP(st_on && j < 8, j * 8 + 3);
      // lazy rescale: keep the old base while this half's row max grew by <= 8 (P <= 2^8)
      const float m_cand = fmaxf(m, mx * scale_log2);
      const bool move = (m_cand > m + 8.0f) || (m == -INFINITY && m_cand != -INFINITY);
      float resc = 1.f;
      if (move) {
        resc = ex2_approx(m - m_cand);  // 0 on the first valid tile (m = -inf)
        m = m_cand;
      }
      bool o_done = false;
      if (j > 0 && __any_sync(0xffffffffu, move)) {
        // O (TMEM) *= resc for the rows whose base moved; needs O += P_{j-1} V_{j-1} to have completed
        mbar_wait_sleepy(bar(8), (uint32_t)(j - 1) & 1u);
        tc_fence_after();
        o_done = true;
#pragma unroll
        for (int c = 0; c < D; c += 16) {  // this key half's own accumulator: all 64 channels of the row
          uint32_t ov[16];
          tmem_ld16(tmem_o + lane_off + (uint32_t)(hf * D + c), ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * resc);
          tmem_st16(tmem_o + lane_off + (uint32_t)(hf * D + c), ov);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      l *= resc;
      const float nbase = (m == -INFINITY) ? 0.f : -m;
      // P = 2^(S*scale - m) -> 16 bit (in place over the score registers); partial row sum
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t plo[SPLIT ? 32 : 1];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        if (EDGE && i >= nvv) {  // both keys masked: no exponentials
          v[i >> 1] = 0u;
          if constexpr (SPLIT != 0) plo[i >> 1] = 0u;
          continue;
        }
        float t0, t1;
        ffma2(t0, t1, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), scale_log2, scale_log2, nbase, nbase);
        float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
        if (EDGE) p1 = (i + 1 < nvv) ? p1 : 0.f;
        fadd2(rs0, rs1, rs0, rs1, p0, p1);
        const uint32_t hi = pack16x2(p0, p1, kind);
        v[i >> 1] = hi;
        if constexpr (SPLIT != 0) {  // P = hi + lo: the residual of the 16-bit rounding, itself rounded to 16 bit
          const float2 hf2 = unpack16x2(hi, kind);
          plo[i >> 1] = pack16x2(p0 - hf2.x, p1 - hf2.y, kind);
        }
      }
      l += rs0 + rs1;
      STAMP(st_on && j < 8, j * 8 + 4);
      // the P buffer is free once O += P_{j-1} V_{j-1} has completed
      if (j > 0 && !o_done) {
        mbar_wait_sleepy(bar(8), (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      STAMP(st_on && j < 8, j * 8 + 5);
#pragma unroll
      for (int q = 0; q < 8; ++q) {  // 8 x 16 B = this thread's 64 keys, K-major SWIZZLE_128B atom of this half
        const uint32_t addr = p_row + ((((uint32_t)q) ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]),
                     "r"(v[4 * q + 3])
                     : "memory");
        if constexpr (SPLIT != 0)  // the lo part: the next [128 x 64] atom
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + (uint32_t)(HV * ATOM64)), "r"(plo[4 * q]), "r"(plo[4 * q + 1]),
                       "r"(plo[4 * q + 2]), "r"(plo[4 * q + 3])
                       : "memory");
      }
    };
    for (int j = 0; j < n_tiles; ++j) {
      const bool edge = __any_sync(0xffffffffu, j * BKV + BKV > lim_row || row >= len);  // warp-uniform
      STAMP(st_on && j < 8, j * 8 + 0);
      mbar_wait_sleepy(bar(5), (uint32_t)j & 1u);
      tc_fence_after();
      STAMP(st_on && j < 8, j * 8 + 1);
      if (edge) tile(j, std::true_type{});
      else tile(j, std::false_type{});
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(7));
      STAMP(st_on && j < 8, j * 8 + 6);
    }
    // epilogue: O / l
    mbar_wait_sleepy(bar(8), (uint32_t)(n_tiles - 1) & 1u);
    tc_fence_after();
    // HV = 2: merge the two key halves of the row: exchange (base, sum) through the (now free) P buffer, then
    // out = (O_mine * 2^(m_mine - M) + O_other * 2^(m_other - M)) / (l_mine * 2^(m_mine - M) + l_other * 2^(m_other - M))
    float w1, w2 = 0.f;
    if (HV == 2) {
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(xsum_mine), "f"(m), "f"(l) : "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float m_other, l_other;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m_other), "=f"(l_other) : "r"(xsum_other) : "memory");
      const float mm = fmaxf(m, m_other);
      const float a_mine = (m == -INFINITY) ? 0.f : ex2_approx(m - mm);
      const float a_other = (m_other == -INFINITY) ? 0.f : ex2_approx(m_other - mm);
      const float lt = l * a_mine + l_other * a_other;
      const float inv = lt > 0.f ? 1.0f / lt : 0.f;
      w1 = a_mine * inv; w2 = a_other * inv;
    } else {
      w1 = l > 0.f ? 1.0f / l : 0.f;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(b + row) * ldo + (size_t)h * D + hf * DH);
#pragma unroll
    for (int c = 0; c < DH; c += 16) {  // 16 channels at a time (register budget)
      uint32_t o[16], o2[16];
      tmem_ld16(tmem_o + lane_off + (uint32_t)(hf * D + hf * DH + c), o);           // my key half's accumulator, my output channels
      if (HV == 2) tmem_ld16(tmem_o + lane_off + (uint32_t)((hf ^ 1) * D + hf * DH + c), o2);  // the other key half's accumulator, same channels
      tmem_ld_wait();
      if (row < len) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          auto f = [&](int k) { return HV == 2 ? __uint_as_float(o[8 * i + k]) * w1 + __uint_as_float(o2[8 * i + k]) * w2 : __uint_as_float(o[8 * i + k]) * w1; };
          if constexpr (SPLIT != 0) {  // fp32 rows
            float4* d32 = reinterpret_cast<float4*>(out32 + (size_t)(b + row) * ldo + (size_t)h * D + hf * DH + c + 8 * i);
            d32[0] = make_float4(f(0), f(1), f(2), f(3));
            d32[1] = make_float4(f(4), f(5), f(6), f(7));
          } else {
            dst[c / 8 + i] = make_uint4(pack16x2(f(0), f(1), kind), pack16x2(f(2), f(3), kind), pack16x2(f(4), f(5), kind), pack16x2(f(6), f(7), kind));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

#ifdef D3D_ATTN_STAMPS
extern "C" int d3d_debug_attn_stamps(long long* host_out) {
  D3D_CHECK_CUDA(cudaDeviceSynchronize());
  D3D_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_stamps, sizeof(long long) * 256));
  return 0;
}
#endif

namespace {
int g_attn_tc_halves[2] = {2, 1};  // key halves per tile for head_dim 64 / 96 (d3d_attention_tc_set_halves): measured, tools/attn_halves.py

template <int D, int HV, int SPLIT>
int launch_attn_tc(const CUtensorMap& tm, const CUtensorMap& tm32, const CUtensorMap& tmkv, const CUtensorMap& tmkv32, void* out, int64_t ldo,
                   const int* cu_seqlens, const int* lens, int n_seq, int max_len, int q_tile_begin, int q_tile_end, int H, int causal, int kind,
                   float scale, int lo_off, cudaStream_t st) {
  using C = AC<D, HV, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<D, D3D_F16, HV, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<D, D3D_BF16, HV, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int q_end = q_tile_end < d3d_cdiv(max_len, BQ) ? q_tile_end : d3d_cdiv(max_len, BQ);
  if (q_end <= q_tile_begin) return 0;
  dim3 grid(q_end - q_tile_begin, H, n_seq);
  uint16_t* o16 = SPLIT ? nullptr : (uint16_t*)out;
  float* o32 = SPLIT ? (float*)out : nullptr;
  if (kind == D3D_BF16)
    attn_tc_kernel<D, D3D_BF16, HV, SPLIT><<<grid, C::NTHREADS, C::SMEM_BYTES, st>>>(tm, tm32, tmkv, tmkv32, o16, ldo, cu_seqlens, lens, q_tile_begin, H, causal,
                                                                                    scale * 1.4426950408889634f, lo_off, o32);
  else
    attn_tc_kernel<D, D3D_F16, HV, SPLIT><<<grid, C::NTHREADS, C::SMEM_BYTES, st>>>(tm, tm32, tmkv, tmkv32, o16, ldo, cu_seqlens, lens, q_tile_begin, H, causal,
                                                                                   scale * 1.4426950408889634f, lo_off, o32);
  D3D_CHECK_LAUNCH();
  return 0;
}

// tensor maps over a packed [n_rows, n_cols] 16-bit matrix: query tiles 128 rows, key / value tiles kv_rows rows; each as a 64-column
// (SWIZZLE_128B) and a 32-column (SWIZZLE_64B; head_dim 96) atom
int attn_tc_maps(const void* base, int64_t ld, int64_t n_rows, int64_t n_cols, int kind, int kv_rows, CUtensorMap* tm, CUtensorMap* tm32, CUtensorMap* tmkv,
                 CUtensorMap* tmkv32) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
      d3d_set_error("cuTensorMapEncodeTiled entry point not available");
      return D3D_ECUDA;
    }
    fn = (EncodeTiledFn)p;
  }
  cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = kind == D3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  auto encode = [&](CUtensorMap* m, cuuint32_t cols, cuuint32_t rows) {
    cuuint32_t box[2] = {cols, rows};
    return fn(m, dt, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = encode(tm, 64, 128);
  if (r == CUDA_SUCCESS) r = encode(tm32, 32, 128);  // unused at head_dim 64, but the kernel signature is shared
  if (r == CUDA_SUCCESS) r = encode(tmkv, 64, (cuuint32_t)kv_rows);
  if (r == CUDA_SUCCESS) r = encode(tmkv32, 32, (cuuint32_t)kv_rows);
  if (r != CUDA_SUCCESS) {
    d3d_set_error("cuTensorMapEncodeTiled(qkv) failed (%d)", (int)r);
    return D3D_ECUDA;
  }
  return 0;
}
}  // namespace

// key halves per tile (1: 64-key tiles, one softmax thread per row, more CTAs per SM; 2: 128-key tiles, two threads per row) per head dim
extern "C" int d3d_attention_tc_set_halves(int halves_d64, int halves_d96) {
  D3D_REQUIRE((halves_d64 == 1 || halves_d64 == 2) && (halves_d96 == 1 || halves_d96 == 2), "halves are 1 or 2");
  g_attn_tc_halves[0] = halves_d64;
  g_attn_tc_halves[1] = halves_d96;
  return 0;
}

extern "C" int d3d_attention_tc(const void* qkv, int64_t ld, int64_t n_rows, void* out, int64_t ldo, const int* cu_seqlens, int n_seq,
                                int max_len, int H, int Dh, int causal, int kind, float scale, void* stream) {
  return d3d_attention_tc_ex(qkv, ld, n_rows, out, ldo, cu_seqlens, nullptr, n_seq, max_len, 0, 1 << 30, H, Dh, causal, kind, scale, stream);
}

extern "C" int d3d_attention_tc_ex(const void* qkv, int64_t ld, int64_t n_rows, void* out, int64_t ldo, const int* seq_start, const int* seq_len,
                                   int n_seq, int max_len, int q_tile_begin, int q_tile_end, int H, int Dh, int causal, int kind, float scale,
                                   void* stream) {
  const int* cu_seqlens = seq_start;
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(q_tile_begin >= 0 && q_tile_end >= q_tile_begin, "query tile range");
  D3D_REQUIRE(qkv && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == 64 || Dh == 96, "tcgen05 attention is built for head_dim 64 and 96");
  D3D_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "16-byte aligned rows");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
  const int hv = g_attn_tc_halves[Dh == 96 ? 1 : 0];
  CUtensorMap tm, tm32, tmkv, tmkv32;
  D3D_TRY(attn_tc_maps(qkv, ld, n_rows, 3LL * H * Dh, kind, 64 * hv, &tm, &tm32, &tmkv, &tmkv32));
  cudaStream_t st = (cudaStream_t)stream;
#define ATT_GO(DV, HVV) \
  return launch_attn_tc<DV, HVV, 0>(tm, tm32, tmkv, tmkv32, out, ldo, cu_seqlens, seq_len, n_seq, max_len, q_tile_begin, q_tile_end, H, causal, kind, scale, 0, st)
  if (Dh == 96) {
    if (hv == 1) ATT_GO(96, 1);
    ATT_GO(96, 2);
  }
  if (hv == 1) ATT_GO(64, 1);
  ATT_GO(64, 2);
#undef ATT_GO
}

// The <= 1e-3 "precise" mode on the tensor cores (tcgen05): qkv_hl = [T, lo_off + 3 H Dh] fp16, the hi parts of [q | k | v] in columns
// [0, 3 H Dh) and the lo parts lo_off columns to the right (d3d_split16 with 2 terms); out = fp32 [T, H Dh].  Same contract as
// d3d_attention_split (the mma.sync kernel, which stays for sequences < 256 tokens).
extern "C" int d3d_attention_split_tc(const void* qkv_hl, int64_t ld, int64_t n_rows, int64_t lo_off, float* out, int64_t ldo, const int* cu_seqlens,
                                      int n_seq, int max_len, int H, int Dh, int causal, float scale, void* stream) {
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(qkv_hl && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == 64 || Dh == 96, "tcgen05 attention is built for head_dim 64 and 96");
  D3D_REQUIRE(ld % 8 == 0 && ldo % 4 == 0 && lo_off % 8 == 0 && lo_off >= 3LL * H * Dh && ld >= lo_off + 3LL * H * Dh &&
                  ((uintptr_t)qkv_hl % 16) == 0 && ((uintptr_t)out % 16) == 0,
              "16-byte aligned rows, lo part behind the hi part");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
  CUtensorMap tm, tm32, tmkv, tmkv32;
  D3D_TRY(attn_tc_maps(qkv_hl, ld, n_rows, lo_off + 3LL * H * Dh, D3D_F16, 64, &tm, &tm32, &tmkv, &tmkv32));
  cudaStream_t st = (cudaStream_t)stream;
  if (Dh == 96)
    return launch_attn_tc<96, 1, 1>(tm, tm32, tmkv, tmkv32, out, ldo, cu_seqlens, nullptr, n_seq, max_len, 0, 1 << 30, H, causal, D3D_F16, scale, (int)lo_off, st);
  return launch_attn_tc<64, 1, 1>(tm, tm32, tmkv, tmkv32, out, ldo, cu_seqlens, nullptr, n_seq, max_len, 0, 1 << 30, H, causal, D3D_F16, scale, (int)lo_off, st);
}
