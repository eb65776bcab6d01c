// tcgen05 flash attention (forward, non-causal or causal, head_dim 64) over packed variable-length sequences.
//
//   S = Q K^T      : tcgen05.mma kind::f16, M=128 (queries) x N=128 (keys) x K=64, Q and K tiles TMA-loaded (SWIZZLE_128B,
//                    K-major) straight out of the packed [T, 3*H*64] QKV matrix, accumulator in TMEM (128 columns)
//   softmax        : 4 warps, one query row per thread (TMEM lane == row): tcgen05.ld of S, running max / sum in registers,
//                    P written to shared memory in the UMMA K-major SWIZZLE_128B layout
//   O_j = P V      : tcgen05.mma M=128 x N=64 x K=128 with V as an MN-major B operand (V rows are keys, d contiguous),
//                    accumulator in TMEM (64 columns); the softmax warps fold O_j into their fp32 row with the running rescale
// One CTA = 128 queries of one (sequence, head); K/V double-buffered; 2 CTAs per SM (112 KB smem, 256 TMEM columns each) so one
// CTA's softmax overlaps the other's MMAs.  Replaces nn.MultiheadAttention's attention in CLIPM:181-183 and the LLaVA tower.
#include "common.cuh"

namespace {

constexpr int D = 64;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int NTHREADS = 256;
constexpr int TILE_BYTES = 128 * 64 * 2;                 // one [128 x 64] 16-bit tile = 16 KB
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = TILE_BYTES;                       // 2 stages
constexpr int SMEM_V = 3 * TILE_BYTES;                   // 2 stages
constexpr int SMEM_P = 5 * TILE_BYTES;                   // [128 x 128] = 2 atoms of [128 x 64]
constexpr int SMEM_BAR = 7 * TILE_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR + 128;               // + barriers; 2 CTAs/SM: 2 x (114816 + 1024 reserved) <= 228 KB
constexpr int TMEM_COLS = 256;                           // S: [0,128)  O: [128,192)

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B operand descriptor (8-row groups 1024 B apart) -- same as the GEMM's
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B operand descriptor for a [K rows x 64 (MN, contiguous)] tile with 128-byte rows:
// SBO = 1024 B between 8-row K groups; LBO = byte distance between 64-element MN atoms (only one atom here)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(TILE_BYTES >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int kind, int m, int n, int b_mn_major) {
  return (1u << 4) | ((uint32_t)kind << 7) | ((uint32_t)kind << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(NTHREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, uint16_t* __restrict__ out, long long ldo, const int* __restrict__ cu, int H,
               int causal, int kind, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = smem_u32(smem_raw);
  if (sbase & 1023u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bar0 = sbase + SMEM_BAR;
  // barriers: 0 q_full | 1,2 kv_full | 3,4 kv_empty | 5 s_full | 6 s_free | 7 p_ready | 8 o_full | 9 o_free ; tmem ptr at +96
  auto bar = [&](int i) { return bar0 + 8u * i; };
  const uint32_t tmem_ptr_addr = bar0 + 96;

  const int seq = blockIdx.z, h = blockIdx.y;
  const int b = cu[seq], len = cu[seq + 1] - b;
  const int qt = causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int q0 = qt * BQ;
  if (q0 >= len) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kmax = causal ? min(len, q0 + BQ) : len;
  const int n_tiles = (kmax + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_qkv) : "memory");
  if (warp == 1 && lane == 0) {
    mbar_init(bar(0), 1);
    mbar_init(bar(1), 1); mbar_init(bar(2), 1);
    mbar_init(bar(3), 1); mbar_init(bar(4), 1);
    mbar_init(bar(5), 1);
    mbar_init(bar(6), 4);
    mbar_init(bar(7), 4);
    mbar_init(bar(8), 1);
    mbar_init(bar(9), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(bar(0), TILE_BYTES);
      tma_load_2d(sbase + SMEM_Q, &tm_qkv, bar(0), h * D, b + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        mbar_wait(bar(3 + st), (((uint32_t)j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(bar(1 + st), 2 * TILE_BYTES);
        tma_load_2d(sbase + SMEM_K + st * TILE_BYTES, &tm_qkv, bar(1 + st), (H + h) * D, b + j * BKV);
        tma_load_2d(sbase + SMEM_V + st * TILE_BYTES, &tm_qkv, bar(1 + st), (2 * H + h) * D, b + j * BKV);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc(kind, BQ, BKV, 0);
      const uint32_t idesc_o = make_idesc(kind, BQ, D, 1);
      const uint64_t dq = desc_kmajor(sbase + SMEM_Q);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        mbar_wait(bar(1 + st), ((uint32_t)j >> 1) & 1u);  // K_j, V_j landed
        tc_fence_after();
        const uint64_t dk = desc_kmajor(sbase + SMEM_K + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) umma_f16(tmem_s, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
        umma_commit(bar(5));
      };
      mbar_wait(bar(0), 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j + 1 < n_tiles) {
          mbar_wait(bar(6), (uint32_t)j & 1u);  // softmax has pulled S_j out of TMEM
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(bar(7), (uint32_t)j & 1u);      // P_j is in shared memory
        if (j > 0) mbar_wait(bar(9), (uint32_t)(j - 1) & 1u);  // O_{j-1} has been folded into registers
        tc_fence_after();
        const uint64_t dp = desc_kmajor(sbase + SMEM_P);
        const uint64_t dv = desc_mnmajor(sbase + SMEM_V + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // A = P: 16 keys = 32 B inside a 64-key atom (+2), next atom +16 KB; B = V: 16 key rows = 2048 B (+128)
          const uint64_t a = dp + (uint64_t)((k & 3) * 2 + (k >> 2) * (TILE_BYTES >> 4));
          umma_f16(tmem_o, a, dv + (uint64_t)(k * 128), idesc_o, k ? 1u : 0u);
        }
        umma_commit(bar(3 + st));  // K_j / V_j stage free
        umma_commit(bar(8));       // O_j ready
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / correction / epilogue: one query row per thread =====================
    const int qd = warp & 3;
    const int r = qd * 32 + lane;        // row inside the tile == TMEM lane
    const int row = q0 + r;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    const uint32_t p_row = sbase + SMEM_P + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    for (int j = 0; j < n_tiles; ++j) {
      const int c0 = j * BKV;
      mbar_wait(bar(5), (uint32_t)j & 1u);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int key = c0 + c + i;
          const bool ok = key < len && (!causal || key <= row);
          mx = fmaxf(mx, ok ? __uint_as_float(v[i]) : -INFINITY);
        }
      }
      const float m_new = fmaxf(m, mx * scale_log2);
      const float base = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f(m - base);
      m = m_new;
      // pass 2: P = exp2(S*scale - m) -> 16-bit -> swizzled shared memory; row sum
      float rs = 0.f;
      // (the P buffer is free: this thread already waited for O_{j-1}, i.e. PV_{j-1} has retired)
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
        if (c == BKV - 32) {  // S_j fully read: the MMA warp may overwrite it with S_{j+1}
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(6));
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int key = c0 + c + i;
          const bool ok0 = key < len && (!causal || key <= row);
          const bool ok1 = key + 1 < len && (!causal || key + 1 <= row);
          const float p0 = ok0 ? exp2f(__uint_as_float(v[i]) * scale_log2 - base) : 0.f;
          const float p1 = ok1 ? exp2f(__uint_as_float(v[i + 1]) * scale_log2 - base) : 0.f;
          rs += p0 + p1;
          pk[i >> 1] = pack16x2(p0, p1, kind);
        }
        // 32 keys = 4 chunks of 16 B; key c -> atom c/64, chunk (c%64)/8 XOR (row%8)
        const uint32_t atom = p_row + (uint32_t)(c >> 6) * (uint32_t)TILE_BYTES;
        const uint32_t ch0 = (uint32_t)((c & 63) >> 3);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t addr = atom + (((ch0 + q) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]),
                       "r"(pk[4 * q + 3])
                       : "memory");
        }
      }
      l = l * corr + rs;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(7));
      // fold the previous tile's O into the running row while the tensor core works on PV_j:  o = o*corr_prev ... (done below)
      // wait for O_j and accumulate
      mbar_wait(bar(8), (uint32_t)j & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < D; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_o + lane_off + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = o[c + i] * corr + __uint_as_float(v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(9));
    }
    if (row < len) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(b + row) * ldo + (size_t)h * D);
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        dst[i] = make_uint4(pack16x2(o[8 * i] * inv, o[8 * i + 1] * inv, kind), pack16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv, kind),
                            pack16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv, kind), pack16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv, kind));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

extern "C" int d3d_attention_tc(const void* qkv, int64_t ld, int64_t n_rows, void* out, int64_t ldo, const int* cu_seqlens, int n_seq,
                                int max_len, int H, int Dh, int causal, int kind, float scale, void* stream) {
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(qkv && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == D, "tcgen05 attention is built for head_dim 64");
  D3D_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "16-byte aligned rows");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
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
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)(3 * H * D), (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&tm, kind == D3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)qkv, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    d3d_set_error("cuTensorMapEncodeTiled(qkv) failed (%d)", (int)r);
    return D3D_ECUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    D3D_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid(d3d_cdiv(max_len, BQ), H, n_seq);
  attn_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tm, (uint16_t*)out, ldo, cu_seqlens, H, causal, kind,
                                                                      scale * 1.4426950408889634f);
  D3D_CHECK_LAUNCH();
  return 0;
}
