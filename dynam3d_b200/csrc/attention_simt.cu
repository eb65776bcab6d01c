// Variable-length multi-head self-attention, fp32 CUDA-core version (online softmax).
// Used for the small pooled encoders (patch->instance, instance->zone: sequences of 1..577 tokens, FF:595,683,728,754)
// and as the exact-arithmetic fallback shape for ViT / Phi-3 attention when the tensor-core kernel does not apply.
// Packed layout: qkv [T, 3*H*D] 16-bit (q | k | v), out [T, H*D] 16-bit; sequences given by cu_seqlens.
#include "common.cuh"

namespace {

constexpr int QT = 32;      // queries per block (8 per warp)
constexpr int KC = 64;      // keys per shared-memory chunk
constexpr int NWARP = 4;
constexpr int QPW = QT / NWARP;

template <int D>
__global__ void __launch_bounds__(NWARP * 32) attn_simt_kernel(const void* __restrict__ qkv, long long ld, void* __restrict__ out,
                                                               long long ldo, const int* __restrict__ cu, int H, int causal, int kind,
                                                               float scale) {
  constexpr int DPL = D / 32;  // output dims per lane
  __shared__ __half2 sK[KC][D / 2 + 1];
  __shared__ __half2 sV[KC][D / 2];
  __shared__ float sQ[QT][D];
  __shared__ float sP[NWARP][KC];
  const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int b = cu[seq], len = cu[seq + 1] - b;
  if (q0 >= len) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t qoff = (size_t)h * D, koff = (size_t)(H + h) * D, voff = (size_t)(2 * H + h) * D;

  for (int i = threadIdx.x; i < QT * D; i += blockDim.x) {
    const int qi = i / D, d = i % D;
    sQ[qi][d] = (q0 + qi < len) ? ld16(qkv, (size_t)(b + q0 + qi) * ld + qoff + d, kind) * scale : 0.f;
  }
  float m[QPW], l[QPW], o[QPW][DPL];
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < DPL; ++j) o[i][j] = 0.f;
  }
  const int kmax = causal ? min(len, q0 + QT) : len;
  for (int c0 = 0; c0 < kmax; c0 += KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < KC * (D / 2); i += blockDim.x) {
      const int r = i / (D / 2), d2 = i % (D / 2);
      __half2 kv = __float2half2_rn(0.f), vv = kv;
      if (c0 + r < len) {
        const size_t row = (size_t)(b + c0 + r) * ld;
        if (kind == D3D_BF16) {
          const __nv_bfloat162 kb = reinterpret_cast<const __nv_bfloat162*>((const __nv_bfloat16*)qkv + row + koff)[d2];
          const __nv_bfloat162 vb = reinterpret_cast<const __nv_bfloat162*>((const __nv_bfloat16*)qkv + row + voff)[d2];
          // keep bf16 values exactly: stage as two floats is too large; bf16 -> fp16 would round, so store raw bits and
          // reinterpret on use (see LOADK below)
          kv = *reinterpret_cast<const __half2*>(&kb);
          vv = *reinterpret_cast<const __half2*>(&vb);
        } else {
          kv = reinterpret_cast<const __half2*>((const __half*)qkv + row + koff)[d2];
          vv = reinterpret_cast<const __half2*>((const __half*)qkv + row + voff)[d2];
        }
      }
      sK[r][d2] = kv;
      sV[r][d2] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const int ql = warp * QPW + qi;
      const int q = q0 + ql;
      if (q >= len) continue;                 // warp-uniform
      if (causal && c0 > q) continue;         // whole chunk masked
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int d2 = 0; d2 < D / 2; ++d2) {
        const uint32_t ka = *reinterpret_cast<const uint32_t*>(&sK[lane][d2]);
        const uint32_t kb = *reinterpret_cast<const uint32_t*>(&sK[lane + 32][d2]);
        const float2 k0 = unpack16x2(ka, kind), k1 = unpack16x2(kb, kind);
        const float2 qv = *reinterpret_cast<const float2*>(&sQ[ql][2 * d2]);
        s0 = fmaf(qv.x, k0.x, s0); s0 = fmaf(qv.y, k0.y, s0);
        s1 = fmaf(qv.x, k1.x, s1); s1 = fmaf(qv.y, k1.y, s1);
      }
      const int j0 = c0 + lane, j1 = c0 + lane + 32;
      const bool ok0 = j0 < len && (!causal || j0 <= q);
      const bool ok1 = j1 < len && (!causal || j1 <= q);
      if (!ok0) s0 = -INFINITY;
      if (!ok1) s1 = -INFINITY;
      const float mx = warp_max(fmaxf(s0, s1));
      const float m_new = fmaxf(m[qi], mx);
      const float p0 = ok0 ? __expf(s0 - m_new) : 0.f;
      const float p1 = ok1 ? __expf(s1 - m_new) : 0.f;
      const float corr = (m[qi] == -INFINITY) ? 0.f : __expf(m[qi] - m_new);
      l[qi] = l[qi] * corr + warp_sum(p0 + p1);
      m[qi] = m_new;
      sP[warp][lane] = p0;
      sP[warp][lane + 32] = p1;
      __syncwarp();
      float acc[DPL];
#pragma unroll
      for (int j = 0; j < DPL; ++j) acc[j] = o[qi][j] * corr;
      const int jn = min(KC, kmax - c0);
      for (int j = 0; j < jn; ++j) {
        const float p = sP[warp][j];
        const __half* vrow = reinterpret_cast<const __half*>(&sV[j][0]);
#pragma unroll
        for (int t = 0; t < DPL; ++t) {
          const uint16_t raw = reinterpret_cast<const uint16_t*>(vrow)[lane + 32 * t];
          const float vv = kind == D3D_BF16 ? __uint_as_float((uint32_t)raw << 16) : __half2float(__ushort_as_half(raw));
          acc[t] = fmaf(p, vv, acc[t]);
        }
      }
#pragma unroll
      for (int j = 0; j < DPL; ++j) o[qi][j] = acc[j];
      __syncwarp();
    }
  }
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const int q = q0 + warp * QPW + qi;
    if (q >= len) continue;
    const float inv = 1.0f / l[qi];
#pragma unroll
    for (int t = 0; t < DPL; ++t) st16(out, (size_t)(b + q) * ldo + (size_t)h * D + lane + 32 * t, o[qi][t] * inv, kind);
  }
}

}  // namespace

extern "C" int d3d_attention_simt(const void* qkv, int64_t ld, void* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                                  int Dh, int causal, int kind, float scale, void* stream) {
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(qkv && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == 64 || Dh == 96, "head_dim 64 or 96");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
  dim3 grid(d3d_cdiv(max_len, QT), H, n_seq);
  cudaStream_t st = (cudaStream_t)stream;
  if (Dh == 64) attn_simt_kernel<64><<<grid, NWARP * 32, 0, st>>>(qkv, ld, out, ldo, cu_seqlens, H, causal, kind, scale);
  else attn_simt_kernel<96><<<grid, NWARP * 32, 0, st>>>(qkv, ld, out, ldo, cu_seqlens, H, causal, kind, scale);
  D3D_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32-in / fp32-out variant for the "precise" pipeline (no 16-bit rounding of q, k, v or the output)
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int KC32 = 32;

template <int D>
__global__ void __launch_bounds__(NWARP * 32) attn_simt_f32_kernel(const float* __restrict__ qkv, long long ld, float* __restrict__ out,
                                                                   long long ldo, const int* __restrict__ cu, int H, int causal, float scale) {
  constexpr int DPL = D / 32;
  __shared__ float sK[KC32][D + 1];
  __shared__ float sV[KC32][D];
  __shared__ float sQ[QT][D];
  __shared__ float sP[NWARP][KC32];
  const int seq = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int b = cu[seq], len = cu[seq + 1] - b;
  if (q0 >= len) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t qoff = (size_t)h * D, koff = (size_t)(H + h) * D, voff = (size_t)(2 * H + h) * D;
  for (int i = threadIdx.x; i < QT * D; i += blockDim.x) {
    const int qi = i / D, d = i % D;
    sQ[qi][d] = (q0 + qi < len) ? qkv[(size_t)(b + q0 + qi) * ld + qoff + d] * scale : 0.f;
  }
  float m[QPW], l[QPW], o[QPW][DPL];
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < DPL; ++j) o[i][j] = 0.f;
  }
  const int kmax = causal ? min(len, q0 + QT) : len;
  for (int c0 = 0; c0 < kmax; c0 += KC32) {
    __syncthreads();
    for (int i = threadIdx.x; i < KC32 * D; i += blockDim.x) {
      const int r = i / D, d = i % D;
      const bool ok = c0 + r < len;
      const size_t row = (size_t)(b + c0 + r) * ld;
      sK[r][d] = ok ? qkv[row + koff + d] : 0.f;
      sV[r][d] = ok ? qkv[row + voff + d] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const int ql = warp * QPW + qi;
      const int q = q0 + ql;
      if (q >= len) continue;
      if (causal && c0 > q) continue;
      float s0 = 0.f;
#pragma unroll 8
      for (int d = 0; d < D; ++d) s0 = fmaf(sQ[ql][d], sK[lane][d], s0);
      const int j0 = c0 + lane;
      const bool ok0 = j0 < len && (!causal || j0 <= q);
      if (!ok0) s0 = -INFINITY;
      const float mx = warp_max(s0);
      const float m_new = fmaxf(m[qi], mx);
      const float p0 = ok0 ? expf(s0 - m_new) : 0.f;
      const float corr = (m[qi] == -INFINITY) ? 0.f : expf(m[qi] - m_new);
      l[qi] = l[qi] * corr + warp_sum(p0);
      m[qi] = m_new;
      sP[warp][lane] = p0;
      __syncwarp();
      float acc[DPL];
#pragma unroll
      for (int j = 0; j < DPL; ++j) acc[j] = o[qi][j] * corr;
      const int jn = min(KC32, kmax - c0);
      for (int j = 0; j < jn; ++j) {
        const float p = sP[warp][j];
#pragma unroll
        for (int t = 0; t < DPL; ++t) acc[t] = fmaf(p, sV[j][lane + 32 * t], acc[t]);
      }
#pragma unroll
      for (int j = 0; j < DPL; ++j) o[qi][j] = acc[j];
      __syncwarp();
    }
  }
#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const int q = q0 + warp * QPW + qi;
    if (q >= len) continue;
    const float inv = 1.0f / l[qi];
#pragma unroll
    for (int t = 0; t < DPL; ++t) out[(size_t)(b + q) * ldo + (size_t)h * D + lane + 32 * t] = o[qi][t] * inv;
  }
}
}  // namespace

extern "C" int d3d_attention_f32(const float* qkv, int64_t ld, float* out, int64_t ldo, const int* cu_seqlens, int n_seq, int max_len, int H,
                                 int Dh, int causal, float scale, void* stream) {
  if (n_seq == 0 || max_len == 0) return 0;
  D3D_REQUIRE(qkv && out && cu_seqlens, "args");
  D3D_REQUIRE(Dh == 64 || Dh == 96, "head_dim 64 or 96");
  D3D_REQUIRE(n_seq <= 65535 && H <= 65535, "grid limits");
  dim3 grid(d3d_cdiv(max_len, QT), H, n_seq);
  cudaStream_t st = (cudaStream_t)stream;
  if (Dh == 64) attn_simt_f32_kernel<64><<<grid, NWARP * 32, 0, st>>>(qkv, ld, out, ldo, cu_seqlens, H, causal, scale);
  else attn_simt_f32_kernel<96><<<grid, NWARP * 32, 0, st>>>(qkv, ld, out, ldo, cu_seqlens, H, causal, scale);
  D3D_CHECK_LAUNCH();
  return 0;
}
