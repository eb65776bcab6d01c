// Token builders for the layer-wise pooling of the 3D memory (patch -> instance -> zone, FF:580-597, 662-676, 717-727,
// 743-753) and the merge-discriminator input (FF:613-617).  Sequences from several episodes are packed into one batch;
// every sequence carries its own base pointers (an episode's patch / instance pool), so one launch serves all episodes.
#include "common.cuh"

namespace {

// feature row (16-bit, 8 wide, K padded for the tensor-core GEMM):
//   mode 0 (patch -> instance, FF:584-591): [xyz - centre (3), |xyz|, sin(dir), cos(dir), scale, 0]
//   mode 1 (instance -> zone,  FF:719-723): [xyz - centre (3), |xyz|, 0, 0, 0, 0]; seq_dir[s] == 1 selects the voxel-centre keys (Q5)
// |xyz| is the norm of the ABSOLUTE position.  Aggregate-token rows (src < 0) are zero.
__global__ void pool_features_kernel(const long long* __restrict__ seq_xyz, const long long* __restrict__ seq_dir,
                                     const long long* __restrict__ seq_scale, const float* __restrict__ centre,
                                     const int* __restrict__ tok_seq, const int* __restrict__ tok_src, int T, int mode,
                                     void* __restrict__ out, int kind) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int src = tok_src[t];
  if (src >= 0) {
    const int s = tok_seq[t];
    const float* xyz = reinterpret_cast<const float*>(seq_xyz[s]) + (size_t)src * 3;
    float x = xyz[0], y = xyz[1], z = xyz[2];
    if (mode == 1 && seq_dir[s] == 1) {
      // Q5 (FF:739-741): an UPDATED zone is embedded from its members' voxel-centre keys floor(p / L) * L + L / 2, computed here from the
      // instance positions (L = float bits in seq_scale[s]) -- same fp32 expression as the host planner, no key array to build and upload
      const float L = __int_as_float((int)seq_scale[s]);
      const float hl = __fdiv_rn(L, 2.0f);
      x = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(x, L)), L), hl);
      y = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(y, L)), L), hl);
      z = __fadd_rn(__fmul_rn(floorf(__fdiv_rn(z, L)), L), hl);
    }
    f[0] = x - centre[(size_t)s * 3];
    f[1] = y - centre[(size_t)s * 3 + 1];
    f[2] = z - centre[(size_t)s * 3 + 2];
    f[3] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    if (mode == 0) {
      const float d = reinterpret_cast<const float*>(seq_dir[s])[src];
      f[4] = sinf(d);
      f[5] = cosf(d);
      f[6] = reinterpret_cast<const float*>(seq_scale[s])[src];
    }
  }
  if (kind == D3D_OUT_F32) {
#pragma unroll
    for (int i = 0; i < 8; ++i) ((float*)out)[(size_t)t * 8 + i] = f[i];
    return;
  }
  uint4 p = make_uint4(pack16x2(f[0], f[1], kind), pack16x2(f[2], f[3], kind), pack16x2(f[4], f[5], kind), pack16x2(f[6], f[7], kind));
  reinterpret_cast<uint4*>(out)[t] = p;
}

// X[t] = (src < 0) ? agg : emb[t] + fts[src]   (FF:592-593, 674-676, 725-727); fts rows are fp16 (patches) or fp32 (instances).
// X16 (optional) receives the same rows rounded to the GEMM operand type, so the first layer needs no separate cast pass.
__global__ void pool_assemble_kernel(const float* __restrict__ emb, const long long* __restrict__ seq_fts, int fts_is_f32,
                                     const int* __restrict__ tok_seq, const int* __restrict__ tok_src, const float* __restrict__ agg, int T,
                                     int D, float* __restrict__ X, void* __restrict__ X16, int kind) {
  const int t = blockIdx.x;
  if (t >= T) return;
  const int src = tok_src[t];
  float4* dst = reinterpret_cast<float4*>(X + (size_t)t * D);
  uint2* dst16 = X16 ? reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(X16) + (size_t)t * D) : nullptr;
  const int n4 = D >> 2;  // D % 4 == 0 (checked by the entry)
  const float4* e = reinterpret_cast<const float4*>(emb + (size_t)t * D);
  const int s = src < 0 ? 0 : tok_seq[t];
  const float4* f32 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(seq_fts[s]) + (size_t)(src < 0 ? 0 : src) * D);
  const uint2* f16 = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(seq_fts[s]) + (size_t)(src < 0 ? 0 : src) * D);
  for (int c = threadIdx.x; c < n4; c += blockDim.x) {
    float4 v;
    if (src < 0) {
      v = reinterpret_cast<const float4*>(agg)[c];
    } else {
      const float4 a = e[c];
      if (fts_is_f32) {
        const float4 b = f32[c];
        v = make_float4(b.x + a.x, b.y + a.y, b.z + a.z, b.w + a.w);
      } else {
        const uint2 b = f16[c];
        const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&b.x)), b1 = __half22float2(*reinterpret_cast<const __half2*>(&b.y));
        v = make_float4(b0.x + a.x, b0.y + a.y, b1.x + a.z, b1.y + a.w);
      }
    }
    dst[c] = v;
    if (dst16) dst16[c] = make_uint2(pack16x2(v.x, v.y, kind), pack16x2(v.z, v.w, kind));
  }
}

// merge-discriminator input rows (FF:613-617): row (g, j) = [inst_fts[idx[g,j]] | view_fts[g] | centre[g] - inst_pos[idx[g,j]] | 0 pad]
__global__ void disc_input_kernel(const float* __restrict__ inst_fts, const float* __restrict__ inst_pos, const int* __restrict__ idx,
                                  const float* __restrict__ view_fts, const float* __restrict__ centre, int G, int K, int D, int ldo,
                                  void* __restrict__ out, int kind) {
  const int r = blockIdx.x;  // g * K + j
  if (r >= G * K) return;
  const int g = r / K;
  const int id = idx[r];
  const float* a = inst_fts + (size_t)id * D;
  const float* b = view_fts + (size_t)g * D;
  const size_t o = (size_t)r * ldo;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    st16(out, o + c, a[c], kind);
    st16(out, o + D + c, b[c], kind);
  }
  for (int c = threadIdx.x; c < ldo - 2 * D; c += blockDim.x) {
    float v = 0.f;
    if (c < 3) v = centre[(size_t)g * 3 + c] - inst_pos[(size_t)id * 3 + c];
    st16(out, o + 2 * D + c, v, kind);
  }
}

// 6-d patch info rows for the policy's patch_position_embedding (POL:432-433):
// [rel_x, rel_y, rel_z, sin(direction), cos(direction), scale, 0, 0] from the [5, n, P] planes of d3d_patch_3d_info
__global__ void patch_info_rows_kernel(const float* __restrict__ info5, long long plane, long long n, void* __restrict__ out, int kind) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = info5[3 * plane + i];
  if (kind == D3D_OUT_F32) {
    float* o = (float*)out + i * 8;
    o[0] = info5[i]; o[1] = info5[plane + i]; o[2] = info5[2 * plane + i]; o[3] = sinf(d); o[4] = cosf(d); o[5] = info5[4 * plane + i];
    o[6] = 0.f; o[7] = 0.f;
    return;
  }
  uint4 p = make_uint4(pack16x2(info5[i], info5[plane + i], kind), pack16x2(info5[2 * plane + i], sinf(d), kind),
                       pack16x2(cosf(d), info5[4 * plane + i], kind), 0u);
  reinterpret_cast<uint4*>(out)[i] = p;
}

// out16[r] = [a[r] (D fp32) | b[r] (D fp32)] as 16-bit (instance/zone projector input, POL:434-435)
__global__ void concat2_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int D, void* __restrict__ out, int kind) {
  const int r = blockIdx.x;
  if (r >= n) return;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    st16(out, (size_t)r * 2 * D + c, a[(size_t)r * D + c], kind);
    st16(out, (size_t)r * 2 * D + D + c, b[(size_t)r * D + c], kind);
  }
}

// out16[r] = [x[r,0], x[r,1], x[r,2], 0...] (8 wide) -- 3-d relative positions as a GEMM A operand (POL:89-99)
__global__ void pos3_rows_kernel(const float* __restrict__ x, int n, void* __restrict__ out, int kind) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (kind == D3D_OUT_F32) {
    float* o = (float*)out + (size_t)r * 8;
    o[0] = x[(size_t)r * 3]; o[1] = x[(size_t)r * 3 + 1]; o[2] = x[(size_t)r * 3 + 2];
    o[3] = o[4] = o[5] = o[6] = o[7] = 0.f;
    return;
  }
  uint4 p = make_uint4(pack16x2(x[(size_t)r * 3], x[(size_t)r * 3 + 1], kind), pack16x2(x[(size_t)r * 3 + 2], 0.f, kind), 0u, 0u);
  reinterpret_cast<uint4*>(out)[r] = p;
}

}  // namespace

extern "C" int d3d_pool_features(const int64_t* seq_xyz, const int64_t* seq_dir, const int64_t* seq_scale, const float* centre,
                                 const int* tok_seq, const int* tok_src, int T, int mode, void* out16, int kind, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(seq_xyz && centre && tok_seq && tok_src && out16, "args");
  D3D_REQUIRE(mode == 1 || (seq_dir && seq_scale), "mode 0 needs direction and scale pools");
  pool_features_kernel<<<d3d_cdiv(T, 128), 128, 0, (cudaStream_t)stream>>>((const long long*)seq_xyz, (const long long*)seq_dir,
                                                                           (const long long*)seq_scale, centre, tok_seq, tok_src, T, mode,
                                                                           out16, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

// X16 != nullptr also writes the rows in the 16-bit operand type `kind` (in-library callers; the public entry below keeps its signature)
int d3d_pool_assemble_cast(const float* emb, const int64_t* seq_fts, int fts_is_f32, const int* tok_seq, const int* tok_src, const float* agg,
                           int T, int D, float* X, void* X16, int kind, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(emb && seq_fts && tok_seq && tok_src && agg && X, "args");
  D3D_REQUIRE(D % 4 == 0 && ((uintptr_t)emb % 16) == 0 && ((uintptr_t)X % 16) == 0 && ((uintptr_t)agg % 16) == 0, "width % 4, 16-byte rows");
  pool_assemble_kernel<<<T, 192, 0, (cudaStream_t)stream>>>(emb, (const long long*)seq_fts, fts_is_f32, tok_seq, tok_src, agg, T, D, X, X16, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_pool_assemble(const float* emb, const int64_t* seq_fts, int fts_is_f32, const int* tok_seq, const int* tok_src,
                                 const float* agg, int T, int D, float* X, void* stream) {
  return d3d_pool_assemble_cast(emb, seq_fts, fts_is_f32, tok_seq, tok_src, agg, T, D, X, nullptr, D3D_F16, stream);
}

extern "C" int d3d_disc_input(const float* inst_fts, const float* inst_pos, const int* idx, const float* view_fts, const float* centre,
                              int G, int K, int D, int ldo, void* out16, int kind, void* stream) {
  if (G * K == 0) return 0;
  D3D_REQUIRE(inst_fts && inst_pos && idx && view_fts && centre && out16, "args");
  D3D_REQUIRE(ldo >= 2 * D + 3, "row too narrow");
  disc_input_kernel<<<G * K, 256, 0, (cudaStream_t)stream>>>(inst_fts, inst_pos, idx, view_fts, centre, G, K, D, ldo, out16, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_patch_info_rows(const float* info5, int64_t n, void* out16, int kind, void* stream) {
  if (n == 0) return 0;
  D3D_REQUIRE(info5 && out16, "args");
  patch_info_rows_kernel<<<d3d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(info5, n, n, out16, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_concat2_cast(const float* a, const float* b, int n, int D, void* out16, int kind, void* stream) {
  if (n == 0) return 0;
  concat2_cast_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(a, b, n, D, out16, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_pos3_rows(const float* x, int n, void* out16, int kind, void* stream) {
  if (n == 0) return 0;
  pos3_rows_kernel<<<d3d_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(x, n, out16, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

// ================================================================================================
// batched (pointer-table) variants: one launch serves every episode of the rank
// ================================================================================================
namespace {

// copy n independent byte blocks (16-byte aligned, sizes multiple of 16): grid (chunks, n)
__global__ void copy_blocks_kernel(const long long* __restrict__ src, const long long* __restrict__ dst, const long long* __restrict__ nbytes) {
  const int b = blockIdx.y;
  const uint4* s = reinterpret_cast<const uint4*>(src[b]);
  uint4* d = reinterpret_cast<uint4*>(dst[b]);
  const long long n16 = nbytes[b] >> 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
}

// dst_row[r] (absolute row address) <- src[src_idx ? src_idx[r] : r], fp32 rows of width D
__global__ void scatter_rows_ptr_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ src_idx,
                                        const long long* __restrict__ dst_row, int n, int D) {
  const int r = blockIdx.x;
  if (r >= n) return;
  const long long s = src_idx ? src_idx[r] : r;
  float* d = reinterpret_cast<float*>(dst_row[r]);
  for (int c = threadIdx.x; c < D; c += blockDim.x) d[c] = src[s * lds + c];
}

__device__ __forceinline__ float dist2b(float qx, float qy, float qz, float rx, float ry, float rz) {
  const float dx = __fsub_rn(qx, rx), dy = __fsub_rn(qy, ry), dz = __fsub_rn(qz, rz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// K = 2 nearest per query, each query with its own reference set (its episode's instance slots); one warp per query.
// Same arithmetic and (d2, index) lexicographic order as knn_warp_kernel.  Missing neighbours (n_ref < 2): d2 = +inf, idx = -1.
__global__ void __launch_bounds__(128) knn2_batched_kernel(const long long* __restrict__ ref_ptr, const int* __restrict__ n_ref,
                                                           const float* __restrict__ qry, int n_q, float* __restrict__ out_d,
                                                           int* __restrict__ out_i) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= n_q) return;
  const float* refs = reinterpret_cast<const float*>(ref_ptr[q]);
  const int n = n_ref[q];
  const float qx = qry[(size_t)q * 3], qy = qry[(size_t)q * 3 + 1], qz = qry[(size_t)q * 3 + 2];
  float d0 = INFINITY, d1 = INFINITY;
  int i0 = 0x7fffffff, i1 = 0x7fffffff;
  for (int r = lane; r < n; r += 32) {
    const float d = dist2b(qx, qy, qz, refs[(size_t)r * 3], refs[(size_t)r * 3 + 1], refs[(size_t)r * 3 + 2]);
    if (d < d0 || (d == d0 && r < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = r; }
    else if (d < d1 || (d == d1 && r < i1)) { d1 = d; i1 = r; }
  }
  for (int j = 0; j < 2; ++j) {
    unsigned long long key = ((unsigned long long)__float_as_uint(d0) << 32) | (unsigned)i0;
    unsigned long long best = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    if (lane == 0) {
      const int bi = (int)(best & 0xffffffffu);
      out_d[(size_t)q * 2 + j] = __uint_as_float((unsigned)(best >> 32));
      out_i[(size_t)q * 2 + j] = bi == 0x7fffffff ? -1 : bi;
    }
    if (key == best) { d0 = d1; i0 = i1; d1 = INFINITY; i1 = 0x7fffffff; }
  }
}

// batched discriminator rows: row (q, j) for j < K from per-query instance pools; idx < 0 -> zero row
__global__ void disc_input_batched_kernel(const long long* __restrict__ fts_ptr, const long long* __restrict__ pos_ptr,
                                          const int* __restrict__ idx, const float* __restrict__ view_fts, const float* __restrict__ centre,
                                          int Q, int K, int D, int ldo, void* __restrict__ out, int kind) {
  const int r = blockIdx.x;
  if (r >= Q * K) return;
  const int q = r / K;
  const int id = idx[r];
  const size_t o = (size_t)r * ldo;
  if (id < 0) {
    for (int c = threadIdx.x; c < ldo; c += blockDim.x) st16(out, o + c, 0.f, kind);
    return;
  }
  const float* a = reinterpret_cast<const float*>(fts_ptr[q]) + (size_t)id * D;
  const float* p = reinterpret_cast<const float*>(pos_ptr[q]) + (size_t)id * 3;
  const float* b = view_fts + (size_t)q * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    st16(out, o + c, a[c], kind);
    st16(out, o + D + c, b[c], kind);
  }
  for (int c = threadIdx.x; c < ldo - 2 * D; c += blockDim.x) {
    float v = 0.f;
    if (c < 3) v = centre[(size_t)q * 3 + c] - p[c];
    st16(out, o + 2 * D + c, v, kind);
  }
}

}  // namespace

extern "C" int d3d_copy_blocks(const int64_t* src_ptr, const int64_t* dst_ptr, const int64_t* nbytes, int n, void* stream) {
  if (n == 0) return 0;
  D3D_REQUIRE(src_ptr && dst_ptr && nbytes, "args");
  copy_blocks_kernel<<<dim3(32, n), 256, 0, (cudaStream_t)stream>>>((const long long*)src_ptr, (const long long*)dst_ptr, (const long long*)nbytes);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_scatter_rows_ptr(const float* src, int64_t lds, const int* src_idx, const int64_t* dst_row_ptr, int n, int D, void* stream) {
  if (n == 0) return 0;
  D3D_REQUIRE(src && dst_row_ptr, "args");
  scatter_rows_ptr_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(src, lds, src_idx, (const long long*)dst_row_ptr, n, D);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_knn2_batched(const int64_t* ref_ptr, const int* n_ref, const float* queries, int n_q, float* out_d2, int* out_idx,
                                void* stream) {
  if (n_q == 0) return 0;
  D3D_REQUIRE(ref_ptr && n_ref && queries && out_d2 && out_idx, "args");
  knn2_batched_kernel<<<d3d_cdiv((long long)n_q * 32, 128), 128, 0, (cudaStream_t)stream>>>((const long long*)ref_ptr, n_ref, queries, n_q, out_d2,
                                                                                          out_idx);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_disc_input_batched(const int64_t* fts_ptr, const int64_t* pos_ptr, const int* idx, const float* view_fts,
                                      const float* centre, int Q, int K, int D, int ldo, void* out16, int kind, void* stream) {
  if (Q * K == 0) return 0;
  D3D_REQUIRE(fts_ptr && pos_ptr && idx && view_fts && centre && out16, "args");
  D3D_REQUIRE(ldo >= 2 * D + 3, "row too narrow");
  disc_input_batched_kernel<<<Q * K, 256, 0, (cudaStream_t)stream>>>((const long long*)fts_ptr, (const long long*)pos_ptr, idx, view_fts, centre, Q,
                                                                      K, D, ldo, out16, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}
