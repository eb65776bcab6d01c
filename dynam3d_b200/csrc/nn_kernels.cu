// Normalisation / elementwise / gather kernels around the tensor-core GEMMs (all fp32 math, 16-bit stores where the
// consumer is a GEMM A operand).  One warp per row everywhere: rows are 768..4096 wide, so a row lives in registers.
#include "common.cuh"

namespace {

constexpr int MAX_D = 4096;

// ------------------------------------------------------------------------------------------------
// LayerNorm (fp32 statistics, two-pass), optional row gather, optional GELU, fp32 and/or 16-bit output
//   y[r] = act( LN(x[idx ? idx[r] : r]) * gamma + beta )
// torch.nn.LayerNorm semantics (biased variance), CLIPM:153-159, FF:141,146,151,155,159
// ------------------------------------------------------------------------------------------------
template <int VPL>  // float4 vectors per lane: D = 128 * VPL
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long ldx, const int* __restrict__ row_index,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int T,
                                                        int D, int act, float* __restrict__ out32, long long ld32,
                                                        void* __restrict__ out16, long long ld16, int kind16) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  const long long src = row_index ? (long long)row_index[row] : (long long)row;
  const float4* xr = reinterpret_cast<const float4*>(x + src * ldx);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = reinterpret_cast<const float4*>(gamma)[c4];
    const float4 b = reinterpret_cast<const float4*>(beta)[c4];
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (act == D3D_ACT_GELU) { y.x = gelu_erf(y.x); y.y = gelu_erf(y.y); y.z = gelu_erf(y.z); y.w = gelu_erf(y.w); }
    if (out32) reinterpret_cast<float4*>(out32 + (long long)row * ld32)[c4] = y;
    if (out16) {
      uint2 p = make_uint2(pack16x2(y.x, y.y, kind16), pack16x2(y.z, y.w, kind16));
      reinterpret_cast<uint2*>((uint16_t*)out16 + (long long)row * ld16)[c4] = p;
    }
  }
}

// RMSNorm (HF LlamaRMSNorm / Phi3RMSNorm): y = w * x * rsqrt(mean(x^2) + eps)
template <int VPL>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const float* __restrict__ x, long long ldx, const int* __restrict__ row_index,
                                                      const float* __restrict__ w, float eps, int T, int D, float* __restrict__ out32,
                                                      long long ld32, void* __restrict__ out16, long long ld16, int kind16) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  const long long src = row_index ? (long long)row_index[row] : (long long)row;
  const float4* xr = reinterpret_cast<const float4*>(x + src * ldx);
  float4 v[VPL];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[lane + 32 * i];
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float r = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = reinterpret_cast<const float4*>(w)[c4];
    float4 y = make_float4(v[i].x * r * g.x, v[i].y * r * g.y, v[i].z * r * g.z, v[i].w * r * g.w);
    if (out32) reinterpret_cast<float4*>(out32 + (long long)row * ld32)[c4] = y;
    if (out16) {
      uint2 p = make_uint2(pack16x2(y.x, y.y, kind16), pack16x2(y.z, y.w, kind16));
      reinterpret_cast<uint2*>((uint16_t*)out16 + (long long)row * ld16)[c4] = p;
    }
  }
}

template <typename F>
int dispatch_vpl(int D, F&& f) {
  switch (D / 128) {
    case 6: return f(std::integral_constant<int, 6>());    // 768
    case 8: return f(std::integral_constant<int, 8>());    // 1024
    case 24: return f(std::integral_constant<int, 24>());  // 3072
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 4: return f(std::integral_constant<int, 4>());
    case 32: return f(std::integral_constant<int, 32>());  // 4096
    default:
      d3d_set_error("row width %d not supported (need 128, 256, 512, 768, 1024, 3072 or 4096)", D);
      return D3D_EINVAL;
  }
}

// ------------------------------------------------------------------------------------------------
// rotary embedding (HF rotate_half form) applied in place to q and k of a packed [T, 3*H*Dh] 16-bit QKV buffer
// ------------------------------------------------------------------------------------------------
__global__ void rope_kernel(void* __restrict__ qkv, long long ld, const int* __restrict__ pos, const float* __restrict__ inv_freq, int T,
                            int H, int Dh, int kind) {
  const int half = Dh / 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)T * 2 * H * half;
  if (idx >= total) return;
  const int i = (int)(idx % half);
  const int h = (int)((idx / half) % (2 * H));  // q heads then k heads
  const int t = (int)(idx / ((long long)half * 2 * H));
  const float ang = (float)pos[t] * inv_freq[i];
  float sn, cs;
  sincosf(ang, &sn, &cs);
  const size_t base = (size_t)t * ld + (size_t)h * Dh;
  const float a = ld16(qkv, base + i, kind), b = ld16(qkv, base + i + half, kind);
  // products rounded separately (no FMA contraction), as in x*cos + rotate_half(x)*sin evaluated by PyTorch
  st16(qkv, base + i, __fsub_rn(__fmul_rn(a, cs), __fmul_rn(b, sn)), kind);
  st16(qkv, base + i + half, __fadd_rn(__fmul_rn(b, cs), __fmul_rn(a, sn)), kind);
}

// The same rotation from a per-token cos/sin table (built once per prefill, reused by all 32 layers), 8 elements per thread:
// tab [T, Dh] fp32 = [cos(pos*inv_freq[0..half)) | sin(...)]; 16-bit QKV only.  Same sincosf on the same angle as rope_kernel.
__global__ void rope_table_kernel(const int* __restrict__ pos, const float* __restrict__ inv_freq, int T, int half, float* __restrict__ tab) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * half) return;
  const int t = idx / half, i = idx - t * half;
  float sn, cs;
  sincosf((float)pos[t] * inv_freq[i], &sn, &cs);
  tab[(size_t)t * 2 * half + i] = cs;
  tab[(size_t)t * 2 * half + half + i] = sn;
}

__global__ void rope_apply_kernel(uint16_t* __restrict__ qkv, long long ld, const float* __restrict__ tab, int T, int H, int Dh, int kind) {
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

// ------------------------------------------------------------------------------------------------
// embedding gather: out32[t] = table16[ids[t]]
// ------------------------------------------------------------------------------------------------
__global__ void embed_gather_kernel(const void* __restrict__ table, int kind, const int* __restrict__ ids, int T, int D,
                                    float* __restrict__ out, long long ldo) {
  const int row = blockIdx.x;
  if (row >= T) return;
  const size_t src = (size_t)ids[row] * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) out[(long long)row * ldo + c] = ld16(table, src + c, kind);
}

// ------------------------------------------------------------------------------------------------
// a1: image preprocessing + im2col for the 14x14/14 patch embedding (ENC:267-284, CLIPM:220-222)
//   u8 NHWC -> (bicubic resize to RxR, A=-0.75, align_corners=False, round+clamp to u8 as torchvision does) -> /255
//   -> (x-mean)/std -> 16-bit -> patches [N*g*g, Kpad], column = c*p*p + ky*p + kx
// ------------------------------------------------------------------------------------------------
// The expressions below are written exactly like PyTorch's CUDA bicubic kernel (ATen/native/cuda/UpSample.cuh: cubic_convolution1/2,
// get_cubic_upsample_coefficients, cubic_interp1d; UpSampleBicubic2d.cu: x pass per row, then the y pass) so that nvcc contracts them into
// the same FMA sequence: resized pixels are bit-identical to torch's CUDA `interpolate(mode="bicubic")` -- the kernel the reference's
// torchvision Resize (ENC:268) runs for observations that live on the GPU (tests/test_nn_kernels_gpu.py).  PyTorch's CPU kernel evaluates
// the same formula in a different order; after rounding to uint8 it differs on < 1e-4 of the pixels, from this kernel as from torch's own.
__device__ __forceinline__ float cubic_convolution1(float x, float A) { return ((A + 2) * x - (A + 3)) * x * x + 1; }
__device__ __forceinline__ float cubic_convolution2(float x, float A) { return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float* coeffs) {
  const float A = -0.75f;
  float x1 = t;
  coeffs[0] = cubic_convolution2(x1 + 1.0f, A);
  coeffs[1] = cubic_convolution1(x1, A);
  float x2 = 1.0f - t;
  coeffs[2] = cubic_convolution1(x2, A);
  coeffs[3] = cubic_convolution2(x2 + 1.0f, A);
}
__device__ __forceinline__ float cubic_interp1d(float x0, float x1, float x2, float x3, const float* coeffs) {
  return x0 * coeffs[0] + x1 * coeffs[1] + x2 * coeffs[2] + x3 * coeffs[3];
}

struct NormParams {
  float mean[3];
  float std[3];
};

__global__ void preprocess_im2col_kernel(const uint8_t* __restrict__ img, int N, int Hin, int Win, int R, int p, NormParams np,
                                         void* __restrict__ out, int kpad, int kind) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * R * R;
  if (idx >= total) return;
  const int ox = (int)(idx % R), oy = (int)((idx / R) % R), n = (int)(idx / ((long long)R * R));
  const uint8_t* src = img + (size_t)n * Hin * Win * 3;
  float px[3];
  if (Hin == R && Win == R) {
    for (int c = 0; c < 3; ++c) px[c] = (float)src[((size_t)oy * Win + ox) * 3 + c];
  } else {
    const float sy = (float)Hin / (float)R, sx = (float)Win / (float)R;
    const float fy = sy * ((float)oy + 0.5f) - 0.5f, fx = sx * ((float)ox + 0.5f) - 0.5f;
    const int iy = (int)floorf(fy), ix = (int)floorf(fx);
    float wy[4], wx[4];
    cubic_coeffs(fy - (float)iy, wy);
    cubic_coeffs(fx - (float)ix, wx);
    int xs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = min(max(ix - 1 + j, 0), Win - 1);
    for (int c = 0; c < 3; ++c) {
      float rows[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint8_t* r = src + (size_t)min(max(iy - 1 + i, 0), Hin - 1) * Win * 3 + c;
        rows[i] = cubic_interp1d((float)r[xs[0] * 3], (float)r[xs[1] * 3], (float)r[xs[2] * 3], (float)r[xs[3] * 3], wx);
      }
      const float acc = cubic_interp1d(rows[0], rows[1], rows[2], rows[3], wy);
      px[c] = fminf(fmaxf(rintf(acc), 0.0f), 255.0f);  // torchvision casts the resized float image back to uint8
    }
  }
  const int g = R / p;
  const int prow = n * g * g + (oy / p) * g + (ox / p);
  const int ky = oy % p, kx = ox % p;
  for (int c = 0; c < 3; ++c) {
    const float v = (px[c] / 255.0f - np.mean[c]) / np.std[c];
    st16(out, (size_t)prow * kpad + c * p * p + ky * p + kx, v, kind);
  }
}

// ------------------------------------------------------------------------------------------------
// a15 input path (POL:438): HF CLIPImageProcessor resizes with Pillow -- ImagingResample, 8 bits per channel: fixed-point taps (22 bits),
// horizontal pass -> uint8 intermediate -> vertical pass, clip8((acc + 2^21) >> 22).  One pass along `axis`; taps / bounds are tables
// built on the host (ops.pil_bicubic_tables, Pillow's precompute_coeffs + normalize_coeffs_8bpc).  Integer work: bit-exact.
//   src [N, H, W, C] u8 -> dst [N, H, Wout, C] (axis = 1, along W) or [N, Hout, W, C] (axis = 0, along H)
// ------------------------------------------------------------------------------------------------
__global__ void pil_resample_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int N, int H, int W, int C, int out_size,
                                         int axis, const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const int Ho = axis == 0 ? out_size : H, Wo = axis == 1 ? out_size : W;
  const long long total = (long long)N * Ho * Wo * C;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int x = (int)((idx / C) % Wo), y = (int)((idx / ((long long)C * Wo)) % Ho), n = (int)(idx / ((long long)C * Wo * Ho));
  const int o = axis == 1 ? x : y;
  const int first = bounds[o * 2], cnt = bounds[o * 2 + 1];
  const int* k = kk + (size_t)o * ksize;
  const uint8_t* base = src + (size_t)n * H * W * C + c;
  int acc = 1 << 21;
  if (axis == 1) {
    const uint8_t* r = base + (size_t)y * W * C;
    for (int t = 0; t < cnt; ++t) acc += (int)r[(size_t)(first + t) * C] * k[t];
  } else {
    const uint8_t* r = base + (size_t)x * C;
    for (int t = 0; t < cnt; ++t) acc += (int)r[(size_t)(first + t) * W * C] * k[t];
  }
  dst[idx] = (uint8_t)min(max(acc >> 22, 0), 255);
}

// zero the K padding columns [k, kpad) of an im2col matrix
__global__ void zero_pad_cols_kernel(void* out, long long rows, int k, int kpad, int kind) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = kpad - k;
  if (idx >= rows * w) return;
  st16(out, (size_t)((idx / w) * kpad + k + (idx % w)), 0.f, kind);
}

// precise-mode operand split: out[r, c] = hi = fp16(x), out[r, K + c] = lo = fp16(x - hi), and with terms == 3 also
// out[r, 2K + c] = hi (for weights that are not fp16-representable: A.[W_hi|W_hi|W_lo] = A_hi W_hi + A_lo W_hi + A_hi W_lo)
__global__ void split16_kernel(const float* __restrict__ in, long long ldi, __half* __restrict__ out, long long ldo, int T, int K, int terms) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)T * K) return;
  const long long r = i / K, c = i % K;
  const float x = in[r * ldi + c];
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  __half* o = out + r * ldo;
  o[c] = hi;
  o[K + c] = lo;
  if (terms == 3) o[2 * (long long)K + c] = hi;
}

// ViT token assembly + ln_pre (CLIPM:223-225): X[n, 0] = cls + pos[0]; X[n, 1+i] = conv[n*g2+i] + pos[1+i]; X = ln_pre(X)
template <int VPL>
__global__ void __launch_bounds__(256) vit_embed_ln_kernel(const float* __restrict__ conv, const float* __restrict__ cls,
                                                           const float* __restrict__ pos, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int N, int tokens, int D,
                                                           float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N * tokens) return;
  const int n = row / tokens, t = row % tokens;
  const float4* a = t == 0 ? reinterpret_cast<const float4*>(cls)
                           : reinterpret_cast<const float4*>(conv + ((size_t)n * (tokens - 1) + (t - 1)) * D);
  const float4* pp = reinterpret_cast<const float4*>(pos + (size_t)t * D);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 x = a[lane + 32 * i], q = pp[lane + 32 * i];
    v[i] = make_float4(x.x + q.x, x.y + q.y, x.z + q.z, x.w + q.w);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)D;
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a0 = v[i].x - mean, b0 = v[i].y - mean, c0 = v[i].z - mean, d0 = v[i].w - mean;
    q2 += (a0 * a0 + b0 * b0) + (c0 * c0 + d0 * d0);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q2) / (float)D + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = reinterpret_cast<const float4*>(gamma)[c4];
    const float4 b = reinterpret_cast<const float4*>(beta)[c4];
    float4 y = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                           (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    reinterpret_cast<float4*>(out + (size_t)row * D)[c4] = y;
  }
}

// generic row ops -------------------------------------------------------------------------------
// dst[dst_idx[r]] = src[src_idx[r]] for fp32 rows of width D (instance / zone slot writes, FF:644-648, 730, 756)
__global__ void scatter_rows_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ src_idx, float* __restrict__ dst,
                                    long long ldd, const int* __restrict__ dst_idx, int n, int D) {
  const int r = blockIdx.x;
  if (r >= n) return;
  const long long s = src_idx ? src_idx[r] : r, d = dst_idx ? dst_idx[r] : r;
  for (int c = threadIdx.x; c < D; c += blockDim.x) dst[d * ldd + c] = src[s * lds + c];
}

// out[r] = a[r] + b[r % period] (fp32), e.g. patch_features + patch_position_fts (POL:453)
__global__ void add_rows_kernel(float* __restrict__ a, const float* __restrict__ b, long long n, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * D) a[i] += b[i];
}

__global__ void cast_rows_kernel(const float* __restrict__ in, long long ldi, void* __restrict__ out, long long ldo, int T, int D, int kind) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)T * D) return;
  const long long r = i / D, c = i % D;
  st16(out, (size_t)(r * ldo + c), in[r * ldi + c], kind);
}

}  // namespace

extern "C" int d3d_layernorm(const float* x, int64_t ldx, const int* row_index, const float* gamma, const float* beta, float eps, int T,
                             int D, int act, float* out32, int64_t ld32, void* out16, int64_t ld16, int kind16, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(x && gamma && beta && (out32 || out16), "args");
  D3D_REQUIRE(D % 128 == 0 && D <= MAX_D && ldx % 4 == 0, "row width");
  D3D_REQUIRE(act == D3D_ACT_NONE || act == D3D_ACT_GELU, "act");
  const int grid = d3d_cdiv((long long)T * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dispatch_vpl(D, [&](auto vpl) {
    layernorm_kernel<decltype(vpl)::value><<<grid, 256, 0, st>>>(x, ldx, row_index, gamma, beta, eps, T, D, act, out32, ld32, out16, ld16, kind16);
    return 0;
  });
  D3D_TRY(rc);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_rmsnorm(const float* x, int64_t ldx, const int* row_index, const float* w, float eps, int T, int D, float* out32,
                           int64_t ld32, void* out16, int64_t ld16, int kind16, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(x && w && (out32 || out16), "args");
  D3D_REQUIRE(D % 128 == 0 && D <= MAX_D && ldx % 4 == 0, "row width");
  const int grid = d3d_cdiv((long long)T * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dispatch_vpl(D, [&](auto vpl) {
    rmsnorm_kernel<decltype(vpl)::value><<<grid, 256, 0, st>>>(x, ldx, row_index, w, eps, T, D, out32, ld32, out16, ld16, kind16);
    return 0;
  });
  D3D_TRY(rc);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_rope(void* qkv, int64_t ld, const int* pos, const float* inv_freq, int T, int H, int Dh, int kind, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(qkv && pos && inv_freq && Dh % 2 == 0, "args");
  const long long total = (long long)T * 2 * H * (Dh / 2);
  rope_kernel<<<d3d_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(qkv, ld, pos, inv_freq, T, H, Dh, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_rope_table(const int* pos, const float* inv_freq, int T, int Dh, float* tab, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(pos && inv_freq && tab && Dh % 2 == 0, "args");
  rope_table_kernel<<<d3d_cdiv((long long)T * (Dh / 2), 256), 256, 0, (cudaStream_t)stream>>>(pos, inv_freq, T, Dh / 2, tab);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_rope_apply(void* qkv, int64_t ld, const float* tab, int T, int H, int Dh, int kind, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(qkv && tab, "args");
  D3D_REQUIRE(kind == D3D_F16 || kind == D3D_BF16, "16-bit QKV only (fp32 buffers use d3d_rope)");
  D3D_REQUIRE(Dh % 16 == 0 && ld % 8 == 0 && ((uintptr_t)qkv % 16) == 0 && ((uintptr_t)tab % 16) == 0, "16-byte vectors");
  const long long total = (long long)T * 2 * H * (Dh / 16);
  rope_apply_kernel<<<d3d_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((uint16_t*)qkv, ld, tab, T, H, Dh, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_embed_gather(const void* table, int kind, const int* ids, int T, int D, float* out, int64_t ldo, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(table && ids && out, "args");
  embed_gather_kernel<<<T, 256, 0, (cudaStream_t)stream>>>(table, kind, ids, T, D, out, ldo);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_preprocess_im2col(const uint8_t* img, int N, int Hin, int Win, int R, int patch, const float* mean3_h,
                                     const float* std3_h, void* out, int kpad, int kind, void* stream) {
  D3D_REQUIRE(img && out && mean3_h && std3_h && N > 0, "args");
  D3D_REQUIRE(R % patch == 0 && kpad >= 3 * patch * patch && kpad % 8 == 0, "patch geometry");
  NormParams np;
  for (int c = 0; c < 3; ++c) { np.mean[c] = mean3_h[c]; np.std[c] = std3_h[c]; }
  cudaStream_t st = (cudaStream_t)stream;
  const int g = R / patch;
  const long long rows = (long long)N * g * g;
  const int k = 3 * patch * patch;
  if (kpad > k) {
    zero_pad_cols_kernel<<<d3d_cdiv(rows * (kpad - k), 256), 256, 0, st>>>(out, rows, k, kpad, kind);
    D3D_CHECK_LAUNCH();
  }
  const long long total = (long long)N * R * R;
  preprocess_im2col_kernel<<<d3d_cdiv(total, 256), 256, 0, st>>>(img, N, Hin, Win, R, patch, np, out, kpad, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_pil_resample_pass(const uint8_t* src, uint8_t* dst, int N, int H, int W, int C, int out_size, int axis, const int* bounds,
                                     const int* kk, int ksize, void* stream) {
  D3D_REQUIRE(src && dst && bounds && kk && N > 0 && (axis == 0 || axis == 1) && ksize > 0, "args");
  const long long total = (long long)N * (axis == 0 ? out_size : H) * (axis == 1 ? out_size : W) * C;
  pil_resample_pass_kernel<<<d3d_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, N, H, W, C, out_size, axis, bounds, kk, ksize);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_vit_embed_ln(const float* conv, const float* cls, const float* pos, const float* gamma, const float* beta, float eps,
                                int N, int tokens, int D, float* out, void* stream) {
  D3D_REQUIRE(conv && cls && pos && gamma && beta && out, "args");
  D3D_REQUIRE(D % 128 == 0 && D <= MAX_D, "row width");
  const int grid = d3d_cdiv((long long)N * tokens * 32, 256);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dispatch_vpl(D, [&](auto vpl) {
    vit_embed_ln_kernel<decltype(vpl)::value><<<grid, 256, 0, st>>>(conv, cls, pos, gamma, beta, eps, N, tokens, D, out);
    return 0;
  });
  D3D_TRY(rc);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_scatter_rows(const float* src, int64_t lds, const int* src_idx, float* dst, int64_t ldd, const int* dst_idx, int n, int D,
                                void* stream) {
  if (n == 0) return 0;
  D3D_REQUIRE(src && dst, "args");
  scatter_rows_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(src, lds, src_idx, dst, ldd, dst_idx, n, D);
  D3D_CHECK_LAUNCH();
  return 0;
}

namespace {
// dst[dst_idx ? dst_idx[r] : r] = src[src_idx ? src_idx[r] : r] for 16-bit rows of D elements (D % 8 == 0: 16-byte vectors)
__global__ void copy_rows16_kernel(const uint4* __restrict__ src, long long lds8, const int* __restrict__ src_idx, uint4* __restrict__ dst,
                                   long long ldd8, const int* __restrict__ dst_idx, int n, int D8) {
  const int r = blockIdx.x;
  if (r >= n) return;
  const uint4* s = src + (size_t)(src_idx ? src_idx[r] : r) * lds8;
  uint4* d = dst + (size_t)(dst_idx ? dst_idx[r] : r) * ldd8;
  for (int c = threadIdx.x; c < D8; c += blockDim.x) d[c] = s[c];
}
}  // namespace

extern "C" int d3d_gather_rows16(const void* src, int64_t lds, const int* idx, void* dst, int64_t ldd, int n, int D, void* stream) {
  if (n == 0) return 0;
  D3D_REQUIRE(src && idx && dst, "args");
  D3D_REQUIRE(D % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0, "16-byte aligned rows");
  copy_rows16_kernel<<<n, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, lds / 8, idx, (uint4*)dst, ldd / 8, nullptr, n, D / 8);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_scatter_rows16(const void* src, int64_t lds, void* dst, int64_t ldd, const int* dst_idx, int n, int D, void* stream) {
  if (n == 0) return 0;
  D3D_REQUIRE(src && dst_idx && dst, "args");
  D3D_REQUIRE(D % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0, "16-byte aligned rows");
  copy_rows16_kernel<<<n, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, lds / 8, nullptr, (uint4*)dst, ldd / 8, dst_idx, n, D / 8);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_add_inplace(float* a, const float* b, int64_t n_rows, int D, void* stream) {
  if (n_rows == 0) return 0;
  add_rows_kernel<<<d3d_cdiv(n_rows * D, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n_rows, D);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_cast16(const float* in, int64_t ldi, void* out, int64_t ldo, int T, int D, int kind, void* stream) {
  if (T == 0) return 0;
  cast_rows_kernel<<<d3d_cdiv((long long)T * D, 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, out, ldo, T, D, kind);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_split16(const float* in, int64_t ldi, void* out16, int64_t ldo, int T, int K, int terms, void* stream) {
  if (T == 0) return 0;
  D3D_REQUIRE(in && out16 && (terms == 2 || terms == 3) && ldo >= (int64_t)terms * K, "args");
  split16_kernel<<<d3d_cdiv((long long)T * K, 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, (__half*)out16, ldo, T, K, terms);
  D3D_CHECK_LAUNCH();
  return 0;
}
