// Candidate-waypoint predictor (the reference's BinaryDistPredictor_TRM + heat-map NMS, POL:188-292):
// the kernels that are specific to it.  The linear layers run on d3d_gemm with split fp16x2/x3 operands (fp32-class accuracy: the
// arg-max of the heat map must not move), LayerNorm on d3d_layernorm.
//   wp_relu_kernel               : in-place ReLU on fp32 rows (nn.ReLU of TRM:27-31 / TRM:60-64)
//   wp_neighbor_attention_kernel : BERT self-attention over the 12 views of an episode with the additive neighbour mask of
//                                  utils.py:90-102 ((1 - mask) * -10000, WBERT:184-185): one warp per (episode, head, query)
//   wp_heatmap_nms_kernel        : POL:226-247 + utils.py:37-66: softmax over the 120 x 12 heat map, circular wrap by one angle row,
//                                  `max_predictions` rounds of {first arg-max, copy, box suppression}, un-wrap.  One CTA per episode.
#include <math.h>

#include "common.cuh"

namespace {

__global__ void wp_relu_kernel(float* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], 0.f);
}

// qkv: [n_ep * n_img, 3 * H * DH] fp32 rows = [q | k | v]; add_mask: [n_img, n_img] fp32 (0 or -10000); out: [n_ep * n_img, H * DH]
template <int DH, int MAXV>
__global__ void __launch_bounds__(128) wp_neighbor_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ add_mask, int n_ep,
                                                                    int n_img, int H, float scale, float* __restrict__ out) {
  constexpr int PL = DH / 32;  // channels per lane
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_ep * H * n_img) return;
  const int i = w % n_img, h = (w / n_img) % H, e = w / (n_img * H);
  const long long ld = 3LL * H * DH;
  const float* qp = qkv + (long long)(e * n_img + i) * ld + h * DH;
  float q[PL];
#pragma unroll
  for (int c = 0; c < PL; ++c) q[c] = qp[lane + 32 * c];
  float s[MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    s[j] = -INFINITY;
    if (j < n_img) {
      const float* kp = qkv + (long long)(e * n_img + j) * ld + (H + h) * DH;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < PL; ++c) d += q[c] * kp[lane + 32 * c];
      d = warp_sum(d);
      s[j] = d * scale + add_mask[i * n_img + j];  // scores / sqrt(d) + mask (WBERT:66-69)
      mx = fmaxf(mx, s[j]);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    s[j] = j < n_img ? expf(s[j] - mx) : 0.f;
    den += s[j];
  }
  float o[PL];
#pragma unroll
  for (int c = 0; c < PL; ++c) o[c] = 0.f;
#pragma unroll
  for (int j = 0; j < MAXV; ++j) {
    if (j < n_img) {
      const float p = s[j] / den;
      const float* vp = qkv + (long long)(e * n_img + j) * ld + (2 * H + h) * DH;
#pragma unroll
      for (int c = 0; c < PL; ++c) o[c] += p * vp[lane + 32 * c];
    }
  }
  float* op = out + (long long)(e * n_img + i) * H * DH + h * DH;
#pragma unroll
  for (int c = 0; c < PL; ++c) op[lane + 32 * c] = o[c];
}

// first maximum (lowest index among equals) of v[0..n) by a 256-thread CTA; result in *s_val / *s_idx
__device__ __forceinline__ void block_argmax_first(const float* v, int n, float* s_v, int* s_i) {
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float x = v[k];
    if (x > best || (x == best && k < bi)) { best = x; bi = k; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? s_v[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < nw ? s_i[threadIdx.x] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (threadIdx.x == 0) { s_v[0] = best; s_i[0] = bi; }
  }
  __syncthreads();
}

constexpr int WP_MAX_CELLS = 2048;  // (n_angles + 2) * n_classes of the reference: 122 * 12 = 1464

__global__ void __launch_bounds__(256) wp_heatmap_nms_kernel(const float* __restrict__ logits, int n_angles, int n_classes, int max_predictions,
                                                             float sigma_x, float sigma_y, float* __restrict__ prob, float* __restrict__ nms) {
  __shared__ float s_pred[WP_MAX_CELLS], s_supp[WP_MAX_CELLS], s_out[WP_MAX_CELLS];
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const int e = blockIdx.x, n = n_angles * n_classes;
  const float* lg = logits + (long long)e * n;
  // ---- softmax over the whole map (POL:226-230) ----
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < n; k += blockDim.x) mx = fmaxf(mx, lg[k]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_v[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_v[w]);
  __syncthreads();
  float sum = 0.f;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float ex = expf(lg[k] - mx);
    s_out[k] = ex;  // scratch
    sum += ex;
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += s_v[w];
  __syncthreads();
  // ---- wrap by one angle row on both sides (POL:234-238): rows [last, 0 .. n_angles-1, first] ----
  const int rows = n_angles + 2, cells = rows * n_classes;
  for (int k = threadIdx.x; k < cells; k += blockDim.x) {
    const int r = k / n_classes, c = k - r * n_classes;
    const int src = r == 0 ? n_angles - 1 : (r == rows - 1 ? 0 : r - 1);
    const float p = s_out[src * n_classes + c] / sum;
    s_pred[k] = p;
    s_supp[k] = p;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x)
    if (prob) prob[(long long)e * n + k] = s_pred[n_classes + k];
  for (int k = threadIdx.x; k < cells; k += blockDim.x) s_out[k] = 0.f;
  __syncthreads();
  // ---- utils.py:37-66: max_predictions rounds of arg-max + box suppression ----
  for (int it = 0; it < max_predictions; ++it) {
    block_argmax_first(s_supp, cells, s_v, s_i);
    const int ix = s_i[0];
    if (threadIdx.x == 0) s_out[ix] = s_pred[ix];  // flat_output[ix] = flat_pred[ix]
    // y = ix / W is a TRUE division in the reference (utils.py:55): a float row coordinate, so the box is not centred on a row
    const float y_mu = (float)ix / (float)n_classes;
    const float x_mu = (float)(ix % n_classes);
    for (int k = threadIdx.x; k < cells; k += blockDim.x) {
      const int r = k / n_classes, c = k - r * n_classes;
      const float yd = (float)r - y_mu;
      const float xd0 = (float)c - x_mu;
      const float xd = fminf(fabsf(xd0), fabsf(xd0 + (float)n_classes));  // circular_x over the class axis (utils.py:23-24)
      const float g = (fabsf(xd) <= sigma_x && fabsf(yd) <= sigma_y) ? 1.f : 0.f;
      s_supp[k] *= (1.f - g);
    }
    __syncthreads();
  }
  // output[output < 0] = 0 (utils.py:65) and the un-wrap [:, 1:-1, :] (POL:246)
  for (int k = threadIdx.x; k < n; k += blockDim.x) nms[(long long)e * n + k] = fmaxf(s_out[n_classes + k], 0.f);
}

}  // namespace

extern "C" int d3d_wp_relu(float* x, int64_t n, void* stream) {
  if (n == 0) return 0;
  D3D_REQUIRE(x != nullptr, "args");
  wp_relu_kernel<<<d3d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_wp_neighbor_attention(const float* qkv, const float* add_mask, int n_episodes, int n_img, int H, int Dh, float scale, float* out,
                                         void* stream) {
  if (n_episodes == 0) return 0;
  D3D_REQUIRE(qkv && add_mask && out, "args");
  D3D_REQUIRE(Dh == 64 && n_img >= 1 && n_img <= 12, "head_dim 64, at most 12 views per episode");
  const long long warps = (long long)n_episodes * H * n_img;
  wp_neighbor_attention_kernel<64, 12><<<d3d_cdiv(warps * 32, 128), 128, 0, (cudaStream_t)stream>>>(qkv, add_mask, n_episodes, n_img, H, scale, out);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_wp_heatmap_nms(const float* logits, int n_episodes, int n_angles, int n_classes, int max_predictions, float sigma_x, float sigma_y,
                                  float* prob, float* nms, void* stream) {
  if (n_episodes == 0) return 0;
  D3D_REQUIRE(logits && nms, "args");
  D3D_REQUIRE(n_angles >= 1 && n_classes >= 1 && (n_angles + 2) * n_classes <= WP_MAX_CELLS && max_predictions >= 0, "heat-map size");
  wp_heatmap_nms_kernel<<<n_episodes, 256, 0, (cudaStream_t)stream>>>(logits, n_angles, n_classes, max_predictions, sigma_x, sigma_y, prob, nms);
  D3D_CHECK_LAUNCH();
  return 0;
}
