// Kernels of the Pretrain novel-view patch renderer (Dynam3D_Pretrain/src_3dff/models/feature_fields.py = PFF,
// render_view_3d_patch PFF:494-625): ray sample points, per-ray important-sample selection on the K-NN result, neighbour
// gather into GEMM operands, and the volume rendering of PFF:446-474.  The K-NN itself is d3d_knn3d, the MLPs run on the
// tcgen05 GEMM (tinycudann's CutlassMLP = bias-free fp16 layers with LeakyReLU(0.01)).
#include "common.cuh"

namespace {

// ray_xyz[(ray, s)] = fp32( (rel_x*cos - rel_y*sin) + cam_x , (rel_x*sin + rel_y*cos) + cam_y , rel_z + cam_z ) evaluated in fp64
// exactly like the numpy expression PFF:523-528 (rel_x = rel_y[s] * tan_x[ray], rel_z = rel_y[s] * tan_z[ray], all float64)
__global__ void ray_points_kernel(const double* __restrict__ rel_y, const float* __restrict__ tan_x, const float* __restrict__ tan_z, int n_rays,
                                  int n_samples, double cs, double sn, double cx, double cy, double cz, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * n_samples) return;
  const int ray = i / n_samples, s = i % n_samples;
  const double y = rel_y[s];
  const double x = __dmul_rn(y, (double)tan_x[ray]);
  const double z = __dmul_rn(y, (double)tan_z[ray]);
  out[(size_t)i * 3 + 0] = (float)__dadd_rn(__dsub_rn(__dmul_rn(x, cs), __dmul_rn(y, sn)), cx);
  out[(size_t)i * 3 + 1] = (float)__dadd_rn(__dadd_rn(__dmul_rn(x, sn), __dmul_rn(y, cs)), cy);
  out[(size_t)i * 3 + 2] = (float)__dadd_rn(z, cz);
}

// PFF:542-552: dist = sqrt(d2); neighbours at >= radius are dropped (idx = -1, dist = radius); per-sample density proxy
// 1 / sum_k dist; the n_top samples of largest density per ray (ties: lowest sample index first).  One block per ray.
__global__ void __launch_bounds__(256) ray_topk_kernel(const float* __restrict__ d2, int* __restrict__ idx, int n_samples, int K, float radius,
                                                       int n_top, int* __restrict__ topk) {
  extern __shared__ float dens[];  // [n_samples]
  const int ray = blockIdx.x;
  for (int s = threadIdx.x; s < n_samples; s += blockDim.x) {
    const size_t base = ((size_t)ray * n_samples + s) * K;
    float sum = 0.f;
    for (int k = 0; k < K; ++k) {
      float d = __fsqrt_rn(d2[base + k]);
      if (d >= radius) { d = radius; idx[base + k] = -1; }
      sum = __fadd_rn(sum, d);
    }
    dens[s] = __fdiv_rn(1.0f, sum);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // n_top (8) rounds of arg-max over 501 values: tiny
    for (int t = 0; t < n_top; ++t) {
      int best = 0;
      float bv = -1.f;
      for (int s = 0; s < n_samples; ++s)
        if (dens[s] > bv) { bv = dens[s]; best = s; }
      topk[ray * n_top + t] = best;
      dens[best] = -2.f;
    }
  }
}

// gather the selected sample points: out[(ray, t)] = ray_xyz[(ray, topk[ray, t])]
__global__ void gather_samples_kernel(const float* __restrict__ ray_xyz, const int* __restrict__ topk, int n_rays, int n_samples, int n_top,
                                      float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * n_top) return;
  const int ray = i / n_top;
  const size_t src = ((size_t)ray * n_samples + topk[i]) * 3;
  out[(size_t)i * 3] = ray_xyz[src]; out[(size_t)i * 3 + 1] = ray_xyz[src + 1]; out[(size_t)i * 3 + 2] = ray_xyz[src + 2];
}

// PFF:588-615: second K-NN post-processing + neighbour gather.  For sample point p (= ray*n_top + t) and neighbour k:
//   pos rows  [P*K, 8]  : [rot(-cam)(patch_xyz - sample_xyz) (3), sin(dir_patch - cam - dir_ray), cos(...), scale, 0, 0]; invalid -> [far,far,far,0,0,0]
//   feat rows [P, K*D]  : patch_fts[idx] (16-bit), zeros for invalid neighbours
__global__ void nerf_gather_kernel(const float* __restrict__ d2, int* __restrict__ idx, const float* __restrict__ sample_xyz,
                                   const float* __restrict__ patch_xyz, const float* __restrict__ patch_dir, const float* __restrict__ patch_scale,
                                   const __half* __restrict__ patch_fts, const float* __restrict__ ray_dir, int n_top, int K, int D, float radius,
                                   float far_, float cam_dir, float cs, float sn, void* __restrict__ pos_rows, int pos_kind,
                                   __half* __restrict__ feat_rows) {
  const int r = blockIdx.x;  // p * K + k
  const int p = r / K;
  const int ray = p / n_top;
  int id = idx[r];
  if (__fsqrt_rn(d2[r]) >= radius) id = -1;
  if (threadIdx.x == 0) {
    idx[r] = id;
    float f[8] = {far_, far_, far_, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (id >= 0) {
      const float x = patch_xyz[(size_t)id * 3] - sample_xyz[(size_t)p * 3];
      const float y = patch_xyz[(size_t)id * 3 + 1] - sample_xyz[(size_t)p * 3 + 1];
      f[2] = patch_xyz[(size_t)id * 3 + 2] - sample_xyz[(size_t)p * 3 + 2];
      f[0] = x * cs - y * sn;   // cs = cos(-cam), sn = sin(-cam)
      f[1] = f[0] * sn + y * cs;  // literal PFF:596-599: the reference's `x` is a VIEW that already holds the rotated value
      const float a = (patch_dir[id] - cam_dir) - ray_dir[ray];
      f[3] = sinf(a); f[4] = cosf(a);
      f[5] = patch_scale[id];
    }
    for (int i = 0; i < 8; ++i) st16(pos_rows, (size_t)r * 8 + i, f[i], pos_kind);
  }
  __half* dst = feat_rows + (size_t)r * D;
  if (id >= 0) {
    const uint4* src = reinterpret_cast<const uint4*>(patch_fts + (size_t)id * D);
    for (int i = threadIdx.x; i < D / 8; i += blockDim.x) reinterpret_cast<uint4*>(dst)[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < D / 8; i += blockDim.x) reinterpret_cast<uint4*>(dst)[i] = make_uint4(0, 0, 0, 0);
  }
}

// out16[i] = fp16( fp16(a32[i]) + b16[i] )   (PFF:478-482: both terms are cast to fp16 before the add)
__global__ void add_half_kernel(const float* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __hadd(__float2half_rn(a[i]), b[i]);
}

// raw2feature (PFF:446-474), one block per ray: softplus density at the n_top selected samples, alpha compositing over all samples
__global__ void __launch_bounds__(256) volume_render_kernel(const float* __restrict__ feat, const float* __restrict__ density,
                                                            const int* __restrict__ topk, const float* __restrict__ rel_dist, int n_samples,
                                                            int n_top, int D, float* __restrict__ feature_map, float* __restrict__ depth_map) {
  extern __shared__ float sm[];   // dens[n_samples] | w[n_samples]
  float* dens = sm;
  float* w = sm + n_samples;
  __shared__ float sw[32];
  __shared__ float red[256];
  const int ray = blockIdx.x;
  for (int s = threadIdx.x; s < n_samples; s += blockDim.x) dens[s] = 0.f;
  __syncthreads();
  if (threadIdx.x < n_top) {
    const float x = density[ray * n_top + threadIdx.x];
    dens[topk[ray * n_top + threadIdx.x]] = x > 20.f ? x : log1pf(expf(x));  // F.softplus (beta 1, threshold 20)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float T = 1.f, wsum = 0.f, dsum = 0.f;
    for (int s = 0; s < n_samples; ++s) {
      const float dist = s + 1 < n_samples ? fabsf(rel_dist[s + 1] - rel_dist[s]) : 1e10f;
      const float alpha = 1.f - expf(-fmaxf(dens[s], 0.f) * dist);
      w[s] = alpha * T;
      T *= (1.f - alpha + 1e-10f);
      wsum += w[s];
      dsum += w[s] * rel_dist[s];
    }
    depth_map[ray] = dsum / fmaxf(wsum, 1e-7f);
    for (int t = 0; t < n_top; ++t) sw[t] = w[topk[ray * n_top + t]];
  }
  __syncthreads();
  float sq = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < n_top; ++t) acc += sw[t] * feat[((size_t)ray * n_top + t) * D + c];
    feature_map[(size_t)ray * D + c] = acc;
    sq += acc * acc;
  }
  red[threadIdx.x] = sq;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float inv = 1.f / fmaxf(sqrtf(red[0]), 1e-7f);
  for (int c = threadIdx.x; c < D; c += blockDim.x) feature_map[(size_t)ray * D + c] *= inv;
}

}  // namespace

extern "C" int d3d_ray_points_habitat(const double* rel_y, const float* tan_x, const float* tan_z, int n_rays, int n_samples, double cos_h,
                                      double sin_h, double cam_x, double cam_y, double cam_z, float* out_xyz, void* stream) {
  D3D_REQUIRE(rel_y && tan_x && tan_z && out_xyz && n_rays > 0 && n_samples > 0, "args");
  ray_points_kernel<<<d3d_cdiv((long long)n_rays * n_samples, 256), 256, 0, (cudaStream_t)stream>>>(rel_y, tan_x, tan_z, n_rays, n_samples, cos_h,
                                                                                                    sin_h, cam_x, cam_y, cam_z, out_xyz);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_ray_topk(const float* d2, int* idx, int n_rays, int n_samples, int K, float radius, int n_top, int* topk, void* stream) {
  D3D_REQUIRE(d2 && idx && topk && n_top <= n_samples && n_samples * 4 <= 48 * 1024, "args");
  ray_topk_kernel<<<n_rays, 256, n_samples * sizeof(float), (cudaStream_t)stream>>>(d2, idx, n_samples, K, radius, n_top, topk);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_gather_samples(const float* ray_xyz, const int* topk, int n_rays, int n_samples, int n_top, float* out, void* stream) {
  D3D_REQUIRE(ray_xyz && topk && out, "args");
  gather_samples_kernel<<<d3d_cdiv((long long)n_rays * n_top, 128), 128, 0, (cudaStream_t)stream>>>(ray_xyz, topk, n_rays, n_samples, n_top, out);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_nerf_gather(const float* d2, int* idx, const float* sample_xyz, const float* patch_xyz, const float* patch_dir,
                               const float* patch_scale, const void* patch_fts16, const float* ray_dir, int n_points, int n_top, int K, int D,
                               float radius, float far_, float cam_dir, float cos_neg, float sin_neg, void* pos_rows, int pos_kind,
                               void* feat_rows16, void* stream) {
  D3D_REQUIRE(d2 && idx && sample_xyz && patch_xyz && patch_dir && patch_scale && patch_fts16 && ray_dir && pos_rows && feat_rows16, "args");
  D3D_REQUIRE(D % 8 == 0, "feature width");
  nerf_gather_kernel<<<n_points * K, 128, 0, (cudaStream_t)stream>>>(d2, idx, sample_xyz, patch_xyz, patch_dir, patch_scale,
                                                                      (const __half*)patch_fts16, ray_dir, n_top, K, D, radius, far_, cam_dir,
                                                                      cos_neg, sin_neg, pos_rows, pos_kind, (__half*)feat_rows16);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_add_half(const float* a32, const void* b16, void* out16, int64_t n, void* stream) {
  if (n == 0) return 0;
  add_half_kernel<<<d3d_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(a32, (const __half*)b16, (__half*)out16, n);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_volume_render(const float* feat, const float* density, const int* topk, const float* rel_dist, int n_rays, int n_samples,
                                 int n_top, int D, float* feature_map, float* depth_map, void* stream) {
  D3D_REQUIRE(feat && density && topk && rel_dist && feature_map && depth_map && n_top <= 32, "args");
  volume_render_kernel<<<n_rays, 256, 2 * n_samples * sizeof(float), (cudaStream_t)stream>>>(feat, density, topk, rel_dist, n_samples, n_top, D,
                                                                                              feature_map, depth_map);
  D3D_CHECK_LAUNCH();
  return 0;
}
