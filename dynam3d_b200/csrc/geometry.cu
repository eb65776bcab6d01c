// HBM-bound fp32 geometry kernels of the 3D-token path: depth preprocessing, habitat unprojection,
// frustum culling, exact 3-D K-NN (replaces torch_kdtree), centroids, agent-frame token export.
//
// Bit-exactness contract (oracle/geometry.py): every fp32 operation is an individually rounded IEEE op in the
// reference's order -- hence the explicit __f*_rn intrinsics (nvcc would otherwise contract a*b+c into FMA).
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// a3: preprocess_depth on full-resolution images (POL:171-186): zero -> column max, then metres
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float depth_to_metres(float d, float a, float span) {
  // min*100 + d*(max-min)*100 ; /100   (each step rounded to fp32 like the torch ops)
  float t = __fmul_rn(__fmul_rn(d, span), 100.0f);
  t = __fadd_rn(a, t);
  return __fdiv_rn(t, 100.0f);
}

__global__ void depth_full_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, float a, float span) {
  // block (32, 8): 32 columns x 8 row-lanes of one image
  __shared__ float smax[8][33];
  const int img = blockIdx.y;
  const int col = blockIdx.x * 32 + threadIdx.x;
  const float* src = in + (size_t)img * H * W;
  float m = -INFINITY;
  if (col < W)
    for (int r = threadIdx.y; r < H; r += 8) m = fmaxf(m, src[(size_t)r * W + col]);
  smax[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (col >= W) return;
#pragma unroll
  for (int y = 0; y < 8; ++y) m = fmaxf(m, smax[y][threadIdx.x]);
  float* dst = out + (size_t)img * H * W;
  for (int r = threadIdx.y; r < H; r += 8) {
    float d = src[(size_t)r * W + col];
    d = __fmul_rn(d, 1.0f);
    if (d == 0.0f) d = m;
    dst[(size_t)r * W + col] = depth_to_metres(d, a, span);
  }
}

// ------------------------------------------------------------------------------------------------
// a3: nearest-resize to the 24x24 patch grid + preprocess_depth (POL:336-341)
// ------------------------------------------------------------------------------------------------
struct GridIdx {
  int r[32];
  int c[32];
};

__global__ void depth_grid_kernel(const float* __restrict__ obs, float* __restrict__ out, int H, int W, int views, int literal_q1,
                                  GridIdx gi, int gh, int gw, float a, float span) {
  __shared__ float tile[32][33];
  const int unit = blockIdx.x;  // b * views + i
  const int b = unit / views, i = unit % views;
  const int r = threadIdx.y, c = threadIdx.x;
  if (r < gh && c < gw) {
    float v;
    if (literal_q1) v = obs[((size_t)b * H + i) * W + gi.r[r]];  // image b, row i, column idx[r]  (Q1)
    else v = obs[((size_t)unit * H + gi.r[r]) * W + gi.c[c]];
    tile[r][c] = v;
  }
  __syncthreads();
  if (r < gh && c < gw) {
    float m = -INFINITY;
    for (int y = 0; y < gh; ++y) m = fmaxf(m, tile[y][c]);
    float d = tile[r][c];
    if (d == 0.0f) d = m;
    out[(size_t)unit * gh * gw + r * gw + c] = depth_to_metres(d, a, span);
  }
}

// ------------------------------------------------------------------------------------------------
// a4 / a5: unprojection (FF:276-293, 296-326, 548-554)
// ------------------------------------------------------------------------------------------------
struct PixelTables {
  float tan_x[32];
  float tan_z[32];
  float neg_atan_x[32];
  float tan_h;
  float two_pi;
  int W, H;
};

__device__ __forceinline__ float np_mod_f32(float a, float b) {
  // numpy float32 `%`: fmod, then shift into [0,b) for b > 0; zero result takes the sign of b
  float m = fmodf(a, b);
  if (m != 0.0f) {
    if ((b < 0.0f) != (m < 0.0f)) m = __fadd_rn(m, b);
  } else {
    m = copysignf(0.0f, b);
  }
  return m;
}

// pose row: [px, py, pz (internal frame), cos(theta), sin(theta), theta] all fp32 (host rounds from double)
__global__ void unproject_kernel(const float* __restrict__ depth, const float* __restrict__ pose, PixelTables t, float* __restrict__ xyz,
                                 float* __restrict__ dir, float* __restrict__ scale, int n_units) {
  const int P = t.W * t.H;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_units * P) return;
  const int u = idx / P, p = idx % P;
  const int r = p / t.W, c = p % t.W;
  const float* ps = pose + (size_t)u * 6;
  const float d = depth[idx];
  const float cs = ps[3], sn = ps[4];
  const float dx = __fmul_rn(d, t.tan_x[c]);
  const float dz = __fmul_rn(d, t.tan_z[r]);
  const float sc = __fdiv_rn(__fmul_rn(__fmul_rn(d, t.tan_h), 2.0f), (float)t.W);
  const float dr = np_mod_f32(__fadd_rn(t.neg_atan_x[c], ps[5]), t.two_pi);
  const float rx = __fsub_rn(__fmul_rn(dx, cs), __fmul_rn(d, sn));
  const float ry = __fadd_rn(__fmul_rn(dx, sn), __fmul_rn(d, cs));
  xyz[(size_t)idx * 3 + 0] = __fadd_rn(rx, ps[0]);
  xyz[(size_t)idx * 3 + 1] = __fadd_rn(ry, ps[1]);
  xyz[(size_t)idx * 3 + 2] = __fadd_rn(dz, ps[2]);
  dir[idx] = dr;
  scale[idx] = sc;
}

// out [5, n, P]: rel_x, rel_y(=depth), rel_z, direction mod 2pi, scale
__global__ void patch_info_kernel(const float* __restrict__ depth, PixelTables t, float* __restrict__ out, int n_units) {
  const int P = t.W * t.H;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)n_units * P;
  if (idx >= n_units * P) return;
  const int p = idx % P;
  const int r = p / t.W, c = p % t.W;
  const float d = depth[idx];
  out[idx] = __fmul_rn(d, t.tan_x[c]);
  out[plane + idx] = d;
  out[2 * plane + idx] = __fmul_rn(d, t.tan_z[r]);
  out[3 * plane + idx] = np_mod_f32(t.neg_atan_x[c], t.two_pi);
  out[4 * plane + idx] = __fdiv_rn(__fmul_rn(__fmul_rn(d, t.tan_h), 2.0f), (float)t.W);
}

// ------------------------------------------------------------------------------------------------
// a6: frustum cull (FF:88-115, 349-360).  One thread per stored patch, loop over the step's views.
// ------------------------------------------------------------------------------------------------
struct CullParams {
  float fx, fy, cx, cy;
  float near_, far_, eps;
  int H, W, n_views;
};

// cam row: [cx, cy, cz (internal frame), cos(-heading), sin(-heading)]
__global__ void frustum_cull_kernel(float* __restrict__ xyz, float* __restrict__ dir, float* __restrict__ scale,
                                    const float* __restrict__ depth, const float* __restrict__ cam, CullParams p, int n,
                                    uint8_t* __restrict__ mask, int* __restrict__ n_deleted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool del = false;
  if (i < n) {
    const float x = xyz[(size_t)i * 3], y = xyz[(size_t)i * 3 + 1], z = xyz[(size_t)i * 3 + 2];
    for (int v = 0; v < p.n_views && !del; ++v) {
      const float* cm = cam + v * 5;
      const float px = __fsub_rn(x, cm[0]), py = __fsub_rn(y, cm[1]), pz = __fsub_rn(z, cm[2]);
      const float relx = __fsub_rn(__fmul_rn(px, cm[3]), __fmul_rn(py, cm[4]));
      const float rely = __fadd_rn(__fmul_rn(px, cm[4]), __fmul_rn(py, cm[3]));
      const float vx = relx, vy = -pz, vz = rely;
      const float uh = __fadd_rn(__fmul_rn(p.fx, vx), __fmul_rn(p.cx, vz));
      const float vh = __fadd_rn(__fmul_rn(p.fy, vy), __fmul_rn(p.cy, vz));
      const float uf = __fdiv_rn(uh, vz), vf = __fdiv_rn(vh, vz);
      if (!(isfinite(uf) && isfinite(vf))) continue;  // torch .to(int64) of inf/nan on x86 -> INT64_MIN -> fails u >= 0
      const float ut = truncf(uf), vt = truncf(vf);
      if (!(vz >= p.near_ && vz <= p.far_)) continue;
      if (!(ut >= 0.0f && ut <= (float)(p.W - 1) && vt >= 0.0f && vt <= (float)(p.H - 1))) continue;
      const int ui = (int)ut, vi = (int)vt;
      const float cd = depth[((size_t)v * p.H + vi) * p.W + ui];
      if (vz < __fadd_rn(cd, p.eps)) del = true;
    }
    mask[i] = del ? 1 : 0;
    if (del) {
      xyz[(size_t)i * 3] = -10000.0f;
      xyz[(size_t)i * 3 + 1] = -10000.0f;
      xyz[(size_t)i * 3 + 2] = -10000.0f;
      dir[i] = 0.0f;
      scale[i] = 0.0f;
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, del);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_deleted, __popc(ballot));
}

// Batched habitat cull: ALL episodes of the rank in one launch (blockIdx.y = episode), and the host no longer scans a mask over every
// stored patch: newly culled rows are compacted on the device into del_idx[job][0 .. n_del[job]) (warp-aggregated atomics; order is
// irrelevant to the bookkeeping, FF:362-393 visits sets) and their fp16 feature rows are zeroed by the warp that found them.
// Per-step host work of the cull therefore scales with the number of DELETED patches, not with the 147 k patches a long rollout stores.
__global__ void frustum_cull_batched_kernel(const d3d_cull_job* __restrict__ jobs, const float* __restrict__ depth_all, const float* __restrict__ cam_all,
                                            CullParams p, int fts_dim, int* __restrict__ del_idx, int del_cap, int* __restrict__ n_del) {
  const d3d_cull_job jb = jobs[blockIdx.y];
  const int n = jb.n_patches;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= n) return;
  float* xyz = (float*)jb.xyz;
  const float* depth = depth_all + (size_t)blockIdx.y * p.n_views * p.H * p.W;
  const float* cam = cam_all + (size_t)blockIdx.y * p.n_views * 5;
  bool del = false;
  if (i < n) {
    const float x = xyz[(size_t)i * 3], y = xyz[(size_t)i * 3 + 1], z = xyz[(size_t)i * 3 + 2];
    for (int v = 0; v < p.n_views && !del; ++v) {
      const float* cm = cam + v * 5;
      const float px = __fsub_rn(x, cm[0]), py = __fsub_rn(y, cm[1]), pz = __fsub_rn(z, cm[2]);
      const float relx = __fsub_rn(__fmul_rn(px, cm[3]), __fmul_rn(py, cm[4]));
      const float rely = __fadd_rn(__fmul_rn(px, cm[4]), __fmul_rn(py, cm[3]));
      const float vx = relx, vy = -pz, vz = rely;
      const float uh = __fadd_rn(__fmul_rn(p.fx, vx), __fmul_rn(p.cx, vz));
      const float vh = __fadd_rn(__fmul_rn(p.fy, vy), __fmul_rn(p.cy, vz));
      const float uf = __fdiv_rn(uh, vz), vf = __fdiv_rn(vh, vz);
      if (!(isfinite(uf) && isfinite(vf))) continue;
      const float ut = truncf(uf), vt = truncf(vf);
      if (!(vz >= p.near_ && vz <= p.far_)) continue;
      if (!(ut >= 0.0f && ut <= (float)(p.W - 1) && vt >= 0.0f && vt <= (float)(p.H - 1))) continue;
      const int ui = (int)ut, vi = (int)vt;
      const float cd = depth[((size_t)v * p.H + vi) * p.W + ui];
      if (vz < __fadd_rn(cd, p.eps)) del = true;
    }
    if (del) {
      xyz[(size_t)i * 3] = -10000.0f;
      xyz[(size_t)i * 3 + 1] = -10000.0f;
      xyz[(size_t)i * 3 + 2] = -10000.0f;
      ((float*)jb.dir)[i] = 0.0f;
      ((float*)jb.scale)[i] = 0.0f;
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, del);
  if (!ballot) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(n_del + blockIdx.y, __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (del) {
    const int pos = base + __popc(ballot & ((1u << lane) - 1));
    if (pos < del_cap) del_idx[(size_t)blockIdx.y * del_cap + pos] = i;
  }
  if (jb.fts16 && fts_dim) {  // the whole warp zeroes the feature row of each culled lane (FF:358)
    const int first = i - lane;
    for (unsigned m = ballot; m; m &= m - 1) {
      const int l = __ffs(m) - 1;
      uint4* dst = reinterpret_cast<uint4*>((uint16_t*)jb.fts16 + (size_t)(first + l) * fts_dim);
      for (int k = lane; k < fts_dim / 8; k += 32) dst[k] = make_uint4(0, 0, 0, 0);
    }
  }
}

// Posed-dataset form of the cull (get_frustum_mask FF:64-84 + z-test FF:349-353): cam row = [view matrix 4x4 row-major (world -> camera) |
// intrinsics 3x3 row-major] fp32.  The two einsums are left-to-right fp32 sums of separately rounded products (oracle/geometry.py).
__global__ void frustum_cull_matrix_kernel(float* __restrict__ xyz, float* __restrict__ dir, float* __restrict__ scale,
                                           const float* __restrict__ depth, const float* __restrict__ cam, CullParams p, int n,
                                           uint8_t* __restrict__ mask, int* __restrict__ n_deleted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool del = false;
  if (i < n) {
    const float x = xyz[(size_t)i * 3], y = xyz[(size_t)i * 3 + 1], z = xyz[(size_t)i * 3 + 2];
    for (int v = 0; v < p.n_views && !del; ++v) {
      const float* M = cam + v * 25;
      const float* K = M + 16;
      float vw[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        vw[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[r * 4], x), __fmul_rn(M[r * 4 + 1], y)), __fmul_rn(M[r * 4 + 2], z)), M[r * 4 + 3]);
      float uv[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        uv[r] = __fadd_rn(__fadd_rn(__fmul_rn(K[r * 3], vw[0]), __fmul_rn(K[r * 3 + 1], vw[1])), __fmul_rn(K[r * 3 + 2], vw[2]));
      const float uf = __fdiv_rn(uv[0], uv[2]), vf = __fdiv_rn(uv[1], uv[2]);
      if (!(isfinite(uf) && isfinite(vf))) continue;
      const float ut = truncf(uf), vt = truncf(vf);
      const float vz = vw[2];
      if (!(vz >= p.near_ && vz <= p.far_)) continue;
      if (!(ut >= 0.0f && ut <= (float)(p.W - 1) && vt >= 0.0f && vt <= (float)(p.H - 1))) continue;
      const int ui = (int)ut, vi = (int)vt;
      const float cd = depth[((size_t)v * p.H + vi) * p.W + ui];
      if (vz < __fadd_rn(cd, p.eps)) del = true;
    }
    mask[i] = del ? 1 : 0;
    if (del) {
      xyz[(size_t)i * 3] = -10000.0f;
      xyz[(size_t)i * 3 + 1] = -10000.0f;
      xyz[(size_t)i * 3 + 2] = -10000.0f;
      dir[i] = 0.0f;
      scale[i] = 0.0f;
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, del);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_deleted, __popc(ballot));
}

// ------------------------------------------------------------------------------------------------
// a4': posed-dataset unprojection (project_depth_to_3d FF:50-60 = open3d create_from_depth_image + nearest resize; rigid transform,
// heading and scale FF:536-546).  One thread per (view, grid patch).  vp [n,16] double = fx, fy, cx, cy, R (9, row-major), T (3).
// open3d arithmetic: z = float(u16)/float(depth_scale) (>= depth_trunc -> 0), x = (u-cx)*z/fx, y = (v-cy)*z/fy in double; the points are
// rounded to fp32 (FF:538) before R @ p + T in double; direction = -asin(dx/|xy|) (-pi if dy < 0), all rounded to fp32 at the end.
// ------------------------------------------------------------------------------------------------
__global__ void unproject_pinhole_kernel(const uint16_t* __restrict__ depth, int n, int H, int W, const double* __restrict__ vp, GridIdx gi,
                                         int gh, int gw, float depth_scale, float depth_trunc, float tan_abs, float* __restrict__ xyz,
                                         float* __restrict__ dir, float* __restrict__ scale, int* __restrict__ n_invalid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * gh * gw) return;
  const int view = idx / (gh * gw), pp = idx - view * gh * gw;
  const int r = gi.r[pp / gw], c = gi.c[pp % gw];
  unsigned d = depth[((size_t)view * H + r) * W + c];
  if (d == 0) d = 1;  // FF:51
  float z32 = __fdiv_rn((float)d, depth_scale);
  if (z32 >= depth_trunc) z32 = 0.f;
  if (!(z32 > 0.f)) atomicAdd(n_invalid, 1);  // open3d drops the pixel; the reference's view(H,W,3) then raises
  const double* q = vp + (size_t)view * 16;
  const double z = (double)z32;
  const double x = __ddiv_rn(__dmul_rn((double)c - q[2], z), q[0]);
  const double y = __ddiv_rn(__dmul_rn((double)r - q[3], z), q[1]);
  const float px = (float)x, py = (float)y, pz = z32;
  scale[idx] = __fdiv_rn(__fmul_rn(__fmul_rn(pz, tan_abs), 2.0f), (float)gw);
  const double X = (double)px, Y = (double)py, Z = (double)pz;
  double w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    w[k] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[4 + 3 * k], X), __dmul_rn(q[5 + 3 * k], Y)), __dmul_rn(q[6 + 3 * k], Z)), q[13 + k]);
  xyz[(size_t)idx * 3] = (float)w[0];
  xyz[(size_t)idx * 3 + 1] = (float)w[1];
  xyz[(size_t)idx * 3 + 2] = (float)w[2];
  double xy = sqrt(__dadd_rn(__dmul_rn(w[0], w[0]), __dmul_rn(w[1], w[1])));
  if (xy < 1e-4) xy = 1e-4;
  double h = -asin(w[0] / xy);
  if (w[1] < 0.0) h = h - 3.14159265358979323846;
  dir[idx] = (float)h;
}

// zero the fp16 feature rows of culled patches (FF:358); one warp per row
__global__ void zero_rows_kernel(uint16_t* __restrict__ fts, const uint8_t* __restrict__ mask, int n, int row_halves) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n || !mask[row]) return;
  uint4* dst = reinterpret_cast<uint4*>(fts + (size_t)row * row_halves);
  for (int i = lane; i < row_halves / 8; i += 32) dst[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// a9 / a18: exact K-NN, squared L2 in fp32 ((dx*dx + dy*dy) + dz*dz), ascending, lowest index on ties
// ------------------------------------------------------------------------------------------------
constexpr int KNN_MAXK = 8;
constexpr int KNN_TILE = 1024;

__device__ __forceinline__ float dist2(float qx, float qy, float qz, float rx, float ry, float rz) {
  const float dx = __fsub_rn(qx, rx), dy = __fsub_rn(qy, ry), dz = __fsub_rn(qz, rz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// lexicographic (d2, idx) insertion into an ascending list held in registers
template <int K>
__device__ __forceinline__ void topk_insert(float (&bd)[K], int (&bi)[K], float d, int i) {
  if (!(d < bd[K - 1] || (d == bd[K - 1] && i < bi[K - 1]))) return;
  bd[K - 1] = d;
  bi[K - 1] = i;
#pragma unroll
  for (int j = K - 1; j > 0; --j) {
    const bool sw = bd[j] < bd[j - 1] || (bd[j] == bd[j - 1] && bi[j] < bi[j - 1]);
    if (sw) {
      const float td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
      const int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
    }
  }
}

// large-Q variant: one thread per query, refs staged through shared memory (coalesced, each ref read once per block)
template <int K>
__global__ void __launch_bounds__(128) knn_thread_kernel(const float* __restrict__ refs, int n_ref, const float* __restrict__ qry, int n_q,
                                                         float* __restrict__ out_d, int* __restrict__ out_i) {
  __shared__ float sref[KNN_TILE * 3];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (q < n_q) { qx = qry[(size_t)q * 3]; qy = qry[(size_t)q * 3 + 1]; qz = qry[(size_t)q * 3 + 2]; }
  float bd[K];
  int bi[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { bd[j] = INFINITY; bi[j] = 0x7fffffff; }
  for (int base = 0; base < n_ref; base += KNN_TILE) {
    const int cnt = min(KNN_TILE, n_ref - base);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += blockDim.x) sref[t] = refs[(size_t)base * 3 + t];
    __syncthreads();
    if (q < n_q) {
      for (int r = 0; r < cnt; ++r) {
        const float d = dist2(qx, qy, qz, sref[3 * r], sref[3 * r + 1], sref[3 * r + 2]);
        topk_insert<K>(bd, bi, d, base + r);
      }
    }
  }
  if (q < n_q) {
#pragma unroll
    for (int j = 0; j < K; ++j) { out_d[(size_t)q * K + j] = bd[j]; out_i[(size_t)q * K + j] = bi[j]; }
  }
}

// small-Q variant: one warp per query; lanes stride over refs, local top-K per lane, then K rounds of warp arg-min
template <int K>
__global__ void __launch_bounds__(128) knn_warp_kernel(const float* __restrict__ refs, int n_ref, const float* __restrict__ qry, int n_q,
                                                       float* __restrict__ out_d, int* __restrict__ out_i) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= n_q) return;
  const float qx = qry[(size_t)q * 3], qy = qry[(size_t)q * 3 + 1], qz = qry[(size_t)q * 3 + 2];
  float bd[K];
  int bi[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { bd[j] = INFINITY; bi[j] = 0x7fffffff; }
  for (int r = lane; r < n_ref; r += 32) {
    const float d = dist2(qx, qy, qz, refs[(size_t)r * 3], refs[(size_t)r * 3 + 1], refs[(size_t)r * 3 + 2]);
    topk_insert<K>(bd, bi, d, r);
  }
  for (int j = 0; j < K; ++j) {
    // d2 >= 0 (or +inf): float bit pattern order == numeric order, so (bits << 32 | idx) sorts lexicographically
    unsigned long long key = ((unsigned long long)__float_as_uint(bd[0]) << 32) | (unsigned)bi[0];
    unsigned long long best = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    if (lane == 0) {
      out_d[(size_t)q * K + j] = __uint_as_float((unsigned)(best >> 32));
      out_i[(size_t)q * K + j] = (int)(best & 0xffffffffu);
    }
    if (key == best) {  // pop my head
#pragma unroll
      for (int t = 0; t < K - 1; ++t) { bd[t] = bd[t + 1]; bi[t] = bi[t + 1]; }
      bd[K - 1] = INFINITY;
      bi[K - 1] = 0x7fffffff;
    }
  }
}

template <int K>
int launch_knn(const float* refs, int n_ref, const float* qry, int n_q, float* out_d, int* out_i, cudaStream_t st) {
  if ((long long)n_q * 32 <= 148LL * 2048) {
    knn_warp_kernel<K><<<d3d_cdiv((long long)n_q * 32, 128), 128, 0, st>>>(refs, n_ref, qry, n_q, out_d, out_i);
  } else {
    knn_thread_kernel<K><<<d3d_cdiv(n_q, 128), 128, 0, st>>>(refs, n_ref, qry, n_q, out_d, out_i);
  }
  D3D_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// centroids: fp64 accumulate (exact for fp32 inputs -> order independent), one rounding to fp32.  One warp per sequence.
// ------------------------------------------------------------------------------------------------
__global__ void seq_centroid_kernel(const float* __restrict__ xyz, const int* __restrict__ member, const int* __restrict__ cu,
                                    int n_seq, float* __restrict__ out) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= n_seq) return;
  const int b = cu[s], e = cu[s + 1];
  double sx = 0, sy = 0, sz = 0;
  for (int t = b + lane; t < e; t += 32) {
    const int m = member[t];
    sx += (double)xyz[(size_t)m * 3];
    sy += (double)xyz[(size_t)m * 3 + 1];
    sz += (double)xyz[(size_t)m * 3 + 2];
  }
  sx = warp_sum_f64(sx); sy = warp_sum_f64(sy); sz = warp_sum_f64(sz);
  if (lane == 0) {
    const double n = (double)(e - b);  // n == 0 -> 0/0 = NaN, as torch's mean of an empty set (Q5)
    out[(size_t)s * 3] = (float)(sx / n);
    out[(size_t)s * 3 + 1] = (float)(sy / n);
    out[(size_t)s * 3 + 2] = (float)(sz / n);
  }
}

// ------------------------------------------------------------------------------------------------
// a13: agent-frame export (FF:818-862): gather ids, rotate/translate, keep dist <= radius, order-preserving compaction
// ------------------------------------------------------------------------------------------------
// agent row: [cx, cy, cz (internal), cos(-heading), sin(-heading)]
__device__ __forceinline__ void export_body(const float* __restrict__ pos, const float* __restrict__ fts, const int* __restrict__ ids,
                                            int n_ids, const float* __restrict__ agent, float radius, int width,
                                            float* __restrict__ out_rel, float* __restrict__ out_fts, int* __restrict__ out_count) {
  // single block; chunks of 1024 ids with a running base keep the dict order
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int start = 0; start < n_ids; start += 1024) {
    const int t = start + tid;
    bool keep = false;
    float rx = 0, ry = 0, rz = 0;
    int id = 0;
    if (t < n_ids) {
      id = ids[t];
      const float px = __fsub_rn(pos[(size_t)id * 3], agent[0]);
      const float py = __fsub_rn(pos[(size_t)id * 3 + 1], agent[1]);
      const float pz = __fsub_rn(pos[(size_t)id * 3 + 2], agent[2]);
      rx = __fsub_rn(__fmul_rn(px, agent[3]), __fmul_rn(py, agent[4]));
      ry = __fadd_rn(__fmul_rn(px, agent[4]), __fmul_rn(py, agent[3]));
      rz = pz;
      const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
      keep = __fsqrt_rn(n2) <= radius;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int in_warp = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    int total = 0;
    for (int w = 0; w < 32; ++w) total += warp_tot[w];
    const int slot = off + in_warp;
    if (keep) {
      out_rel[(size_t)slot * 3] = rx;
      out_rel[(size_t)slot * 3 + 1] = ry;
      out_rel[(size_t)slot * 3 + 2] = rz;
    }
    // feature rows: each kept thread's warp copies cooperatively
    for (int l = 0; l < 32; ++l) {
      const bool k = (bal >> l) & 1u;
      if (!k) continue;
      const int sid = __shfl_sync(0xffffffffu, id, l);
      const int sslot = __shfl_sync(0xffffffffu, slot, l);
      const float4* src = reinterpret_cast<const float4*>(fts + (size_t)sid * width);
      float4* dst = reinterpret_cast<float4*>(out_fts + (size_t)sslot * width);
      for (int i = lane; i < width / 4; i += 32) dst[i] = src[i];
    }
    __syncthreads();
    if (tid == 0) base_s += total;
    __syncthreads();
  }
  if (tid == 0) *out_count = base_s;
}

__global__ void __launch_bounds__(1024) export_kernel(const float* __restrict__ pos, const float* __restrict__ fts, const int* __restrict__ ids,
                                                      int n_ids, const float* __restrict__ agent, float radius, int width,
                                                      float* __restrict__ out_rel, float* __restrict__ out_fts, int* __restrict__ out_count) {
  export_body(pos, fts, ids, n_ids, agent, radius, width, out_rel, out_fts, out_count);
}

// all (episode, token kind) exports of a step in one launch: one block per job
struct ExportJob {
  const float* pos; const float* fts; long long ids_off; float* out_rel; float* out_fts;  // ids_off: first id of the job in `ids_all`
  float agent[5]; float radius; int n_ids; int pad;
};
__global__ void __launch_bounds__(1024) export_batched_kernel(const ExportJob* __restrict__ jobs, const int* __restrict__ ids_all, int width,
                                                              int* __restrict__ out_count) {
  const ExportJob j = jobs[blockIdx.x];
  __shared__ float ag[5];
  if (threadIdx.x < 5) ag[threadIdx.x] = j.agent[threadIdx.x];
  __syncthreads();
  export_body(j.pos, j.fts, ids_all + j.ids_off, j.n_ids, ag, j.radius, width, j.out_rel, j.out_fts, out_count + blockIdx.x);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int d3d_depth_preprocess(const float* obs, float* out, int n_img, int H, int W, float lo, float hi, void* stream) {
  D3D_REQUIRE(obs && out && n_img > 0 && H > 0 && W > 0, "args");
  const float a = (float)((double)lo * 100.0);
  const float span = (float)((double)hi - (double)lo);
  dim3 grid(d3d_cdiv(W, 32), n_img), block(32, 8);
  depth_full_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(obs, out, H, W, a, span);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_depth_patch_grid(const float* obs, float* out, int batch, int views, int H, int W, int gh, int gw, int literal_q1,
                                    const int* row_idx_h, const int* col_idx_h, float lo, float hi, void* stream) {
  D3D_REQUIRE(obs && out && row_idx_h && col_idx_h, "args");
  D3D_REQUIRE(gh <= 32 && gw <= 32 && gh > 0 && gw > 0, "grid up to 32x32");
  D3D_REQUIRE(!literal_q1 || views <= H, "literal (Q1) indexing reads image row i < H");
  GridIdx gi;
  for (int i = 0; i < 32; ++i) { gi.r[i] = i < gh ? row_idx_h[i] : 0; gi.c[i] = i < gw ? col_idx_h[i] : 0; }
  const float a = (float)((double)lo * 100.0);
  const float span = (float)((double)hi - (double)lo);
  depth_grid_kernel<<<batch * views, dim3(32, 32), 0, (cudaStream_t)stream>>>(obs, out, H, W, views, literal_q1, gi, gh, gw, a, span);
  D3D_CHECK_LAUNCH();
  return 0;
}

static int fill_tables(PixelTables& t, const float* tan_x_h, const float* tan_z_h, const float* neg_atan_x_h, float tan_h, int W, int H) {
  D3D_REQUIRE(W > 0 && H > 0 && W <= 32 && H <= 32, "patch grid up to 32x32");
  for (int i = 0; i < 32; ++i) {
    t.tan_x[i] = i < W ? tan_x_h[i] : 0.f;
    t.neg_atan_x[i] = i < W ? neg_atan_x_h[i] : 0.f;
    t.tan_z[i] = i < H ? tan_z_h[i] : 0.f;
  }
  t.tan_h = tan_h;
  t.two_pi = (float)(2.0 * 3.14159265358979323846);
  t.W = W;
  t.H = H;
  return 0;
}

extern "C" int d3d_unproject_habitat(const float* depth, const float* pose, int n_units, int W, int H, const float* tan_x_h,
                                     const float* tan_z_h, const float* neg_atan_x_h, float tan_h, float* xyz, float* dir, float* scale,
                                     void* stream) {
  D3D_REQUIRE(depth && pose && xyz && dir && scale && n_units > 0, "args");
  PixelTables t;
  D3D_TRY(fill_tables(t, tan_x_h, tan_z_h, neg_atan_x_h, tan_h, W, H));
  const long long total = (long long)n_units * W * H;
  unproject_kernel<<<d3d_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(depth, pose, t, xyz, dir, scale, n_units);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_patch_3d_info(const float* depth, int n_units, int W, int H, const float* tan_x_h, const float* tan_z_h,
                                 const float* neg_atan_x_h, float tan_h, float* out5, void* stream) {
  D3D_REQUIRE(depth && out5 && n_units > 0, "args");
  PixelTables t;
  D3D_TRY(fill_tables(t, tan_x_h, tan_z_h, neg_atan_x_h, tan_h, W, H));
  const long long total = (long long)n_units * W * H;
  patch_info_kernel<<<d3d_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(depth, t, out5, n_units);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_frustum_cull(float* xyz, float* dir, float* scale, void* fts16, int n_patches, int fts_dim, const float* depth,
                                int n_views, int H, int W, const float* cam, float fx, float fy, float cx, float cy, float near_,
                                float far_, float eps, uint8_t* mask, int* n_deleted, void* stream) {
  D3D_REQUIRE(xyz && dir && scale && depth && cam && mask && n_deleted, "args");
  D3D_REQUIRE(fts16 == nullptr || fts_dim % 8 == 0, "feature rows must be multiples of 16 B");
  cudaStream_t st = (cudaStream_t)stream;
  D3D_CHECK_CUDA(cudaMemsetAsync(n_deleted, 0, sizeof(int), st));
  if (n_patches == 0) return 0;
  CullParams p{fx, fy, cx, cy, near_, far_, eps, H, W, n_views};
  frustum_cull_kernel<<<d3d_cdiv(n_patches, 256), 256, 0, st>>>(xyz, dir, scale, depth, cam, p, n_patches, mask, n_deleted);
  D3D_CHECK_LAUNCH();
  if (fts16) {
    zero_rows_kernel<<<d3d_cdiv((long long)n_patches * 32, 256), 256, 0, st>>>((uint16_t*)fts16, mask, n_patches, fts_dim);
    D3D_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int d3d_frustum_cull_batched(const d3d_cull_job* jobs, int n_jobs, int max_patches, int fts_dim, const float* depth, int n_views, int H, int W,
                                        const float* cam, float fx, float fy, float cx, float cy, float near_, float far_, float eps, int* del_idx,
                                        int del_cap, int* n_del, void* stream) {
  D3D_REQUIRE(jobs && depth && cam && del_idx && n_del && n_jobs > 0 && n_jobs <= 65535, "args");
  D3D_REQUIRE(fts_dim % 8 == 0, "feature rows must be multiples of 16 B");
  cudaStream_t st = (cudaStream_t)stream;
  D3D_CHECK_CUDA(cudaMemsetAsync(n_del, 0, sizeof(int) * (size_t)n_jobs, st));
  if (max_patches == 0) return 0;
  CullParams p{fx, fy, cx, cy, near_, far_, eps, H, W, n_views};
  dim3 grid(d3d_cdiv(max_patches, 256), n_jobs);
  frustum_cull_batched_kernel<<<grid, 256, 0, st>>>(jobs, depth, cam, p, fts_dim, del_idx, del_cap, n_del);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_frustum_cull_matrix(float* xyz, float* dir, float* scale, void* fts16, int n_patches, int fts_dim, const float* depth,
                                       int n_views, int H, int W, const float* cam25, float near_, float far_, float eps, uint8_t* mask,
                                       int* n_deleted, void* stream) {
  D3D_REQUIRE(xyz && dir && scale && depth && cam25 && mask && n_deleted, "args");
  D3D_REQUIRE(fts16 == nullptr || fts_dim % 8 == 0, "feature rows must be multiples of 16 B");
  cudaStream_t st = (cudaStream_t)stream;
  D3D_CHECK_CUDA(cudaMemsetAsync(n_deleted, 0, sizeof(int), st));
  if (n_patches == 0) return 0;
  CullParams p{0.f, 0.f, 0.f, 0.f, near_, far_, eps, H, W, n_views};
  frustum_cull_matrix_kernel<<<d3d_cdiv(n_patches, 256), 256, 0, st>>>(xyz, dir, scale, depth, cam25, p, n_patches, mask, n_deleted);
  D3D_CHECK_LAUNCH();
  if (fts16) {
    zero_rows_kernel<<<d3d_cdiv((long long)n_patches * 32, 256), 256, 0, st>>>((uint16_t*)fts16, mask, n_patches, fts_dim);
    D3D_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int d3d_unproject_pinhole(const uint16_t* depth, int n_views, int H, int W, const double* view_params, int gh, int gw,
                                     const int* row_idx_h, const int* col_idx_h, float depth_scale, float depth_trunc, float tan_abs,
                                     float* xyz, float* dir, float* scale, int* n_invalid, void* stream) {
  D3D_REQUIRE(depth && view_params && row_idx_h && col_idx_h && xyz && dir && scale && n_invalid, "args");
  D3D_REQUIRE(gh <= 32 && gw <= 32 && gh > 0 && gw > 0, "grid up to 32x32");
  cudaStream_t st = (cudaStream_t)stream;
  D3D_CHECK_CUDA(cudaMemsetAsync(n_invalid, 0, sizeof(int), st));
  if (n_views == 0) return 0;
  GridIdx gi;
  for (int i = 0; i < 32; ++i) { gi.r[i] = i < gh ? row_idx_h[i] : 0; gi.c[i] = i < gw ? col_idx_h[i] : 0; }
  unproject_pinhole_kernel<<<d3d_cdiv((long long)n_views * gh * gw, 256), 256, 0, st>>>(depth, n_views, H, W, view_params, gi, gh, gw, depth_scale,
                                                                                        depth_trunc, tan_abs, xyz, dir, scale, n_invalid);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_knn3d(const float* refs, int n_ref, const float* queries, int n_q, int k, float* out_d2, int* out_idx, void* stream) {
  D3D_REQUIRE(k >= 0 && k <= KNN_MAXK, "k in [0, 8]");
  if (k == 0 || n_q == 0) return 0;
  D3D_REQUIRE(refs && queries && out_d2 && out_idx, "args");
  D3D_REQUIRE(n_ref >= k, "fewer reference points than k");
  cudaStream_t st = (cudaStream_t)stream;
  switch (k) {
    case 1: return launch_knn<1>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
    case 2: return launch_knn<2>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
    case 3: return launch_knn<3>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
    case 4: return launch_knn<4>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
    case 5: return launch_knn<5>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
    case 6: return launch_knn<6>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
    case 7: return launch_knn<7>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
    default: return launch_knn<8>(refs, n_ref, queries, n_q, out_d2, out_idx, st);
  }
}

extern "C" int d3d_seq_centroid(const float* xyz, const int* member, const int* cu_seqlens, int n_seq, float* out, void* stream) {
  if (n_seq == 0) return 0;
  D3D_REQUIRE(xyz && member && cu_seqlens && out, "args");
  seq_centroid_kernel<<<d3d_cdiv((long long)n_seq * 32, 128), 128, 0, (cudaStream_t)stream>>>(xyz, member, cu_seqlens, n_seq, out);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_env_export_batched(const void* jobs, const int* ids_all, int n_jobs, int width, int* out_count, void* stream) {
  if (n_jobs == 0) return 0;
  D3D_REQUIRE(jobs && ids_all && out_count, "args");
  D3D_REQUIRE(width % 4 == 0, "feature width must be a multiple of 4");
  static_assert(sizeof(ExportJob) == 72, "job layout is part of the C ABI (see include/dynam3d_b200.h)");
  export_batched_kernel<<<n_jobs, 1024, 0, (cudaStream_t)stream>>>((const ExportJob*)jobs, ids_all, width, out_count);
  D3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int d3d_env_export(const float* pos, const float* fts, const int* ids, int n_ids, const float* agent, float radius, int width,
                              float* out_rel, float* out_fts, int* out_count, void* stream) {
  D3D_REQUIRE(out_count != nullptr, "args");
  D3D_REQUIRE(width % 4 == 0, "feature width must be a multiple of 4");
  export_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pos, fts, ids, n_ids, agent, radius, width, out_rel, out_fts, out_count);
  D3D_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// a7: FastSAM masks -> 24x24 dense segment labels (FF:411-422).  One block per image.
//   paint masks in order (later mask wins; pixels in no mask keep label 0), nearest-resize to gh x gw, relabel the labels
//   that occur to 0..G-1 in ascending order.
// ------------------------------------------------------------------------------------------------
namespace {
struct SegIdx { int r[32]; int c[32]; };

__global__ void __launch_bounds__(1024) segm_relabel_kernel(const uint8_t* __restrict__ masks, int M, int H, int W, SegIdx si, int gh, int gw,
                                                            long long* __restrict__ out, int* __restrict__ n_seg) {
  extern __shared__ int present[];  // [M+1] occurrence flags, then exclusive ranks
  const int img = blockIdx.x;
  const uint8_t* mk = masks + (size_t)img * M * H * W;
  const int r = threadIdx.y, c = threadIdx.x;
  const int tid = r * blockDim.x + c;
  for (int i = tid; i <= M; i += blockDim.x * blockDim.y) present[i] = 0;
  __syncthreads();
  int label = 0;
  if (r < gh && c < gw) {
    const size_t pix = (size_t)si.r[r] * W + si.c[c];
    for (int g = M - 1; g >= 0; --g)
      if (mk[(size_t)g * H * W + pix]) { label = g; break; }
    present[label] = 1;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < M; ++i) { const int p = present[i]; present[i] = run; run += p; }
    n_seg[img] = run;
  }
  __syncthreads();
  if (r < gh && c < gw) out[(size_t)img * gh * gw + r * gw + c] = (long long)present[label];
}
}  // namespace

extern "C" int d3d_segm_relabel(const uint8_t* masks, int n_img, int M, int H, int W, int gh, int gw, const int* row_idx_h,
                                const int* col_idx_h, int64_t* out, int* n_seg, void* stream) {
  D3D_REQUIRE(masks && out && n_seg && row_idx_h && col_idx_h && n_img > 0 && M > 0, "args");
  D3D_REQUIRE(gh <= 32 && gw <= 32 && M <= 8192, "grid up to 32x32, at most 8192 masks");
  SegIdx si;
  for (int i = 0; i < 32; ++i) { si.r[i] = i < gh ? row_idx_h[i] : 0; si.c[i] = i < gw ? col_idx_h[i] : 0; }
  segm_relabel_kernel<<<n_img, dim3(32, 32), (M + 1) * sizeof(int), (cudaStream_t)stream>>>(masks, M, H, W, si, gh, gw, (long long*)out, n_seg);
  D3D_CHECK_LAUNCH();
  return 0;
}
