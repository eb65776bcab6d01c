"""Thin Python wrappers over the C ABI (one function per exported kernel family).

All tensors are CUDA tensors owned by the caller; nothing here computes on the host or falls back to torch ops.
"""
import ctypes

import torch

from . import _lib as L


GEMM_PROFILE = None  # set to a list to record (flops, start_event, end_event) per tensor-core GEMM launch (bench.py roofline)
# Per-stage roofline profile (bench.py `stages`): set STAGE_PROFILE to a list and every wrapped op appends
# (stage, bound, work, start_event, end_event) with work = algorithmic FLOPs ("tensor") or bytes ("hbm") of the launch.  STAGE_TAG names the
# caller (vit / tower / lm / ff / proj) and is set by the engines.
STAGE_PROFILE = None
STAGE_TAG = "misc"


class _Rec:
    """Context manager: CUDA events around one C-ABI call on the current stream (no-op unless STAGE_PROFILE is a list)."""

    def __init__(self, name, bound, work):
        self.args = (name, bound, work)

    def __enter__(self):
        if STAGE_PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if STAGE_PROFILE is not None:
            self.e1.record()
            STAGE_PROFILE.append((STAGE_TAG + "." + self.args[0], self.args[1], float(self.args[2]), self.e0, self.e1))
        return False


def gemm(a, w, out=None, bias=None, act=L.ACT_NONE, residual=None, out_dtype=None, simt=False):
    """out = epi(a @ w.T).  a [M,K], w [N,K] fp16/bf16 row-major; see d3d_gemm in include/dynam3d_b200.h."""
    assert a.is_cuda and w.is_cuda and a.dtype == w.dtype and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if act == L.ACT_SWIGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=out_dtype or a.dtype)
    assert out.stride(1) == 1 and out.shape[0] == M and out.shape[1] == n_out
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
    args = L.GemmArgs(
        L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), out.stride(0), M, N, K,
        L.kind_of(a.dtype), L.kind_of(out.dtype), L.ptr(bias), int(act),
        L.ptr(residual), residual.stride(0) if residual is not None else 0)
    fn = L.lib().d3d_gemm_simt if simt else L.lib().d3d_gemm
    if STAGE_PROFILE is not None:
        with _Rec("gemm", "tensor", 2.0 * M * N * K):
            L.check(fn(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
        return out
    if GEMM_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(fn(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
        e1.record()
        GEMM_PROFILE.append((2.0 * M * N * K, e0, e1))
        return out
    L.check(fn(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------
# geometry (see include/dynam3d_b200.h for the reference call sites)
# ------------------------------------------------------------------------------------------------
import math

import numpy as np

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_float_p = ctypes.POINTER(ctypes.c_float)


def _hp_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_c_float_p)


def _hp_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_c_int_p)


def nearest_index_table(dst, src):
    """cv2.resize(INTER_NEAREST) source indices: floor(x * (1 / (dst/src))) clamped (host, double arithmetic)."""
    ifx = 1.0 / (float(dst) / float(src))
    return np.array([min(int(math.floor(x * ifx)), src - 1) for x in range(dst)], dtype=np.int32)


def pixel_tables(hfov, vfov, W=24, H=24):
    """Per-column / per-row tangent tables built with the reference's own expressions (FF:283-287)."""
    half_W, half_H = W // 2, H // 2
    tan_h = math.tan(math.pi * hfov / 360.0)
    tan_v = math.tan(math.pi * vfov / 360.0)
    tan_x = (np.array([i / half_W + 1 / W for i in range(-half_W, half_W)], np.float32) * tan_h).astype(np.float32)
    tan_z = (np.array([i / half_H - 1 / H for i in range(half_H, -half_H, -1)], np.float32) * tan_v).astype(np.float32)
    neg_atan_x = (-np.arctan(tan_x)).astype(np.float32)
    return tan_x, tan_z, neg_atan_x, float(np.float32(tan_h))


def depth_preprocess(obs, lo=0.0, hi=10.0, out=None):
    """obs [n,H,W] (or [n,H,W,1]) fp32 -> metres [n,H,W]."""
    n, H, W = obs.shape[0], obs.shape[1], obs.shape[2]
    assert obs.dtype == torch.float32 and obs.is_contiguous()
    if out is None:
        out = torch.empty((n, H, W), device=obs.device, dtype=torch.float32)
    L.check(L.lib().d3d_depth_preprocess(L.ptr(obs), L.ptr(out), n, H, W, ctypes.c_float(lo), ctypes.c_float(hi), L.stream_ptr()))
    return out


def depth_patch_grid(obs, batch, views, gh=24, gw=24, literal_q1=True, lo=0.0, hi=10.0):
    """obs [n_img,H,W(,1)] fp32 -> [batch*views, gh*gw] metres on the patch grid (POL:336-341)."""
    H, W = obs.shape[1], obs.shape[2]
    assert obs.dtype == torch.float32 and obs.is_contiguous()
    if literal_q1:
        ri, ci = nearest_index_table(gh, W), np.zeros(gw, np.int32)
    else:
        ri, ci = nearest_index_table(gh, H), nearest_index_table(gw, W)
    out = torch.empty((batch * views, gh * gw), device=obs.device, dtype=torch.float32)
    ri, rp = _hp_i32(ri)
    ci, cp = _hp_i32(ci)
    L.check(L.lib().d3d_depth_patch_grid(L.ptr(obs), L.ptr(out), batch, views, H, W, gh, gw, int(bool(literal_q1)), rp, cp,
                                         ctypes.c_float(lo), ctypes.c_float(hi), L.stream_ptr()))
    return out


def pose_rows(positions_hab, headings, num_views):
    """[B*V,6] fp32 rows (x, -z, y, cos(theta), sin(theta), theta), theta = ix*(-pi/6)+heading (FF:521-526,550)."""
    rows = []
    for pos, h in zip(positions_hab, headings):
        for ix in range(num_views):
            th = ix * (-math.pi / 6) + float(h)
            rows.append([float(pos[0]), -float(pos[2]), float(pos[1]), math.cos(th), math.sin(th), th])
    return np.asarray(rows, dtype=np.float32)


def camera_rows(position_hab, headings):
    """[V,5] fp32 rows (x, -z, y, cos(-h), sin(-h)) for the cull / export kernels (FF:95-99, 830-836)."""
    rows = [[float(position_hab[0]), -float(position_hab[2]), float(position_hab[1]), math.cos(-float(h)), math.sin(-float(h))]
            for h in headings]
    return np.asarray(rows, dtype=np.float32)


def unproject_habitat(depth, pose, hfov=90.0, vfov=90.0, W=24, H=24):
    """depth [n,W*H] metres, pose [n,6] (device) -> xyz [n,W*H,3], dir [n,W*H], scale [n,W*H]."""
    n = depth.shape[0]
    tx, tz, na, th = pixel_tables(hfov, vfov, W, H)
    xyz = torch.empty((n, W * H, 3), device=depth.device, dtype=torch.float32)
    d = torch.empty((n, W * H), device=depth.device, dtype=torch.float32)
    s = torch.empty((n, W * H), device=depth.device, dtype=torch.float32)
    tx, txp = _hp_f32(tx); tz, tzp = _hp_f32(tz); na, nap = _hp_f32(na)
    with _Rec("unproject", "hbm", n * (W * H * (4 + 12 + 4 + 4) + 24)):  # 13.8 KB / view
        L.check(L.lib().d3d_unproject_habitat(L.ptr(depth), L.ptr(pose), n, W, H, txp, tzp, nap, ctypes.c_float(th),
                                              L.ptr(xyz), L.ptr(d), L.ptr(s), L.stream_ptr()))
    return xyz, d, s


def patch_3d_info(depth, hfov=90.0, vfov=90.0, W=24, H=24):
    n = depth.shape[0]
    tx, tz, na, th = pixel_tables(hfov, vfov, W, H)
    out = torch.empty((5, n, W * H), device=depth.device, dtype=torch.float32)
    tx, txp = _hp_f32(tx); tz, tzp = _hp_f32(tz); na, nap = _hp_f32(na)
    L.check(L.lib().d3d_patch_3d_info(L.ptr(depth), n, W, H, txp, tzp, nap, ctypes.c_float(th), L.ptr(out), L.stream_ptr()))
    return out


def frustum_cull(xyz, direction, scale, fts16, n_patches, depth, cam, hfov=90.0, vfov=90.0, near=0.0, far=3.0, eps=0.1,
                 mask=None, n_deleted=None):
    """In-place tombstoning of the first `n_patches` rows; returns (mask u8 [n_patches], n_deleted int32 [1])."""
    V, H, W = depth.shape
    fx = float(np.float32(W / np.tan(np.deg2rad(hfov) / 2.0) / 2.0))
    fy = float(np.float32(H / np.tan(np.deg2rad(vfov) / 2.0) / 2.0))
    if mask is None:
        mask = torch.empty((max(n_patches, 1),), device=xyz.device, dtype=torch.uint8)
    if n_deleted is None:
        n_deleted = torch.zeros((1,), device=xyz.device, dtype=torch.int32)
    f = ctypes.c_float
    with _Rec("frustum_cull", "hbm", n_patches * 13 + V * H * W * 4):  # 12 B xyz + 1 B mask per stored patch + the depth maps
        L.check(L.lib().d3d_frustum_cull(L.ptr(xyz), L.ptr(direction), L.ptr(scale), L.ptr(fts16), n_patches,
                                         fts16.shape[1] if fts16 is not None else 0, L.ptr(depth), V, H, W, L.ptr(cam),
                                         f(fx), f(fy), f(W / 2.0), f(H / 2.0), f(near), f(far), f(eps), L.ptr(mask), L.ptr(n_deleted),
                                         L.stream_ptr()))
    return mask[:n_patches], n_deleted


def frustum_cull_matrix(xyz, direction, scale, fts16, n_patches, depth, cam25, near=0.0, far=2.0, eps=0.1):
    """Posed-dataset cull (FF:64-84, 343-353): depth [V,H,W] fp32 (device), cam25 [V,25] fp32 (device) = view matrix | intrinsics."""
    V, H, W = depth.shape
    mask = torch.empty((max(n_patches, 1),), device=xyz.device, dtype=torch.uint8)
    n_deleted = torch.zeros((1,), device=xyz.device, dtype=torch.int32)
    f = ctypes.c_float
    L.check(L.lib().d3d_frustum_cull_matrix(L.ptr(xyz), L.ptr(direction), L.ptr(scale), L.ptr(fts16), n_patches,
                                            fts16.shape[1] if fts16 is not None else 0, L.ptr(depth), V, H, W, L.ptr(cam25), f(near), f(far), f(eps),
                                            L.ptr(mask), L.ptr(n_deleted), L.stream_ptr()))
    return mask[:n_patches], n_deleted


def ray_direction0(fx, gw=24, distance=3.0, depth_trunc=1000.0):
    """rel_direction[0][-1] of get_rays (PFF:390-405; FF:262-273 is the same expression but cannot run)."""
    z32 = np.float32(distance) / np.float32(1.0)
    if z32 >= np.float32(depth_trunc):
        raise ValueError("get_rays: constant depth >= depth_trunc, open3d returns no points (FF:267)")
    z = float(z32)
    x = (0.0 - gw / 2) * z / float(fx)
    return -math.atan(x / z)


def unproject_pinhole(depth_u16, view_params, depth_scale, depth_trunc, tan_abs, gh=24, gw=24):
    """depth_u16 [n,H,W] int16/uint16 bit pattern (device), view_params [n,16] float64 (device) -> xyz [n,gh*gw,3], dir, scale, n_invalid."""
    n, H, W = depth_u16.shape
    assert depth_u16.element_size() == 2 and depth_u16.is_contiguous() and view_params.dtype == torch.float64 and view_params.is_contiguous()
    dev = depth_u16.device
    xyz = torch.empty((n, gh * gw, 3), device=dev, dtype=torch.float32)
    d = torch.empty((n, gh * gw), device=dev, dtype=torch.float32)
    s = torch.empty((n, gh * gw), device=dev, dtype=torch.float32)
    n_invalid = torch.zeros((1,), device=dev, dtype=torch.int32)
    ri, rp = _hp_i32(torch_nearest_index_table(gh, H))
    ci, cp = _hp_i32(torch_nearest_index_table(gw, W))
    f = ctypes.c_float
    L.check(L.lib().d3d_unproject_pinhole(L.ptr(depth_u16), n, H, W, L.ptr(view_params), gh, gw, rp, cp, f(depth_scale), f(depth_trunc),
                                          f(float(np.float32(tan_abs))), L.ptr(xyz), L.ptr(d), L.ptr(s), L.ptr(n_invalid), L.stream_ptr()))
    return xyz, d, s, n_invalid


def knn3d(refs, queries, k):
    """Exact K-NN (squared L2, ascending, lowest index on ties): -> (d2 [Q,k] fp32, idx [Q,k] int32)."""
    assert refs.dtype == torch.float32 and queries.dtype == torch.float32 and refs.is_contiguous() and queries.is_contiguous()
    Q = queries.shape[0]
    d2 = torch.empty((Q, k), device=refs.device, dtype=torch.float32)
    idx = torch.empty((Q, k), device=refs.device, dtype=torch.int32)
    L.check(L.lib().d3d_knn3d(L.ptr(refs), refs.shape[0], L.ptr(queries), Q, k, L.ptr(d2), L.ptr(idx), L.stream_ptr()))
    return d2, idx


def seq_centroid(xyz, member, cu_seqlens, n_seq):
    out = torch.empty((n_seq, 3), device=xyz.device, dtype=torch.float32)
    L.check(L.lib().d3d_seq_centroid(L.ptr(xyz), L.ptr(member), L.ptr(cu_seqlens), n_seq, L.ptr(out), L.stream_ptr()))
    return out


def env_export(pos, fts, ids, agent, radius, out_rel, out_fts, out_count):
    L.check(L.lib().d3d_env_export(L.ptr(pos), L.ptr(fts), L.ptr(ids), ids.numel(), L.ptr(agent), ctypes.c_float(radius), fts.shape[1],
                                   L.ptr(out_rel), L.ptr(out_fts), L.ptr(out_count), L.stream_ptr()))


# ------------------------------------------------------------------------------------------------
# normalisation / elementwise / attention
# ------------------------------------------------------------------------------------------------
def layernorm(x, gamma, beta, eps, out32=None, out16=None, act=L.ACT_NONE, row_index=None, n_rows=None):
    T = n_rows if n_rows is not None else (row_index.numel() if row_index is not None else x.shape[0])
    D = x.shape[1]
    assert x.dtype == torch.float32 and x.stride(1) == 1
    with _Rec("layernorm", "hbm", T * D * (4 + (4 if out32 is not None else 0) + (2 if out16 is not None else 0))):
        L.check(L.lib().d3d_layernorm(L.ptr(x), x.stride(0), L.ptr(row_index), L.ptr(gamma), L.ptr(beta), ctypes.c_float(eps), T, D, int(act),
                                      L.ptr(out32), out32.stride(0) if out32 is not None else 0,
                                      L.ptr(out16), out16.stride(0) if out16 is not None else 0,
                                      L.kind_of(out16.dtype) if out16 is not None else 0, L.stream_ptr()))


def rmsnorm(x, w, eps, out32=None, out16=None, row_index=None, n_rows=None):
    T = n_rows if n_rows is not None else (row_index.numel() if row_index is not None else x.shape[0])
    D = x.shape[1]
    assert x.dtype == torch.float32 and x.stride(1) == 1
    with _Rec("rmsnorm", "hbm", T * D * (4 + (4 if out32 is not None else 0) + (2 if out16 is not None else 0))):
        L.check(L.lib().d3d_rmsnorm(L.ptr(x), x.stride(0), L.ptr(row_index), L.ptr(w), ctypes.c_float(eps), T, D,
                                    L.ptr(out32), out32.stride(0) if out32 is not None else 0,
                                    L.ptr(out16), out16.stride(0) if out16 is not None else 0,
                                    L.kind_of(out16.dtype) if out16 is not None else 0, L.stream_ptr()))


def rope(qkv, pos, inv_freq, H, Dh):
    L.check(L.lib().d3d_rope(L.ptr(qkv), qkv.stride(0), L.ptr(pos), L.ptr(inv_freq), qkv.shape[0], H, Dh, L.kind_of(qkv.dtype), L.stream_ptr()))


def rope_table(pos, inv_freq, Dh):
    tab = torch.empty((pos.numel(), Dh), device=pos.device, dtype=torch.float32)
    L.check(L.lib().d3d_rope_table(L.ptr(pos), L.ptr(inv_freq), pos.numel(), Dh, L.ptr(tab), L.stream_ptr()))
    return tab


def rope_apply(qkv, tab, H, Dh):
    with _Rec("rope", "hbm", qkv.shape[0] * (2 * H * Dh * 2 * 2 + Dh * 4)):  # q and k read + written (16 bit), cos/sin row read
        L.check(L.lib().d3d_rope_apply(L.ptr(qkv), qkv.stride(0), L.ptr(tab), qkv.shape[0], H, Dh, L.kind_of(qkv.dtype), L.stream_ptr()))


def embed_gather(table, ids, out):
    L.check(L.lib().d3d_embed_gather(L.ptr(table), L.kind_of(table.dtype), L.ptr(ids), ids.numel(), table.shape[1], L.ptr(out), out.stride(0),
                                     L.stream_ptr()))


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def preprocess_im2col(img_u8, R=336, patch=14, out_dtype=torch.float16, mean=CLIP_MEAN, std=CLIP_STD):
    """img_u8 [N,H,W,3] uint8 (device) -> im2col patches [N*(R/patch)^2, kpad] with kpad = 3*patch^2 rounded up to 8."""
    N, Hin, Win, _ = img_u8.shape
    assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous()
    k = 3 * patch * patch
    kpad = (k + 7) // 8 * 8
    g = R // patch
    out = torch.empty((N * g * g, kpad), device=img_u8.device, dtype=out_dtype)
    m, mp = _hp_f32(mean)
    s, sp = _hp_f32(std)
    with _Rec("preprocess_im2col", "hbm", N * (Hin * Win * 3 + g * g * kpad * 2)):
        L.check(L.lib().d3d_preprocess_im2col(L.ptr(img_u8), N, Hin, Win, R, patch, mp, sp, L.ptr(out), kpad, L.kind_of(out_dtype), L.stream_ptr()))
    return out


_PIL_TABLES = {}


def pil_bicubic_tables(in_size, out_size):
    """Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc` for the BICUBIC filter (a = -0.5, support 2, widened when down-scaling):
    (bounds int32 [out,2] = first source index / tap count, kk int32 [out,ksize] = 22-bit fixed-point taps).  Host, double arithmetic."""
    def filt(x):
        a = -0.5
        x = abs(x)
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - xmin
        k = [filt((x + xmin - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for w in k:
            ww += w
        for x, w in enumerate(k):
            if ww != 0.0:
                w = w / ww
            kk[xx, x] = int(-0.5 + w * (1 << 22)) if w < 0 else int(0.5 + w * (1 << 22))
        bounds[xx] = (xmin, n)
    return bounds, kk


def pil_bicubic_resize(img_u8, out_h, out_w):
    """PIL `Image.resize((out_w, out_h), BICUBIC)` of uint8 NHWC device images, bit for bit (HF CLIPImageProcessor, POL:438)."""
    assert img_u8.dtype == torch.uint8 and img_u8.is_cuda and img_u8.is_contiguous()
    N, H, W, C = img_u8.shape
    x = img_u8
    for axis, n_in, n_out in ((1, W, out_w), (0, H, out_h)):  # horizontal pass first, uint8 intermediate (ImagingResample)
        if n_in == n_out:
            continue
        key = (n_in, n_out, str(x.device))
        if key not in _PIL_TABLES:
            b, k = pil_bicubic_tables(n_in, n_out)
            _PIL_TABLES[key] = (torch.from_numpy(b).to(x.device), torch.from_numpy(k).to(x.device))
        b, k = _PIL_TABLES[key]
        h, w = x.shape[1], x.shape[2]
        dst = torch.empty((N, n_out if axis == 0 else h, n_out if axis == 1 else w, C), device=x.device, dtype=torch.uint8)
        with _Rec("pil_resize", "hbm", x.numel() + dst.numel()):
            L.check(L.lib().d3d_pil_resample_pass(L.ptr(x), L.ptr(dst), N, h, w, C, n_out, axis, L.ptr(b), L.ptr(k), k.shape[1], L.stream_ptr()))
        x = dst
    return x


def vit_embed_ln(conv, cls, pos, gamma, beta, eps, N, tokens, out):
    L.check(L.lib().d3d_vit_embed_ln(L.ptr(conv), L.ptr(cls), L.ptr(pos), L.ptr(gamma), L.ptr(beta), ctypes.c_float(eps), N, tokens,
                                     conv.shape[1], L.ptr(out), L.stream_ptr()))


def scatter_rows(src, dst, n, src_idx=None, dst_idx=None):
    L.check(L.lib().d3d_scatter_rows(L.ptr(src), src.stride(0), L.ptr(src_idx), L.ptr(dst), dst.stride(0), L.ptr(dst_idx), n, src.shape[1],
                                     L.stream_ptr()))


def add_inplace(a, b):
    assert a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()
    L.check(L.lib().d3d_add_inplace(L.ptr(a), L.ptr(b), a.numel() // a.shape[-1], a.shape[-1], L.stream_ptr()))


def cast16(x, out):
    L.check(L.lib().d3d_cast16(L.ptr(x), x.stride(0), L.ptr(out), out.stride(0), x.shape[0], x.shape[1], L.kind_of(out.dtype), L.stream_ptr()))


def attention_simt(qkv, out, cu_seqlens, n_seq, max_len, H, Dh, causal=False, scale=None):
    scale = (1.0 / math.sqrt(Dh)) if scale is None else scale
    L.check(L.lib().d3d_attention_simt(L.ptr(qkv), qkv.stride(0), L.ptr(out), out.stride(0), L.ptr(cu_seqlens), n_seq, max_len, H, Dh,
                                       int(bool(causal)), L.kind_of(qkv.dtype), ctypes.c_float(scale), L.stream_ptr()))


def attention(qkv, out, cu_seqlens, n_seq, max_len, H, Dh, causal=False, uniform_len=None, impl="auto"):
    """Self-attention over packed sequences.  `impl`: 'tc' (tcgen05, head_dim 64 / 96), 'mma' (legacy tensor cores), 'mixed' (head_dim 64,
    non-causal: sequences of <= 64 tokens one warp per (sequence, head), longer ones on the mma kernel -- the pooling passes), 'simt' (fp32
    CUDA cores) or 'auto' (tc for long head_dim-64 / 96 sequences, mma otherwise, simt for very short ones)."""
    scale = 1.0 / math.sqrt(Dh)
    # algorithmic FLOPs of the launch (QK^T + PV), assuming equal-length sequences (exact for the ViT, ~2 % high for the packed LM batch)
    work = 4.0 * (qkv.shape[0] ** 2 / max(n_seq, 1)) * Dh * H * (0.5 if causal else 1.0)
    with _Rec("attention", "tensor", work):
        if impl == "tc" or (impl == "auto" and Dh in (64, 96) and max_len >= 256):
            L.check(L.lib().d3d_attention_tc(L.ptr(qkv), qkv.stride(0), qkv.shape[0], L.ptr(out), out.stride(0), L.ptr(cu_seqlens), n_seq, max_len, H, Dh,
                                             int(bool(causal)), L.kind_of(qkv.dtype), scale, L.stream_ptr()))
        elif impl == "mixed":
            L.check(L.lib().d3d_attention_mixed(L.ptr(qkv), qkv.stride(0), L.ptr(out), out.stride(0), L.ptr(cu_seqlens), n_seq, max_len, H, Dh,
                                                L.kind_of(qkv.dtype), scale, L.stream_ptr()))
        elif impl == "mma" or (impl == "auto" and Dh in (64, 96) and max_len >= 64):
            L.check(L.lib().d3d_attention_mma(L.ptr(qkv), qkv.stride(0), L.ptr(out), out.stride(0), L.ptr(cu_seqlens), n_seq, max_len, H, Dh,
                                              int(bool(causal)), L.kind_of(qkv.dtype), scale, L.stream_ptr()))
        else:
            attention_simt(qkv, out, cu_seqlens, n_seq, max_len, H, Dh, causal)


# ------------------------------------------------------------------------------------------------
# pooling token builders / discriminator / policy operands
# ------------------------------------------------------------------------------------------------
def pool_features(seq_xyz, seq_dir, seq_scale, centre, tok_seq, tok_src, T, mode, out16):
    L.check(L.lib().d3d_pool_features(L.ptr(seq_xyz), L.ptr(seq_dir), L.ptr(seq_scale), L.ptr(centre), L.ptr(tok_seq), L.ptr(tok_src), T, mode,
                                      L.ptr(out16), L.kind_of(out16.dtype), L.stream_ptr()))


def pool_assemble(emb, seq_fts, fts_is_f32, tok_seq, tok_src, agg, T, X):
    L.check(L.lib().d3d_pool_assemble(L.ptr(emb), L.ptr(seq_fts), int(fts_is_f32), L.ptr(tok_seq), L.ptr(tok_src), L.ptr(agg), T, X.shape[1],
                                      L.ptr(X), L.stream_ptr()))


def disc_input(inst_fts, inst_pos, idx, view_fts, centre, G, K, out16):
    L.check(L.lib().d3d_disc_input(L.ptr(inst_fts), L.ptr(inst_pos), L.ptr(idx), L.ptr(view_fts), L.ptr(centre), G, K, inst_fts.shape[1],
                                   out16.stride(0), L.ptr(out16), L.kind_of(out16.dtype), L.stream_ptr()))


def patch_info_rows(info5, out16):
    n = info5.shape[1] * info5.shape[2]
    L.check(L.lib().d3d_patch_info_rows(L.ptr(info5), n, L.ptr(out16), L.kind_of(out16.dtype), L.stream_ptr()))


def concat2_cast(a, b, out16):
    L.check(L.lib().d3d_concat2_cast(L.ptr(a), L.ptr(b), a.shape[0], a.shape[1], L.ptr(out16), L.kind_of(out16.dtype), L.stream_ptr()))


def pos3_rows(x, out16):
    L.check(L.lib().d3d_pos3_rows(L.ptr(x), x.shape[0], L.ptr(out16), L.kind_of(out16.dtype), L.stream_ptr()))


def torch_nearest_index_table(dst, src):
    """Source indices of F.interpolate(mode='nearest'): min(floor(dst_index * (src/dst) as float32), src-1)."""
    scale = np.float32(src) / np.float32(dst)
    return np.minimum(np.floor(np.arange(dst, dtype=np.float32) * scale).astype(np.int32), src - 1).astype(np.int32)


def segm_relabel(masks_u8, gh=24, gw=24):
    """masks_u8 [n_img, M, H, W] uint8 (device) -> (labels int64 [n_img, gh, gw], n_seg int32 [n_img]) (FF:411-422)."""
    n, M, H, W = masks_u8.shape
    assert masks_u8.dtype == torch.uint8 and masks_u8.is_contiguous()
    out = torch.empty((n, gh, gw), device=masks_u8.device, dtype=torch.int64)
    n_seg = torch.empty((n,), device=masks_u8.device, dtype=torch.int32)
    ri, rp = _hp_i32(torch_nearest_index_table(gh, H))
    ci, cp = _hp_i32(torch_nearest_index_table(gw, W))
    L.check(L.lib().d3d_segm_relabel(L.ptr(masks_u8), n, M, H, W, gh, gw, rp, cp, L.ptr(out), L.ptr(n_seg), L.stream_ptr()))
    return out, n_seg
