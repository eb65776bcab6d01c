"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp32, fixed operation order) of the
geometry / integer part of Dynam3D's per-step hot path.  Parity status: PINNED against the
reference's own Python run in this container (oracle/make_golden.py, tests/test_oracle_vs_reference.py);
the reference ships no tests or golden vectors of its own.

Every function cites the reference lines it restates.  Arithmetic contract (what "bit-exact"
means for the CUDA kernels): every fp32 add / mul / sub / div below is a separately rounded
IEEE-754 binary32 operation in exactly the written order (no FMA contraction); per-pixel
tangent / arctan tables are computed once on the host with the reference's own expressions.

FF   = Dynam3D_VLN/vlnce_baselines/models/feature_fields.py
POL  = Dynam3D_VLN/vlnce_baselines/models/Policy_Dynam3D_VLN.py
"""
import math

import numpy as np

F32 = np.float32
TWO_PI_F32 = F32(2 * math.pi)


# ----------------------------------------------------------------------------------------------
# a3: depth resize + preprocess (POL:171-186, 336-341)
# ----------------------------------------------------------------------------------------------
def cv2_nearest_index(dst, src):
    """Source index table of cv2.resize(..., INTER_NEAREST): floor(x * (1/(dst/src))), clamped.

    cv2 computes inv_scale = dst/src (double), ifx = 1/inv_scale, sx = cvFloor(x*ifx), min(sx, src-1).
    Checked against cv2 4.13 in tests/test_oracle_geometry.py.
    """
    inv_scale = float(dst) / float(src)
    ifx = 1.0 / inv_scale
    return np.array([min(int(math.floor(x * ifx)), src - 1) for x in range(dst)], dtype=np.int64)


def preprocess_depth(depth, depth_scale=(0.0, 10.0)):
    """POL:171-186. depth [B,H,W,1] fp32 in [0,1] -> metres; zeros replaced by the column max."""
    depth = np.asarray(depth, dtype=F32) * F32(1.0)
    lo, hi = depth_scale
    col_max = depth.max(axis=1, keepdims=True)
    col_max = np.broadcast_to(col_max, depth.shape)
    depth = np.where(depth == 0, col_max, depth).astype(F32)
    # torch: min*100.0 + depth*(max-min)*100.0 ; /100.   python scalars are folded in double first
    a = F32(lo * 100.0)
    depth = (a + (depth * F32(hi - lo)).astype(F32) * F32(100.0)).astype(F32)
    depth = (depth / F32(100.0)).astype(F32)
    return depth


def depth_patch_grid(obs_depth, batch_size, num_of_views, gh=24, gw=24, depth_scale=(0.0, 10.0), q1_fix=False):
    """POL:336-341: nearest-resize to the 24x24 patch grid, then preprocess_depth (default scale).

    Literal behaviour (Q1, SURVEY.md quirk list): `observations['depth'][b][i]` selects IMAGE ROW i
    of image b, shape [W,1]; cv2 resizes that [W rows,1 col] array to 24x24, so
    out[b,i,r,c] = depth[b, i, idx[r], 0].  With q1_fix the intended per-view image (b*V+i) is resized.
    Note POL:341 calls preprocess_depth with the DEFAULT scale (0,10) regardless of `depth_scale`.
    """
    obs_depth = np.asarray(obs_depth, dtype=F32)
    n_img, H, W, _ = obs_depth.shape
    out = np.zeros((batch_size * num_of_views, gh, gw, 1), dtype=F32)
    for b in range(batch_size):
        for i in range(num_of_views):
            if q1_fix:
                img = obs_depth[b * num_of_views + i, :, :, 0]
                ri = cv2_nearest_index(gh, H)
                ci = cv2_nearest_index(gw, W)
                out[b * num_of_views + i, :, :, 0] = img[ri][:, ci]
            else:
                row = obs_depth[b, i, :, 0]  # [W]
                ri = cv2_nearest_index(gh, W)  # cv2 sees a [W x 1] image: rows <- W
                out[b * num_of_views + i, :, :, 0] = np.repeat(row[ri][:, None], gw, axis=1)
    out = preprocess_depth(out, (0.0, 10.0))
    return out.reshape(batch_size, num_of_views, gh * gw)


# ----------------------------------------------------------------------------------------------
# a4 / a5: unprojection (FF:276-293, 296-326, 521-526, 548-554)
# ----------------------------------------------------------------------------------------------
def pixel_tables(hfov, vfov, W=24, H=24):
    """Per-column / per-row tangent tables exactly as FF:283-287 builds them (float32 arrays)."""
    half_W, half_H = W // 2, H // 2
    tan_h = math.tan(math.pi * hfov / 360.0)
    tan_v = math.tan(math.pi * vfov / 360.0)
    tan_x = np.array([i / half_W + 1 / W for i in range(-half_W, half_W)], F32) * tan_h  # [W] fp32
    tan_x = tan_x.astype(F32)
    neg_atan_x = (-np.arctan(tan_x)).astype(F32)
    tan_z = np.array([i / half_H - 1 / H for i in range(half_H, -half_H, -1)], F32) * tan_v  # [H]
    tan_z = tan_z.astype(F32)
    # scale = depth * tan_h * 2. / W  -> ((depth * f32(tan_h)) * f32(2.)) / f32(W)
    return tan_x, tan_z, neg_atan_x, F32(tan_h)


def np_mod_f32(a, b):
    """numpy float32 `%` (python-style floor mod), spelled out for the CUDA kernel to mirror."""
    a = np.asarray(a, dtype=F32)
    b = F32(b)
    m = np.fmod(a, b).astype(F32)
    fix = (m != 0) & ((m < 0) != (b < 0))
    m = np.where(fix, (m + b).astype(F32), m)
    m = np.where(m == 0, np.copysign(F32(0), b), m)
    return m.astype(F32)


def unproject_habitat(depth576, heading, hfov=90.0, vfov=90.0, W=24, H=24):
    """FF:276-293 with NumPy-1.21 float32 semantics (Q4): returns rel_x, rel_y, rel_z, direction, scale (fp32 [576])."""
    d = np.asarray(depth576, dtype=F32).reshape(-1)
    tan_x, tan_z, neg_atan_x, tan_h = pixel_tables(hfov, vfov, W, H)
    tx = np.tile(tan_x, H)  # index = r*W + c
    tz = np.repeat(tan_z, W)
    direction = np.tile(neg_atan_x, H)
    depth_x = (d * tx).astype(F32)
    depth_z = (d * tz).astype(F32)
    scale = (((d * tan_h).astype(F32) * F32(2.0)).astype(F32) / F32(W)).astype(F32)
    heading = float(heading)
    hf = F32(heading)
    c, s = F32(math.cos(heading)), F32(math.sin(heading))
    direction = np_mod_f32((direction + hf).astype(F32), TWO_PI_F32)
    rel_x = ((depth_x * c).astype(F32) - (d * s).astype(F32)).astype(F32)
    rel_y = ((depth_x * s).astype(F32) + (d * c).astype(F32)).astype(F32)
    return rel_x, rel_y, depth_z, direction, scale


def habitat_to_internal(position):
    """FF:336, 523, 830: (x, y, z)_habitat -> (x, -z, y)."""
    p = np.asarray(position, dtype=np.float64)
    return np.array([p[0], -p[2], p[1]], dtype=np.float64)


def unproject_view_world(depth576, position_hab, heading, view_ix, hfov=90.0, vfov=90.0):
    """FF:548-554: world-frame patch xyz [576,3] fp32 + direction + scale for panorama view `view_ix`."""
    pos = habitat_to_internal(position_hab)
    theta = view_ix * (-math.pi / 6) + float(heading)
    rx, ry, rz, direction, scale = unproject_habitat(depth576, theta, hfov, vfov)
    # rel_x + position[0]: float32 array + float scalar -> float32 (NumPy-1.21 value-based casting; Q4)
    xyz = np.stack([(rx + F32(pos[0])).astype(F32), (ry + F32(pos[1])).astype(F32), (rz + F32(pos[2])).astype(F32)], axis=-1)
    return xyz.astype(F32), direction, scale


def patch_3d_info(depth_maps, hfov=90.0, vfov=90.0, W=24, H=24):
    """FF:296-326: rel_x, rel_y(=depth), rel_z, direction mod 2pi, scale; each [N,576] fp32."""
    d = np.asarray(depth_maps, dtype=F32).reshape(len(depth_maps), -1)
    tan_x, tan_z, neg_atan_x, tan_h = pixel_tables(hfov, vfov, W, H)
    tx = np.tile(tan_x, H)[None]
    tz = np.repeat(tan_z, W)[None]
    direction = np_mod_f32(np.tile(neg_atan_x, H), TWO_PI_F32)[None].repeat(len(d), 0)
    rel_x = (d * tx).astype(F32)
    rel_z = (d * tz).astype(F32)
    scale = (((d * tan_h).astype(F32) * F32(2.0)).astype(F32) / F32(W)).astype(F32)
    return rel_x, d.copy(), rel_z, direction.astype(F32), scale


# ----------------------------------------------------------------------------------------------
# a6: frustum cull (FF:88-115, 329-360)
# ----------------------------------------------------------------------------------------------
def frustum_mask_habitat(points, depth_img, cam_pos_internal, heading, hfov=90.0, vfov=90.0, far=3.0, near=0.0, eps=0.1):
    """FF:88-115 + the z-test FF:349-353 for ONE view. points [N,3] fp32, depth_img [H,W] fp32 metres.

    Returns bool mask [N]: inside the frustum, within [near, far], and closer than observed depth + eps.
    """
    pts = np.asarray(points, dtype=F32)
    H, W = depth_img.shape
    fx = F32(W / np.tan(np.deg2rad(hfov) / 2.0) / 2.0)
    fy = F32(H / np.tan(np.deg2rad(vfov) / 2.0) / 2.0)
    cx, cy = F32(W / 2.0), F32(H / 2.0)
    h = -float(heading)
    c, s = F32(math.cos(h)), F32(math.sin(h))
    # points[:,0:1] - camera_position[0]: fp32 tensor minus python/NumPy double scalar -> fp32 op with fp32(scalar)
    px = (pts[:, 0] - F32(cam_pos_internal[0])).astype(F32)
    py = (pts[:, 1] - F32(cam_pos_internal[1])).astype(F32)
    pz = (pts[:, 2] - F32(cam_pos_internal[2])).astype(F32)
    rel_x = ((px * c).astype(F32) - (py * s).astype(F32)).astype(F32)
    rel_y = ((px * s).astype(F32) + (py * c).astype(F32)).astype(F32)
    vx, vy, vz = rel_x, (-pz).astype(F32), rel_y  # FF:102
    # einsum(intrinsics, view_points): u_h = fx*x + 0*y + cx*z ; v_h = fy*y + cy*z ; z_h = z
    with np.errstate(all="ignore"):
        uh = ((fx * vx).astype(F32) + (cx * vz).astype(F32)).astype(F32)
        vh = ((fy * vy).astype(F32) + (cy * vz).astype(F32)).astype(F32)
        uf = (uh / vz).astype(F32)
        vf = (vh / vz).astype(F32)
    finite = np.isfinite(uf) & np.isfinite(vf)
    # .to(torch.int64) truncates toward zero; non-finite -> INT64_MIN on x86 (fails u >= 0)
    u = np.where(finite, np.trunc(np.where(finite, uf, 0)), -1).astype(np.int64)
    v = np.where(finite, np.trunc(np.where(finite, vf, 0)), -1).astype(np.int64)
    mask = (vz >= F32(near)) & (vz <= F32(far)) & (u >= 0) & (u <= W - 1) & (v >= 0) & (v <= H - 1) & finite
    cam_depth = depth_img[np.mod(v, H), np.mod(u, W)].astype(F32)
    mask &= vz < (cam_depth + F32(eps)).astype(F32)
    return mask


# ----------------------------------------------------------------------------------------------
# a4': posed-dataset branch (FF:50-60, 250-273, 536-546, 64-84 + 343-344); PFF = Dynam3D_Pretrain/src_3dff/models/feature_fields.py
#
# open3d (pinned open3d==0.19.0, environment.yml) is NOT in the reference tree and not installable here: its
# PointCloud.create_from_depth_image / Image.ConvertDepthToFloatImage are restated from the published algorithm
# (cpp/open3d/geometry/PointCloudFactory.cpp, ImageFactory.cpp): z32 = float(u16) / float(depth_scale); z32 >= depth_trunc -> 0;
# for z > 0: x = (u - cx) * z / fx, y = (v - cy) * z / fy evaluated in double from the float z.  PARITY UNPINNED for this
# third-party step (no reference test holds its outputs); everything around it is pinned against the reference's own functions
# (tests/test_oracle_vs_reference.py::test_posed_*).
# ----------------------------------------------------------------------------------------------
def torch_nearest_index(dst, src):
    """Source index table of F.interpolate(mode='nearest'): min(floor(i * (src/dst) as fp32), src-1)."""
    scale = F32(src) / F32(dst)
    return np.minimum(np.floor(np.arange(dst, dtype=F32) * scale).astype(np.int64), src - 1)


def open3d_depth_to_points(depth_u16, fx, fy, cx, cy, depth_scale, depth_trunc):
    """create_from_depth_image on a uint16 image: [H,W,3] float64 camera-frame points and the validity mask (z > 0)."""
    d = np.asarray(depth_u16).astype(np.uint16)
    z32 = (d.astype(F32) / F32(depth_scale)).astype(F32)
    z32 = np.where(z32 >= F32(depth_trunc), F32(0), z32).astype(F32)
    H, W = z32.shape
    z = z32.astype(np.float64)
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    x = (u - float(cx)) * z / float(fx)
    y = (v - float(cy)) * z / float(fy)
    return np.stack([x, y, z], -1), z32 > 0


def project_depth_to_3d(depth, intrinsic, depth_scale, depth_trunc, gh=24, gw=24):
    """FF:50-60: depth (uint16-valued [H,W]) -> (points [gh*gw,3] float64 camera frame, mask z > 0.002).
    Zero depth is replaced by 1 raw unit (FF:51); open3d drops invalid pixels, which makes the reference's .view(H,W,3) raise --
    restated as ValueError."""
    d = np.array(depth).astype(np.int64)
    d[d == 0] = 1
    K = np.asarray(intrinsic, dtype=np.float64)
    pts, valid = open3d_depth_to_points(d, K[0][0], K[1][1], K[0][2], K[1][2], depth_scale, depth_trunc)
    if not valid.all():
        raise ValueError("open3d dropped %d pixels (depth >= depth_trunc): the reference's view(H,W,3) fails here (FF:55)" % int((~valid).sum()))
    ri, ci = torch_nearest_index(gh, pts.shape[0]), torch_nearest_index(gw, pts.shape[1])
    pts = pts[ri][:, ci].reshape(-1, 3)
    return pts, pts[:, 2] > 0.002


def ray_direction0(fx, gw=24, distance=3.0, depth_trunc=1000.0):
    """rel_direction[0][-1] of get_rays (PFF:390-405 / FF:262-273): -arctan(x/z) of grid pixel (0,0) on a constant-depth image with
    principal point (gw/2, gh/2).  Q14: the VLN copy passes depth_trunc=1. with a 3 m image (FF:267), so open3d would drop every
    point and the reshape at FF:270 raises; the Pretrain copy (depth_trunc=1000., PFF:397) is the working form restated here."""
    z32 = F32(distance) / F32(1.0)
    if z32 >= F32(depth_trunc):
        raise ValueError("get_rays: constant depth >= depth_trunc, open3d returns no points (FF:267, quirk Q14)")
    z = float(z32)
    x = (0.0 - gw / 2) * z / float(fx)
    return -math.atan(x / z)


def heading_angle_of_points(points):
    """FF:250-259 (float64 in, float64 out)."""
    p = np.asarray(points, dtype=np.float64)
    dx, dy = p[:, 0], p[:, 1]
    xy = np.sqrt(np.square(dx) + np.square(dy))
    xy[xy < 1e-4] = 1e-4
    h = -np.arcsin(dx / xy)
    h[dy < 0] = h[dy < 0] - np.pi
    return h


def unproject_posed_view(depth, intrinsic, R, T, depth_scale=1000.0, depth_trunc=1000.0, gh=24, gw=24, ray_distance=3.0, ray_fx=None):
    """FF:533-546: world-frame patch xyz [gh*gw,3] fp32, direction, scale for one posed image (R [3,3], T [3,1] float64).
    `ray_fx`: focal length get_rays was built with -- the reference uses the FIRST image of the batch for every view (FF:503)."""
    pts, _ = project_depth_to_3d(depth, intrinsic, depth_scale, depth_trunc, gh, gw)
    pts32 = pts.astype(F32)
    t = abs(math.tan(ray_direction0(np.asarray(intrinsic)[0][0] if ray_fx is None else ray_fx, gw, ray_distance)))
    scale = (((pts32[:, 2] * F32(t)).astype(F32) * F32(2.0)).astype(F32) / F32(gw)).astype(F32)
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    T = np.asarray(T, dtype=np.float64).reshape(3, 1)
    world = (R @ pts32.astype(np.float64).T + T).T
    return world.astype(F32), heading_angle_of_points(world).astype(F32), scale


def frustum_mask_matrix(points, depth_img, intrinsic, view_matrix, near=0.0, far=2.0, eps=0.1):
    """FF:64-84 + the z-test FF:349-353 for ONE posed view: points [N,3] fp32, depth_img [H,W] fp32, intrinsic [>=3,>=3], view_matrix
    [4,4] (world -> camera), all fp32.  The two einsums are evaluated as left-to-right fp32 sums of separately rounded products."""
    p = np.asarray(points, dtype=F32)
    K = np.asarray(intrinsic, dtype=F32)[:3, :3]
    M = np.asarray(view_matrix, dtype=F32)
    H, W = depth_img.shape

    def row4(r):
        a = ((M[r, 0] * p[:, 0]).astype(F32) + (M[r, 1] * p[:, 1]).astype(F32)).astype(F32)
        a = (a + (M[r, 2] * p[:, 2]).astype(F32)).astype(F32)
        return (a + M[r, 3]).astype(F32)  # homogeneous 1 * m3

    vx, vy, vz = row4(0), row4(1), row4(2)

    def row3(r):
        a = ((K[r, 0] * vx).astype(F32) + (K[r, 1] * vy).astype(F32)).astype(F32)
        return (a + (K[r, 2] * vz).astype(F32)).astype(F32)

    with np.errstate(all="ignore"):
        uh, vh, zh = row3(0), row3(1), row3(2)
        uf = (uh / zh).astype(F32)
        vf = (vh / zh).astype(F32)
    finite = np.isfinite(uf) & np.isfinite(vf)
    u = np.where(finite, np.trunc(np.where(finite, uf, 0)), -1).astype(np.int64)
    v = np.where(finite, np.trunc(np.where(finite, vf, 0)), -1).astype(np.int64)
    mask = (vz >= F32(near)) & (vz <= F32(far)) & (u >= 0) & (u <= W - 1) & (v >= 0) & (v <= H - 1) & finite
    cam_depth = np.asarray(depth_img, dtype=F32)[np.mod(v, H), np.mod(u, W)]
    mask &= vz < (cam_depth + F32(eps)).astype(F32)
    return mask


# ----------------------------------------------------------------------------------------------
# a9: exact K-NN (torch_kdtree semantics pinned by us: FF:606-612)
# ----------------------------------------------------------------------------------------------
def knn3d(refs, queries, k, chunk=2048):
    """Exact K nearest refs for each query: squared L2 in fp32 ((dx*dx + dy*dy) + dz*dz), ascending,
    lowest index wins ties.  Returns (d2 [Q,k] fp32, idx [Q,k] int32).

    torch_kdtree (pinned torch-kdtree==1.0, environment.yml:293) is absent from the reference tree and
    not installable offline -> its tie-break is UNPINNED; cross-checked with scipy cKDTree in tests.
    """
    refs = np.asarray(refs, dtype=F32)
    queries = np.asarray(queries, dtype=F32)
    Q = len(queries)
    out_d = np.zeros((Q, k), dtype=F32)
    out_i = np.zeros((Q, k), dtype=np.int32)
    if k == 0:
        return out_d, out_i
    for s in range(0, Q, chunk):
        q = queries[s:s + chunk]
        dx = (q[:, None, 0] - refs[None, :, 0]).astype(F32)
        dy = (q[:, None, 1] - refs[None, :, 1]).astype(F32)
        dz = (q[:, None, 2] - refs[None, :, 2]).astype(F32)
        d2 = ((dx * dx).astype(F32) + (dy * dy).astype(F32)).astype(F32)
        d2 = (d2 + (dz * dz).astype(F32)).astype(F32)
        if k < d2.shape[1]:
            part = np.argpartition(d2, k - 1, axis=1)[:, :k]
            kth = np.take_along_axis(d2, part, axis=1).max(axis=1, keepdims=True)
            # collect everything <= kth then stable-sort (handles ties deterministically)
            order = np.empty((len(q), k), dtype=np.int64)
            for r in range(len(q)):
                cand = np.nonzero(d2[r] <= kth[r, 0])[0]
                o = cand[np.argsort(d2[r, cand], kind="stable")][:k]
                order[r] = o
        else:
            order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        out_d[s:s + chunk] = np.take_along_axis(d2, order, axis=1)
        out_i[s:s + chunk] = order.astype(np.int32)
    return out_d, out_i


# ----------------------------------------------------------------------------------------------
# a12 helpers: zone voxel keys (FF:694-695)
# ----------------------------------------------------------------------------------------------
def zone_keys(pos, length=2.0):
    """(p // L) * L + L/2 in fp32 (torch floor_divide == floor for finite p, L = 2)."""
    p = np.asarray(pos, dtype=F32)
    L = F32(length)
    with np.errstate(all="ignore"):
        return ((np.floor(p / L).astype(F32) * L).astype(F32) + F32(length / 2)).astype(F32)


# ----------------------------------------------------------------------------------------------
# a13: agent-frame export geometry (FF:829-839)
# ----------------------------------------------------------------------------------------------
def to_agent_frame(pos, agent_position_hab, agent_heading):
    p = np.asarray(pos, dtype=F32).reshape(-1, 3)
    cam = habitat_to_internal(agent_position_hab)
    h = -float(agent_heading)
    c, s = F32(math.cos(h)), F32(math.sin(h))
    px = (p[:, 0] - F32(cam[0])).astype(F32)
    py = (p[:, 1] - F32(cam[1])).astype(F32)
    pz = (p[:, 2] - F32(cam[2])).astype(F32)
    rx = ((px * c).astype(F32) - (py * s).astype(F32)).astype(F32)
    ry = ((px * s).astype(F32) + (py * c).astype(F32)).astype(F32)
    rel = np.stack([rx, ry, pz], axis=-1).astype(F32)
    with np.errstate(all="ignore"):
        n2 = (((rx * rx).astype(F32) + (ry * ry).astype(F32)).astype(F32) + (pz * pz).astype(F32)).astype(F32)
        dist = np.sqrt(n2).astype(F32)
    return rel, dist


# ----------------------------------------------------------------------------------------------
# a7: FastSAM masks -> dense 24x24 labels (FF:411-422), restated with the reference's own torch ops
# ----------------------------------------------------------------------------------------------
def segm_relabel(masks, gh=24, gw=24):
    """masks [M,H,W] float/bool tensor-like (FastSAM `everything_prompt()` output) -> int64 [gh,gw] dense labels."""
    import torch
    masks = torch.as_tensor(np.asarray(masks)).to(torch.float32)
    patch_group = masks[0].clone()
    for group_id in range(masks.shape[0]):
        patch_group[masks[group_id] == 1] = group_id
    patch_group = torch.nn.functional.interpolate(patch_group.unsqueeze(0).unsqueeze(0), (gh, gw), mode="nearest").to(torch.int64).squeeze(0)
    patch_segm = patch_group.clone()
    group_id = 0
    for mask_id in torch.unique(patch_group).cpu().numpy().tolist():
        patch_segm[patch_group == mask_id] = group_id
        group_id += 1
    return patch_segm[0].numpy()
