"""Pins the CPU oracle (oracle/) against the UNMODIFIED reference run in this container (needs /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present (GPU box)")]


@pytest.mark.parametrize("cfg", [
    dict(seed=3, n_steps=5, num_views=1, n_seg=16, seg_kind="voronoi", merge_bias=0.3),
    dict(seed=5, n_steps=8, num_views=1, n_seg=48, seg_kind="blocks", merge_bias=0.3),
    dict(seed=6, n_steps=2, num_views=12, n_seg=16, seg_kind="voronoi", merge_bias=0.3),
    dict(seed=7, n_steps=2, num_views=3, n_seg=16, seg_kind="blocks", merge_bias=0.3, q1_fix=True),
])
def test_feature_fields_state_machine_matches_reference(cfg):
    from oracle import ref_compare as RC
    ff, orc, recs = RC.run_pair(**cfg)
    saw_merge = False
    for i, r in enumerate(recs):
        assert RC.snapshots_equal(r["ref"], r["orc"]) == [], f"step {i}"
        er, eo = r["env_ref"], r["env_orc"]
        for k in er:
            a, b = er[k][0].numpy(), eo[k][0]
            assert a.shape == b.shape, (i, k)
            # fp32 CPU both sides; the oracle accumulates centroids in fp64 (one rounding) -> last-ulp differences only
            assert np.allclose(a, b, atol=2e-5, rtol=1e-5, equal_nan=True), (i, k, np.abs(a - b).max())
        if r["merge"] is not None and r["merge"][0].any():
            saw_merge = True
    assert saw_merge or cfg["num_views"] == 3


def test_geometry_matches_reference_functions():
    from oracle import geometry as G
    mod = ref_shim.load_reference_feature_fields_module()
    ff = ref_shim.make_reference_feature_fields()
    rng = np.random.default_rng(0)
    depth = rng.uniform(0.1, 10, size=(1, 576)).astype(np.float32)
    for heading in (0.3, 4.0, 6.2):
        rx, ry, rz, d, s = ff.project_depth_to_3d_habitat(depth, heading)
        ox, oy, oz, od, os_ = G.unproject_habitat(depth[0], heading)
        for a, b in ((rx[0], ox), (ry[0], oy), (rz[0], oz), (d, od), (s, os_)):
            assert np.array_equal(np.asarray(a, np.float32), b)
    info = ff.get_patch_3d_info(depth)
    want = G.patch_3d_info(depth)
    for a, b in zip(info, want):
        assert np.array_equal(a[..., 0].numpy(), b)
    pts = rng.uniform(-5, 5, size=(50000, 3)).astype(np.float32)
    dimg = rng.uniform(0.5, 6, size=(256, 256)).astype(np.float32)
    cam = np.array([0.5, -0.25, 1.25])
    m, dep, u, v = mod.get_frustum_mask_habitat(torch.from_numpy(pts), 256, 256, 90.0, 90.0, cam, 1.1, far=3.0)
    u, v = u % 256, v % 256
    m = m & (dep < torch.from_numpy(dimg)[v, u] + 0.1)
    assert np.array_equal(m.numpy(), G.frustum_mask_habitat(pts, dimg, cam, 1.1))


def test_cv2_nearest_table_and_preprocess_depth():
    import cv2
    from oracle import geometry as G
    for src in (256, 224, 336, 240):
        img = np.arange(src * src, dtype=np.float32).reshape(src, src)
        out = cv2.resize(img, (24, 24), interpolation=cv2.INTER_NEAREST)
        idx = G.cv2_nearest_index(24, src)
        assert np.array_equal(out, img[idx][:, idx])
    # literal Q1 path (POL:339): a [W,1] slice resized to 24x24
    row = np.random.default_rng(1).random((256, 1)).astype(np.float32)
    lit = cv2.resize(row, (24, 24), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(lit, np.repeat(row[G.cv2_nearest_index(24, 256)], 24, axis=1))
    # preprocess_depth vs the torch expression (POL:171-186)
    d = np.random.default_rng(2).random((2, 24, 24, 1)).astype(np.float32)
    d[d < 0.1] = 0
    t = torch.from_numpy(d.copy()) * 1.0
    mx, _ = t.max(dim=1, keepdim=True)
    mx = mx.expand(-1, 24, -1, -1)
    t[t == 0] = mx[t == 0]
    t = (0.0 * 100.0 + t * (10.0 - 0.0) * 100.0) / 100.0
    assert np.array_equal(t.numpy(), G.preprocess_depth(d))


@pytest.mark.parametrize("pretrain", [False, True])
def test_vit_restatement_matches_reference_module(pretrain):
    from dynam3d_b200 import synth
    from oracle import nn_ops as NN
    mod = ref_shim.load_reference_clip_model_module(pretrain=pretrain)
    width, layers, heads, res, out = 256, 3, 4, 112, 128
    vit = mod.VisionTransformer(res, 14, width, layers, heads, out).eval()
    sd = synth.vit_state_dict(9, width=width, layers=layers, resolution=res, out_dim=out)
    vit.load_state_dict(sd, strict=True)
    x = synth.hash_uniform((2, 3, res, res), 77, 2.0)
    with torch.no_grad():
        ref_cls, ref_patch = vit(x)
    cls, patch = NN.vit_forward(x, sd, layers, heads, rnd=None, ln_post_on_patches=not pretrain)
    assert (ref_cls - cls).abs().max().item() < 2e-5 and (ref_patch - patch).abs().max().item() < 2e-5


def _posed_inputs(seed, H=240, W=320):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    depth = (1500 + 900 * np.sin(xx / 37.0 + seed) + 600 * np.cos(yy / 23.0) + rng.integers(0, 40, (H, W))).astype(np.uint16)
    depth[rng.integers(0, H, 30), rng.integers(0, W, 30)] = 0  # missing returns -> 1 raw unit (FF:51)
    K = np.array([[285.0 + seed, 0, W / 2 - 3.5], [0, 291.0, H / 2 + 2.25], [0, 0, 1]], np.float64)
    a = 0.3 * seed + 0.2
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.0]])
    T = rng.uniform(-2, 2, (3, 1))
    return depth, K, R, T


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_posed_unprojection_matches_reference_functions(seed):
    """a4': FF:50-60 (through the open3d stand-in), FF:250-259 and the glue of FF:533-546 against oracle/geometry.py."""
    from oracle import geometry as G
    mod = ref_shim.load_reference_feature_fields_module()
    ff = ref_shim.make_reference_feature_fields()
    depth, K, R, T = _posed_inputs(seed)
    pts_ref, mask_ref = mod.project_depth_to_3d(torch.from_numpy(depth.astype(np.int32)), K, 1000.0, 1000.0, 24, 24)
    pts, mask = G.project_depth_to_3d(depth, K, 1000.0, 1000.0, 24, 24)
    assert pts_ref.dtype == np.float64 and np.array_equal(pts_ref, pts) and np.array_equal(mask_ref, mask)
    # the reference's glue, literally (FF:536-546), with the working get_rays direction (Q14)
    points = pts_ref.astype(np.float32)
    t = abs(np.tan(G.ray_direction0(K[0][0], 24, 3.0)))
    scale_ref = points[:, -1] * float(t) * 2. / 24
    world = (R @ points.T + T).T
    xyz, direction, scale = G.unproject_posed_view(depth, K, R, T)
    assert scale_ref.dtype == np.float32 and np.array_equal(scale_ref, scale)
    assert np.array_equal(world.astype(np.float32), xyz)
    assert np.array_equal(ff.get_heading_angle(world).astype(np.float32), direction)


def test_posed_get_rays_quirk_q14():
    """The VLN copy of get_rays asks open3d for a 3 m image with depth_trunc=1 (FF:267): every point is dropped and FF:270 raises."""
    ff = ref_shim.make_reference_feature_fields()
    with pytest.raises(ValueError):
        ff.get_rays(np.array([[300.0, 0, 12], [0, 300.0, 12], [0, 0, 1]]))


@pytest.mark.parametrize("seed", [0, 3])
def test_posed_frustum_mask_matches_reference_function(seed):
    """a6 (dataset form): get_frustum_mask FF:64-84 + z-test FF:349-353 against oracle/geometry.frustum_mask_matrix."""
    from oracle import geometry as G
    mod = ref_shim.load_reference_feature_fields_module()
    rng = np.random.default_rng(seed)
    depth_u16, K, R, T = _posed_inputs(seed)
    H, W = depth_u16.shape
    depth_m = (depth_u16.astype(np.float32) / 1000.0).astype(np.float32)
    pts = rng.uniform(-3, 3, (40000, 3)).astype(np.float32)
    # world -> camera matrix of the pose (R, T) (camera -> world)
    M = np.eye(4)
    M[:3, :3] = R.T
    M[:3, 3:] = -R.T @ T
    M32, K32 = M.astype(np.float32), np.eye(4, dtype=np.float32)
    K32[:3, :3] = K.astype(np.float32)
    mask_ref, d_ref, u, v = mod.get_frustum_mask(torch.from_numpy(pts), H, W, torch.from_numpy(K32), torch.from_numpy(M32))
    cam_depth = torch.from_numpy(depth_m)[v % H, u % W]
    mask_ref = (mask_ref & (d_ref < cam_depth + 0.1)).numpy()
    mask = G.frustum_mask_matrix(pts, depth_m, K32, M32)
    # CPU einsum may contract / reorder the 4-term sums: identical except for points within one ulp of a pixel or depth boundary
    assert (mask_ref != mask).sum() <= 2 and mask.sum() > 100


def test_render_oracle_matches_reference_renderer():
    """a18: the unmodified Pretrain `render_view_3d_patch` (PFF:494-625) vs oracle/render_oracle.py on the fixture scene."""
    from oracle import nn_ops as NN
    from oracle import render_oracle as RO
    from oracle.make_golden import render_scene
    xyz, dr, sc, fts, pos, head, P = render_scene()
    ff = ref_shim.make_reference_pretrain_feature_fields()
    ff.load_state_dict(P, strict=False)
    f_ref, p_ref = ref_shim.reference_render_view(ff, xyz, dr, sc, fts, pos, head)
    want = RO.render_view_3d_patch(P, xyz, dr, sc, fts, pos, head, rnd=NN.round_fp16)
    valid = (want["idx"] >= 0).any(-1).any(-1)
    assert valid.sum() > 100
    assert np.array_equal(p_ref[valid], want["positions"][valid])  # rays without neighbours tie over all samples: topk order unspecified
    assert np.abs(f_ref - want["feature_map"]).max() < 5e-4


def test_waypoint_oracle_matches_reference():
    """8(f) rank 3: oracle/waypoint_oracle.py vs the reference's own BinaryDistPredictor_TRM (TRM_net.py:66-88) and `nms` (waypoint_pred/utils.py:37-66,
    applied as POL:226-247): logits to fp32 round-off, identical candidate cells, probabilities to 1e-5 relative."""
    import torch
    from dynam3d_b200 import synth
    from oracle import waypoint_oracle as WO
    trm, utils = ref_shim.load_reference_waypoint_predictor()
    for wseed, xseed, B in ((3, 4, 3), (8, 9, 2)):
        sd = synth.waypoint_state_dict(wseed)
        net = trm.BinaryDistPredictor_TRM(device="cpu").eval()
        net.load_state_dict(sd, strict=True)
        x = synth.waypoint_depth_embedding(xseed, B)
        with torch.no_grad():
            ref = net(None, x)
            bx = torch.softmax(ref.reshape(B, -1), 1).reshape(B, 120, 12)
            wrap = torch.cat((bx[:, -1:], bx, bx[:, :1]), 1)
            ref_map = utils.nms(wrap.unsqueeze(1), max_predictions=5, sigma=(7.0, 5.0)).squeeze(1)[:, 1:-1, :].numpy()
        got = WO.predictor_logits({k: v.numpy() for k, v in sd.items()}, x.numpy())
        assert np.abs(got - ref.numpy()).max() < 1e-4 * max(1.0, float(ref.abs().max()))
        prob, nms_map = WO.heatmap_nms(got)
        assert np.abs(prob - bx.numpy()).max() < 1e-5
        assert np.array_equal(nms_map != 0, ref_map != 0)
        assert np.abs(nms_map - ref_map).max() < 1e-5
        assert np.array_equal(WO.attention_mask(), utils.get_attention_mask(12, 1).numpy().reshape(12, 12))
    # the literal quirks of the post-processing on a hand-made map: float row coordinate of the box, circular class axis, exhausted map -> index 0
    lg = np.full((1, 120, 12), -30.0, dtype=np.float32)
    lg[0, 40, 3] = 5.0
    lg[0, 44, 9] = 4.0   # inside the box of the first peak in y (|44 - 40.25| <= 5) and in x through the wrap (|9 - 3 - 12| = 6 <= 7)
    t = torch.from_numpy(lg)
    bx = torch.softmax(t.reshape(1, -1), 1).reshape(1, 120, 12)
    wrap = torch.cat((bx[:, -1:], bx, bx[:, :1]), 1)
    ref_map = utils.nms(wrap.unsqueeze(1), max_predictions=5, sigma=(7.0, 5.0)).squeeze(1)[:, 1:-1, :].numpy()
    _, nms_map = WO.heatmap_nms(lg)
    assert np.array_equal(nms_map != 0, ref_map != 0) and np.abs(nms_map - ref_map).max() < 1e-7
