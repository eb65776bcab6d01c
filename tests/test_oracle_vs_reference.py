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
