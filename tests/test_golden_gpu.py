"""The CUDA engine vs golden vectors produced by the UNMODIFIED reference in the build container (tests/golden/, made by
oracle/make_golden.py): discrete 3D-memory state per step must be identical to the reference's; token features are
floating point (reference = fp32 CPU, engine = fp16 GEMM operands like the reference's autocast path)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ff_traj_*.npz"))), ids=os.path.basename)
def test_engine_reproduces_reference_trajectory(path):
    from dynam3d_b200 import ops
    from dynam3d_b200.feature_fields import Feature_Fields
    from oracle import ref_compare as RC
    from oracle.make_golden import load_ff_fixture
    cfg, steps, gold = load_ff_fixture(path)
    V = cfg["num_views"]
    eng = Feature_Fields(batch_size=1)
    eng.load_state_dict(RC.ff_params(cfg["weight_seed"], cfg["merge_bias"]), strict=True)
    for t, (st, g) in enumerate(zip(steps, gold)):
        obs = torch.from_numpy(st["depth"][..., 0]).cuda()
        d576 = ops.depth_patch_grid(obs, 1, V, literal_q1=not cfg.get("q1_fix", False)).view(1, V, 576)
        full = ops.depth_preprocess(obs).view(1, V, obs.shape[1], obs.shape[2])
        pos, head = [st["position"]], [st["heading"]]
        eng.delete_old_features_from_camera_frustum(full, pos, head, num_of_views=V)
        eng.update_feature_fields(d576, torch.from_numpy(st["grid"]).cuda(), batch_position=pos, batch_heading=head, num_of_views=V,
                                  batch_patch_segm=st["segm"][None])
        assert RC.snapshots_equal(g["snap"], eng.snapshot(0)) == [], f"step {t}"
        if "knn_idx" in g:
            assert np.array_equal(eng._last(0)["knn"][1], g["knn_idx"]), "K-NN indices"
            assert np.array_equal(eng._last(0)["knn"][0], g["knn_d2"]), "K-NN squared distances"
            assert np.array_equal(eng._last(0)["merge"].astype(np.uint8), g["merge"]), "merge decisions"
        env = eng.get_environment_features(pos, head)
        assert np.allclose(env["batch_instance_relative_position"][0].cpu().numpy(), g["inst_rel"], atol=2e-5, equal_nan=True)
        assert np.allclose(env["batch_zone_relative_position"][0].cpu().numpy(), g["zone_rel"], atol=2e-5, equal_nan=True)
        # per-token feature checksums (sum over 768 LayerNorm-ed channels): fp16-operand noise ~1e-3 per channel, random sign
        assert np.allclose(env["batch_instance_fts"][0].sum(-1).cpu().numpy(), g["inst_fts_sum"], atol=0.15)
        assert np.allclose(env["batch_zone_fts"][0].sum(-1).cpu().numpy(), g["zone_fts_sum"], atol=0.15)


def test_geometry_kernels_vs_reference_vectors():
    from dynam3d_b200 import ops
    z = np.load(os.path.join(GOLD, "geometry.npz"))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    depth = z["depth"]
    pose = ops.pose_rows([np.zeros(3, np.float32)] * 4, z["headings"].tolist(), 1)
    xyz, d, s = ops.unproject_habitat(dev(depth), dev(pose))
    for i in range(4):
        w = z[f"unproj{i}"]
        assert np.array_equal(xyz[i].cpu().numpy(), w[:3].T) and np.array_equal(d[i].cpu().numpy(), w[3]) and np.array_equal(s[i].cpu().numpy(), w[4])
    assert np.array_equal(ops.patch_3d_info(dev(depth)).cpu().numpy(), z["info5"])
    n = len(z["cull_pts"])
    want = np.unpackbits(z["cull_mask"])[:n].astype(bool)
    cam = dev(ops.camera_rows(np.array([z["cull_cam"][0], z["cull_cam"][2], -z["cull_cam"][1]], np.float32), [float(z["cull_heading"][0])]))
    mask, _ = ops.frustum_cull(dev(z["cull_pts"].copy()), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), None, n,
                               dev(z["cull_depth"].astype(np.float32)[None]), cam)
    assert np.array_equal(mask.cpu().numpy().astype(bool), want)


def test_posed_kernels_vs_reference_vectors():
    """a4': d3d_unproject_pinhole / d3d_frustum_cull_matrix against the vectors the reference's own functions produced (posed.npz)."""
    from dynam3d_b200 import ops
    z = np.load(os.path.join(GOLD, "posed.npz"))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    vp = np.zeros((3, 16))
    for i in range(3):
        K = z[f"K{i}"]
        vp[i] = [K[0][0], K[1][1], K[0][2], K[1][2], *z[f"R{i}"].reshape(9), *z[f"T{i}"].reshape(3)]
    depth = np.stack([z[f"depth{i}"] for i in range(3)])
    t = abs(np.tan(ops.ray_direction0(140.0, 24, 3.0)))
    xyz, d, s, bad = ops.unproject_pinhole(dev(depth.view(np.int16)), dev(vp), 1000.0, 1000.0, t)
    assert int(bad.item()) == 0
    for i in range(3):
        assert np.array_equal(xyz[i].cpu().numpy(), z[f"xyz{i}"]) and np.array_equal(d[i].cpu().numpy(), z[f"dir{i}"])
        assert np.array_equal(s[i].cpu().numpy(), z[f"scale{i}"])
    n = len(z["cull_pts"])
    want = np.unpackbits(z["cull_mask"])[:n].astype(bool)
    cam25 = np.concatenate([z["cull_M"].reshape(16), z["K1"].astype(np.float32)[:3, :3].reshape(9)]).astype(np.float32)[None]
    mask, _ = ops.frustum_cull_matrix(dev(z["cull_pts"].copy()), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), None, n,
                                      dev((depth[1].astype(np.float32) / 1000.0).astype(np.float32)[None]), dev(cam25))
    assert (mask.cpu().numpy().astype(bool) != want).sum() <= 1  # torch's CPU einsum may contract the 4-term sums


def test_renderer_vs_reference_renderer_output():
    """a18 on the C ABI vs the stored output of the reference's own render_view_3d_patch (PFF:494-625; tests/golden/render.npz)."""
    import os
    from dynam3d_b200.pretrain_render import NerfRenderer
    from oracle.make_golden import render_scene
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render.npz"))
    xyz, dr, sc, fts, pos, head, P = render_scene()
    ren = NerfRenderer(P)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    fmap, p, depth, aux = ren.render(dev(xyz), dev(dr), dev(sc), dev(fts), pos, head)
    valid = (aux["idx"].cpu().numpy() >= 0).any(-1).any(-1)
    assert valid.sum() > 100
    assert np.array_equal(p.cpu().numpy()[valid], z["positions"][valid])  # selected samples: bit-equal wherever the ray has a neighbour
    e = np.abs(fmap.cpu().numpy() - z["feature_map"].astype(np.float32)).max()
    print(f"renderer vs the reference's own output: unit-norm feature err {e:.2e}")
    assert e <= 5e-3
