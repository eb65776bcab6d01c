"""a4' -- the posed-dataset branch (open3d unprojection + pose transform + matrix frustum cull) on the B200 vs the CPU oracle:
bit-exact geometry, identical discrete 3D-memory state on a posed multi-view trajectory."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def posed_view(seed, H=240, W=320):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    depth = (1500 + 900 * np.sin(xx / 37.0 + seed) + 600 * np.cos(yy / 23.0 + 0.3 * seed) + rng.integers(0, 40, (H, W))).astype(np.uint16)
    depth[rng.integers(0, H, 30), rng.integers(0, W, 30)] = 0
    K = np.array([[285.0 + seed % 7, 0, W / 2 - 3.5], [0, 291.0, H / 2 + 2.25], [0, 0, 1]], np.float64)
    a = 0.35 * seed + 0.2
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.0]])
    T = np.array([[0.3 * np.cos(seed)], [0.3 * np.sin(seed)], [1.2]]) + rng.uniform(-0.05, 0.05, (3, 1))
    M = np.eye(4)
    M[:3, :3], M[:3, 3:] = R.T, -R.T @ T
    return depth, K, R, T, M.astype(np.float32)


@pytest.mark.parametrize("shape", [(240, 320), (480, 640), (24, 24)])
def test_unproject_pinhole_bit_exact(shape):
    from dynam3d_b200 import ops
    from oracle import geometry as G
    views = [posed_view(s, *shape) for s in range(5)]
    vp = np.zeros((len(views), 16))
    for i, (d, K, R, T, _) in enumerate(views):
        vp[i] = [K[0][0], K[1][1], K[0][2], K[1][2], *R.reshape(9), *T.reshape(3)]
    fx0 = views[0][1][0][0]
    t = abs(np.tan(ops.ray_direction0(fx0, 24, 3.0)))
    depth = torch.from_numpy(np.stack([v[0] for v in views]).view(np.int16)).cuda()
    xyz, direction, scale, bad = ops.unproject_pinhole(depth, torch.from_numpy(vp).cuda(), 1000.0, 1000.0, t)
    assert int(bad.item()) == 0
    for i, (d, K, R, T, _) in enumerate(views):
        x, dr, sc = G.unproject_posed_view(d, K, R, T, 1000.0, 1000.0, ray_fx=fx0)
        assert np.array_equal(xyz[i].cpu().numpy(), x), i
        assert np.array_equal(scale[i].cpu().numpy(), sc), i
        assert np.array_equal(direction[i].cpu().numpy(), dr), i  # fp64 asin rounded to fp32
    # depth beyond depth_trunc: open3d drops the pixel, the reference's view() raises -> counted
    _, _, _, bad = ops.unproject_pinhole(depth, torch.from_numpy(vp).cuda(), 1000.0, 2.0, t)
    assert int(bad.item()) > 0


def test_frustum_cull_matrix_bit_exact():
    from dynam3d_b200 import ops
    from oracle import geometry as G
    rng = np.random.default_rng(1)
    pts = rng.uniform(-3, 3, (60000, 3)).astype(np.float32)
    views = [posed_view(s) for s in (0, 3, 4)]
    depth_m = np.stack([(v[0].astype(np.float32) / 1000.0).astype(np.float32) for v in views])
    cam25 = np.stack([np.concatenate([v[4].reshape(16), v[1].astype(np.float32)[:3, :3].reshape(9)]) for v in views]).astype(np.float32)
    want = np.zeros(len(pts), bool)
    for i, v in enumerate(views):
        want |= G.frustum_mask_matrix(pts, depth_m[i], v[1].astype(np.float32), v[4])
    xyz = torch.from_numpy(pts.copy()).cuda()
    dr, sc = torch.ones(len(pts), device="cuda"), torch.ones(len(pts), device="cuda")
    fts = torch.ones((len(pts), 768), device="cuda", dtype=torch.float16)
    mask, n = ops.frustum_cull_matrix(xyz, dr, sc, fts, len(pts), torch.from_numpy(depth_m).cuda(), torch.from_numpy(cam25).cuda())
    mask = mask.cpu().numpy().astype(bool)
    assert want.sum() > 200 and np.array_equal(mask, want) and int(n.item()) == int(want.sum())
    assert torch.all(xyz[torch.from_numpy(want).cuda()] == -10000.0) and float(fts[torch.from_numpy(want).cuda()].abs().max()) == 0.0
    assert float(fts[torch.from_numpy(~want).cuda()].min()) == 1.0


def test_posed_trajectory_matches_oracle():
    """Feature_Fields driven through the dataset-branch signature (FF:329,493) for 3 steps of 4 posed views, 2 episodes."""
    from test_feature_fields_gpu import _params, _compare_snap
    from dynam3d_b200 import synth
    from dynam3d_b200.feature_fields import Feature_Fields
    from oracle import nn_ops as NN
    from oracle.ff_oracle import FeatureFieldsOracle
    B, V, n_steps = 2, 4, 3
    sd = _params(11, merge_bias=0.0)
    eng = Feature_Fields(batch_size=B)
    eng.load_state_dict(sd, strict=True)
    orc = FeatureFieldsOracle(sd, batch_size=B, rnd=NN.round_fp16)
    rng = np.random.default_rng(11)
    for t in range(n_steps):
        views = [[posed_view(100 * b + 4 * t + ix) for ix in range(V)] for b in range(B)]
        depth = [np.stack([v[0] for v in views[b]]) for b in range(B)]
        Ks = [[v[1] for v in views[b]] for b in range(B)]
        Rs = [[v[2] for v in views[b]] for b in range(B)]
        Ts = [[v[3] for v in views[b]] for b in range(B)]
        Ms = [[v[4] for v in views[b]] for b in range(B)]
        K32 = [[k.astype(np.float32) for k in Ks[b]] for b in range(B)]
        depth_m = [(depth[b].astype(np.float32) / 1000.0).astype(np.float32) for b in range(B)]
        segm = np.stack([np.stack([synth.make_segmentation(np.random.default_rng(1000 * t + 10 * ix + b), 16, "voronoi") for ix in range(V)]) for b in range(B)])
        grid = (rng.standard_normal((B, V, 576, 768)) * 0.5).astype(np.float16)
        orc.delete_old_features_posed(depth_m, K32, Ms)
        orc.update_feature_fields_posed(depth, grid, segm, Ks, Rs, Ts)
        eng.delete_old_features_from_camera_frustum(depth_m, batch_camera_intrinsic=K32, batch_extrinsic=Ms)
        eng.update_feature_fields(depth, grid, None, batch_camera_intrinsic=Ks, batch_rot=Rs, batch_trans=Ts, depth_scale=1000.0,
                                  depth_trunc=1000.0, batch_patch_segm=segm)
        for b in range(B):
            assert _compare_snap(orc.snapshot(b), eng.snapshot(b)) == [], f"step {t} episode {b}"
            assert np.array_equal(eng.global_patch_position[b].cpu().numpy(), orc.eps[b].patch_pos)
        if t > 0:
            assert any(s["patch_tomb"].any() for s in (eng.snapshot(b) for b in range(B))), "fixture culls nothing"
