"""Bit-exact parity of the geometry / K-NN kernels against the CPU oracle (oracle/geometry.py), through the C ABI."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("q1_fix", [False, True])
@pytest.mark.parametrize("views,size", [(1, 256), (12, 256), (3, 224)])
def test_depth_grid_and_full(views, size, q1_fix):
    from dynam3d_b200 import ops, synth
    from oracle import geometry as G
    B = 2
    rng = np.random.default_rng(5)
    obs = np.concatenate([st["depth"] for b in range(B) for st in synth.make_episode(10 + b, 1, views, depth_size=size)], 0)
    obs[0, :, 7, 0] = 0.0  # an all-zero column: max fill leaves 0
    want = G.depth_patch_grid(obs, B, views, q1_fix=q1_fix).reshape(B * views, 576)
    got = ops.depth_patch_grid(_dev(obs[..., 0]), B, views, literal_q1=not q1_fix).cpu().numpy()
    assert np.array_equal(want, got)
    want_full = G.preprocess_depth(obs, (0.0, 10.0))[..., 0]
    got_full = ops.depth_preprocess(_dev(obs[..., 0])).cpu().numpy()
    assert np.array_equal(want_full, got_full)


@pytest.mark.parametrize("hfov", [90.0, 79.0])
def test_unproject_bit_exact(hfov):
    from dynam3d_b200 import ops
    from oracle import geometry as G
    rng = np.random.default_rng(3)
    n, V = 4, 12
    depth = rng.uniform(0.05, 10.0, size=(n * V, 576)).astype(np.float32)
    pos = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
    head = rng.uniform(0, 2 * math.pi, size=n).tolist()
    pose = ops.pose_rows(pos, head, V)
    xyz, d, s = ops.unproject_habitat(_dev(depth), _dev(pose), hfov, hfov)
    xyz, d, s = xyz.cpu().numpy(), d.cpu().numpy(), s.cpu().numpy()
    for b in range(n):
        for ix in range(V):
            u = b * V + ix
            w_xyz, w_d, w_s = G.unproject_view_world(depth[u], pos[b], head[b], ix, hfov, hfov)
            assert np.array_equal(w_xyz, xyz[u]) and np.array_equal(w_d, d[u]) and np.array_equal(w_s, s[u])
    info = ops.patch_3d_info(_dev(depth), hfov, hfov).cpu().numpy()
    want = G.patch_3d_info(depth, hfov, hfov)
    for i in range(5):
        assert np.array_equal(want[i], info[i])


def test_frustum_cull_bit_exact():
    from dynam3d_b200 import ops, synth
    from oracle import geometry as G
    rng = np.random.default_rng(11)
    steps = synth.make_episode(21, n_steps=1, num_views=12)
    st = steps[0]
    depth_m = G.preprocess_depth(st["depth"], (0.0, 10.0))[..., 0]  # [12,256,256]
    N = 200000
    pts = rng.uniform(-5, 5, size=(N, 3)).astype(np.float32)
    pts[:, 2] = rng.uniform(0, 3, size=N).astype(np.float32)
    pts[:50] = -10000.0
    cam_int = G.habitat_to_internal(st["position"])
    pts[50] = cam_int.astype(np.float32)  # a point exactly at the camera: 0/0
    heads = [st["heading"] + ix * (-math.pi / 6) for ix in range(12)]
    want = np.zeros(N, bool)
    for ix in range(12):
        want |= G.frustum_mask_habitat(np.where(want[:, None], np.float32(-10000.0), pts), depth_m[ix], cam_int, heads[ix])
    xyz = _dev(pts.copy()); dr = torch.ones(N, device="cuda"); sc = torch.ones(N, device="cuda")
    fts = torch.ones((N, 768), device="cuda", dtype=torch.float16)
    cam = _dev(ops.camera_rows(st["position"], heads))
    mask, n_del = ops.frustum_cull(xyz, dr, sc, fts, N, _dev(depth_m), cam)
    mask = mask.cpu().numpy().astype(bool)
    assert np.array_equal(mask, want) and int(n_del.item()) == int(want.sum()) and want.sum() > 100
    assert torch.all(xyz[torch.from_numpy(want).cuda()] == -10000.0)
    assert torch.all(fts[torch.from_numpy(want).cuda()] == 0) and torch.all(fts[torch.from_numpy(~want).cuda()] == 1)
    assert float(dr.sum().item()) == float((~want).sum())


@pytest.mark.parametrize("n_ref,n_q,k", [(3, 5, 2), (577, 16, 2), (4096, 60, 2), (9216, 72144, 4), (5000, 1000, 8), (40, 300000, 1)])
def test_knn_bit_exact(n_ref, n_q, k):
    from dynam3d_b200 import ops
    from oracle import geometry as G
    rng = np.random.default_rng(n_ref + n_q)
    refs = rng.uniform(-8, 8, size=(n_ref, 3)).astype(np.float32)
    refs[::7] = -10000.0  # coincident tombstones: exact ties, lowest index must win
    q = rng.uniform(-8, 8, size=(n_q, 3)).astype(np.float32)
    q[: min(n_q, 3)] = refs[: min(n_q, 3)][: min(n_q, 3)] if n_ref >= 3 else q[:3]
    d2, idx = ops.knn3d(_dev(refs), _dev(q), k)
    sub = slice(0, min(n_q, 6000))
    wd, wi = G.knn3d(refs, q[sub], k)
    assert np.array_equal(idx.cpu().numpy()[sub], wi)
    assert np.array_equal(d2.cpu().numpy()[sub], wd)
    # size-independent properties on the full result: ascending, indices valid, distances consistent with indices
    d2c, idxc = d2.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    assert np.all(np.diff(d2c, axis=1) >= 0) and idxc.min() >= 0 and idxc.max() < n_ref
    chk = rng.integers(0, n_q, size=min(n_q, 2000))
    diff = q[chk][:, None, :] - refs[idxc[chk]]
    rec = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]).astype(np.float32) + (diff[..., 2] * diff[..., 2]).astype(np.float32)
    assert np.array_equal(rec.astype(np.float32), d2c[chk])


def test_centroid_and_export():
    from dynam3d_b200 import ops
    from oracle import geometry as G
    from oracle.ff_oracle import mean_f64
    rng = np.random.default_rng(2)
    xyz = rng.uniform(-50, 50, size=(5000, 3)).astype(np.float32)
    lens = [1, 0, 577, 33, 2000, 7]
    member = rng.integers(0, 5000, size=sum(lens)).astype(np.int32)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    got = ops.seq_centroid(_dev(xyz), _dev(member), _dev(cu), len(lens)).cpu().numpy()
    for s, n in enumerate(lens):
        want = mean_f64(xyz[member[cu[s]:cu[s + 1]]])
        assert np.array_equal(want, got[s], equal_nan=True)
    # export
    fts = rng.standard_normal((5000, 768)).astype(np.float32)
    ids = rng.permutation(5000)[:3000].astype(np.int32)
    pos_h = np.array([1.0, 1.25, -2.0], np.float32); heading = 1.234
    rel, dist = G.to_agent_frame(xyz[ids], pos_h, heading)
    keep = dist <= np.float32(30.0)
    out_rel = torch.empty((3000, 3), device="cuda"); out_fts = torch.empty((3000, 768), device="cuda")
    cnt = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.env_export(_dev(xyz), _dev(fts), _dev(ids), _dev(ops.camera_rows(pos_h, [heading])[0]), 30.0, out_rel, out_fts, cnt)
    n = int(cnt.item())
    assert n == int(keep.sum()) and 100 < n < 3000
    assert np.array_equal(out_rel[:n].cpu().numpy(), rel[keep]) and np.array_equal(out_fts[:n].cpu().numpy(), fts[ids][keep])


@pytest.mark.parametrize("size,M", [(336, 40), (224, 7), (576, 150)])
def test_segm_relabel_bit_exact(size, M):
    """a7 (FF:411-422): FastSAM masks -> dense 24x24 labels, vs the reference's own torch ops restated in oracle/geometry.py."""
    from dynam3d_b200 import ops
    from oracle import geometry as G
    rng = np.random.default_rng(size + M)
    masks = np.zeros((2, M, size, size), np.uint8)
    for i in range(2):
        for g in range(M):
            y0, x0 = rng.integers(0, size - 20, 2)
            h, w = rng.integers(10, size // 2, 2)
            masks[i, g, y0:y0 + h, x0:x0 + w] = 1
    lab, n_seg = ops.segm_relabel(_dev(masks))
    for i in range(2):
        want = G.segm_relabel(masks[i])
        assert np.array_equal(lab[i].cpu().numpy(), want)
        assert int(n_seg[i].item()) == int(want.max()) + 1
