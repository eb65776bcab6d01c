"""Pretrain novel-view patch renderer (PFF:494-625) on the C ABI vs oracle/render_oracle.py: sample selection, K-NN indices and
ray points bit-exact; rendered features floating point (fp16 GEMM operands both sides)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_render_view_3d_patch_matches_oracle():
    from dynam3d_b200 import ops, synth
    from dynam3d_b200.pretrain_render import NerfRenderer
    from oracle import geometry as G
    from oracle import nn_ops as NN
    from oracle import render_oracle as RO
    # a small scene: patches unprojected from 4 views of a synthetic room
    ep = synth.make_episode(31, n_steps=1, num_views=4, n_seg=16)[0]
    d576 = G.depth_patch_grid(ep["depth"], 1, 4, q1_fix=True)[0]
    xyz, dr, sc = [], [], []
    for ix in range(4):
        a, b, c = G.unproject_view_world(d576[ix], ep["position"], ep["heading"], ix)
        xyz.append(a); dr.append(b); sc.append(c)
    xyz, dr, sc = np.concatenate(xyz), np.concatenate(dr), np.concatenate(sc)
    xyz[::97] = -10000.0  # a few tombstones
    fts = synth.hash_uniform((len(xyz), 768), 5, 0.9).numpy().astype(np.float16)
    P = synth.nerf_state_dict(3)
    want = RO.render_view_3d_patch(P, xyz, dr, sc, fts, ep["position"], ep["heading"] + 0.3, rnd=NN.round_fp16)
    ren = NerfRenderer(P)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    fmap, pos, depth, aux = ren.render(dev(xyz), dev(dr), dev(sc), dev(fts), ep["position"], ep["heading"] + 0.3)
    assert np.array_equal(aux["ray_xyz"].cpu().numpy(), want["ray_xyz"]), "ray sample points (fp64 expression rounded once)"
    assert np.array_equal(aux["topk"].cpu().numpy(), want["topk"]), "important samples per ray"
    assert np.array_equal(aux["idx"].cpu().numpy().astype(np.int64), want["idx"]), "K-NN gather indices"
    assert np.array_equal(pos.cpu().numpy(), want["positions"])
    valid = (want["idx"] >= 0).any(-1).any(-1)
    assert 10 < valid.sum() <= 144
    e_x = np.abs(aux["pos_rows"].float().cpu().numpy()[:, :6] - want["xyzds"]).max()
    e_s = np.abs(aux["si"].cpu().numpy() - want["si"]).max()
    print(f"renderer stages: geometry rows err {e_x:.2e} (fp16 storage), aggregated sample input err {e_s:.2e}")
    assert e_x <= 2e-2 and e_s <= 2e-2
    e_d = np.abs(aux["density"].cpu().numpy() - want["density"]).max()
    e_f = np.abs(fmap.cpu().numpy() - want["feature_map"]).max()
    e_z = np.abs(depth.cpu().numpy() - want["depth_map"]).max()
    print(f"renderer: density err {e_d:.2e}, unit-norm feature err {e_f:.2e}, depth err {e_z:.2e}")
    # fp16 layer outputs (tinycudann semantics) both sides; accumulation order differs -> a few fp16 ulps through 5 GEMM layers
    assert e_d <= 2e-2 and e_f <= 5e-3 and e_z <= 5e-2
