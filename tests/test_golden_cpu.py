"""CPU suite: the oracle restatement vs golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py).
These run everywhere (no GPU, no /root/reference)."""
import glob
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_geometry_vs_reference_vectors():
    from oracle import geometry as G
    z = np.load(os.path.join(GOLD, "geometry.npz"))
    for i, h in enumerate(z["headings"]):
        rx, ry, rz, d, s = G.unproject_habitat(z["depth"][i], float(h))
        assert np.array_equal(np.stack([rx, ry, rz, d, s]), z[f"unproj{i}"])
    assert np.array_equal(np.stack(G.patch_3d_info(z["depth"])), z["info5"])
    n = len(z["cull_pts"])
    want = np.unpackbits(z["cull_mask"])[:n].astype(bool)
    got = G.frustum_mask_habitat(z["cull_pts"], z["cull_depth"].astype(np.float32), z["cull_cam"], float(z["cull_heading"][0]))
    assert np.array_equal(got, want) and 100 < want.sum() < n


def test_posed_geometry_vs_reference_vectors():
    """a4': oracle/geometry.py posed-dataset functions against the vectors the reference's own functions produced (posed.npz)."""
    from oracle import geometry as G
    z = np.load(os.path.join(GOLD, "posed.npz"))
    for i in range(3):
        xyz, d, s = G.unproject_posed_view(z[f"depth{i}"], z[f"K{i}"], z[f"R{i}"], z[f"T{i}"], 1000.0, 1000.0, ray_fx=140.0)
        assert np.array_equal(xyz, z[f"xyz{i}"]) and np.array_equal(d, z[f"dir{i}"]) and np.array_equal(s, z[f"scale{i}"])
    n = len(z["cull_pts"])
    want = np.unpackbits(z["cull_mask"])[:n].astype(bool)
    got = G.frustum_mask_matrix(z["cull_pts"], (z["depth1"].astype(np.float32) / 1000.0).astype(np.float32), z["K1"].astype(np.float32), z["cull_M"])
    assert (got != want).sum() <= 1 and want.sum() > 100  # torch's CPU einsum may contract the 4-term sums (one-ulp boundary cases)


def test_vit_restatement_vs_reference_vectors():
    from dynam3d_b200 import synth
    from oracle import nn_ops as NN
    z = np.load(os.path.join(GOLD, "vit_small.npz"))
    sd = synth.vit_state_dict(9, width=256, layers=3, resolution=112, out_dim=128)
    x = synth.hash_uniform((2, 3, 112, 112), 77, 2.0)
    for pre in (False, True):
        c, p = NN.vit_forward(x, sd, 3, 4, rnd=None, ln_post_on_patches=not pre)
        assert np.abs(c.numpy() - z["cls_pre" if pre else "cls"]).max() < 2e-5
        assert np.abs(p.numpy() - z["patch_pre" if pre else "patch"]).max() < 2e-5


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ff_traj_*.npz"))), ids=os.path.basename)
def test_feature_fields_oracle_vs_reference_trajectory(path):
    from oracle import ref_compare as RC
    from oracle.make_golden import load_ff_fixture
    cfg, steps, gold = load_ff_fixture(path)
    _, orc, recs = RC.run_pair(**cfg, steps=steps, with_reference=False)
    for t, (r, g) in enumerate(zip(recs, gold)):
        assert RC.snapshots_equal(g["snap"], r["orc"]) == [], f"step {t}"
        e = r["env_orc"]
        assert np.allclose(e["batch_instance_relative_position"][0], g["inst_rel"], atol=2e-5, equal_nan=True)
        assert np.allclose(e["batch_zone_relative_position"][0], g["zone_rel"], atol=2e-5, equal_nan=True)
        assert np.allclose(e["batch_instance_fts"][0].sum(-1), g["inst_fts_sum"], atol=2e-3)
        assert np.allclose(e["batch_zone_fts"][0].sum(-1), g["zone_fts_sum"], atol=2e-3)
        if "knn_idx" in g:
            assert np.array_equal(r["knn"][1], g["knn_idx"]) and np.array_equal(r["merge"][0].astype(np.uint8), g["merge"])


def test_render_oracle_vs_reference_renderer_output():
    """a18: oracle/render_oracle.py vs the stored output of the reference's own render_view_3d_patch (tests/golden/render.npz)."""
    from oracle import nn_ops as NN
    from oracle import render_oracle as RO
    from oracle.make_golden import render_scene
    z = np.load(os.path.join(GOLD, "render.npz"))
    xyz, dr, sc, fts, pos, head, P = render_scene()
    want = RO.render_view_3d_patch(P, xyz, dr, sc, fts, pos, head, rnd=NN.round_fp16)
    valid = (want["idx"] >= 0).any(-1).any(-1)
    assert valid.sum() > 100
    assert np.array_equal(z["positions"][valid], want["positions"][valid])
    assert np.abs(z["feature_map"].astype(np.float32) - want["feature_map"]).max() < 1e-3  # stored as fp16 (the reference's output dtype)


def test_waypoint_oracle_vs_reference_vectors():
    """8(f) rank 3: oracle/waypoint_oracle.py vs the stored outputs of the reference's own predictor and NMS (tests/golden/waypoint.npz;
    inputs regenerated from the seeds of oracle/make_golden.py)."""
    from dynam3d_b200 import synth
    from oracle import waypoint_oracle as WO
    from oracle.make_golden import WAYPOINT_SEEDS
    g = np.load(os.path.join(GOLD, "waypoint.npz"))
    sd = {k: v.numpy() for k, v in synth.waypoint_state_dict(WAYPOINT_SEEDS[0]).items()}
    x = synth.waypoint_depth_embedding(WAYPOINT_SEEDS[1], WAYPOINT_SEEDS[2]).numpy()
    lg = WO.predictor_logits(sd, x)
    assert np.abs(lg - g["logits"]).max() < 1e-4 * max(1.0, float(np.abs(g["logits"]).max()))
    _, nms_map = WO.heatmap_nms(lg)
    assert np.array_equal(nms_map != 0, g["nms"] != 0) and np.abs(nms_map - g["nms"]).max() < 1e-5
    c = WO.candidates_from_map(nms_map[0])
    assert len(c["cand_angles"]) == 5 and all(0.25 <= d <= 3.0 for d in c["cand_distances"]) and all(0 <= i < 12 for i in c["cand_img_idxes"])
