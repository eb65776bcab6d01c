"""CPU test of the C++ bookkeeping planner (csrc/ff_host.cu) through the C ABI: driven with the ORACLE's per-view numeric results
(centroids, K-NN proposals, merge logits) it must reproduce the oracle's / reference's discrete state on the golden trajectories."""
import ctypes
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _planner_snapshot(lib, h, b, n_patch):
    from dynam3d_b200.feature_fields import Feature_Fields
    ff = Feature_Fields.__new__(Feature_Fields)  # only the read-back helpers are used
    ff._h = h
    ff.eps = [type("E", (), {"n_patch": n_patch})()]
    cnt = np.zeros(8, np.int64)
    lib.d3d_ffh_counts(h, b, cnt.ctypes.data)
    i2p, z2i, p2i = ff._map(b, 0), ff._map(b, 1), ff._p2i(b)
    pos = np.zeros((max(n_patch, 1), 3), np.float32)
    lib.d3d_ffh_get_patch_pos(h, b, pos.ctypes.data)
    for which, m in ((0, i2p), (1, z2i)):  # the light accessor used by the export returns the same dict-order keys
        ids = np.zeros(max(int(cnt[2 if which == 0 else 4]), 1), np.int64); n = np.zeros(1, np.int64)
        assert lib.d3d_ffh_live_ids(h, b, which, ids.ctypes.data, n.ctypes.data) == 0
        assert ids[: int(n[0])].tolist() == list(m.keys())
    ff._h = None
    return {"n_patches": int(cnt[0]), "p2i": {int(k): int(p2i[k]) for k in np.flatnonzero(p2i >= 0)}, "i2p": i2p, "i2p_order": list(i2p.keys()),
            "n_inst_slots": int(cnt[2]), "zone_key_to_id": Feature_Fields._zone_keys_dict(type("F", (), {"_h": h})(), b), "z2i": z2i,
            "z2i_order": list(z2i.keys()), "n_zone_slots": int(cnt[4]), "patch_tomb": (pos[:n_patch, 0] == -10000.0).copy()}, cnt


def _drive_step_planner(cfg, steps, gold):
    """Runs the oracle over `steps` and replays its per-view numeric results into the whole-step planner entry points
    (d3d_ffh_begin_step + d3d_ffh_begin_view_refs + d3d_ffh_finish_view); the planner's discrete state must equal the oracle's (and the
    golden reference state when given) after every step."""
    import copy
    from dynam3d_b200 import _lib as L
    from dynam3d_b200.feature_fields import Feature_Fields
    from oracle import geometry as G
    from oracle import ref_compare as RC
    from oracle.ff_oracle import FeatureFieldsOracle
    lib = L.lib()
    V, P = cfg["num_views"], 576
    orc = FeatureFieldsOracle(RC.ff_params(cfg["weight_seed"], cfg["merge_bias"]), batch_size=1, rnd=None)
    h = lib.d3d_ffh_create(1, 2, 2.0)
    try:
        for t, (st, g) in enumerate(zip(steps, gold if gold is not None else [None] * len(steps))):
            d576 = G.depth_patch_grid(st["depth"], 1, V, q1_fix=cfg.get("q1_fix", False))
            full = G.preprocess_depth(st["depth"], (0.0, 10.0)).reshape(1, V, 256, 256)
            pos, head = [st["position"]], [st["heading"]]
            ep = orc.eps[0]
            n0 = len(ep.patch_pos)
            before = ep.patch_pos[:, 0].copy() == -10000.0 if n0 else np.zeros(0, bool)
            orc.delete_old_features_from_camera_frustum(full, pos, head, num_of_views=V)
            if n0:
                mask = ((ep.patch_pos[:, 0] == -10000.0) & ~before).astype(np.uint8)
                di, dz = np.zeros(4096, np.int64), np.zeros(4096, np.int64)
                nd = (ctypes.c_int * 2)()
                L.check(lib.d3d_ffh_cull(h, 0, mask.ctypes.data, n0, di.ctypes.data, ctypes.addressof(nd), dz.ctypes.data, ctypes.addressof(nd) + 4))
            else:
                lib.d3d_ffh_set_tree(h)
            per_view = []
            for ix in range(V):  # the oracle runs the whole step first; its per-view results are replayed into the planner below
                orc._update_view(ep, np.asarray(d576[0][ix], np.float32), st["grid"][0][ix].astype(np.float16), st["segm"][ix].reshape(-1),
                                 st["position"], float(st["heading"]), ix)
                per_view.append((copy.deepcopy(ep.last_raw), copy.deepcopy(ep.last_knn), copy.deepcopy(ep.last_merge)))
            xyz = np.ascontiguousarray(ep.patch_pos[-V * P:].reshape(V, 1, P, 3)).astype(np.float32)
            segm = np.ascontiguousarray(np.asarray(st["segm"]).reshape(V, 1, P), dtype=np.int64)
            cap = V * P
            base_rows = np.zeros(1, np.int64); view_start = np.zeros(V + 1, np.int32); seq_owner = np.zeros(cap, np.int32)
            members = np.zeros(cap, np.int32); cu_m = np.zeros(cap + 1, np.int32); tok_src = np.zeros(2 * cap, np.int32)
            tok_seq = np.zeros(2 * cap, np.int32); cu_tok = np.zeros(cap + 1, np.int32); info = np.zeros(2, np.int32)
            L.check(lib.d3d_ffh_begin_step(h, xyz.ctypes.data, segm.ctypes.data, P, V, base_rows.ctypes.data, view_start.ctypes.data,
                                           seq_owner.ctypes.data, members.ctypes.data, cu_m.ctypes.data, tok_src.ctypes.data, tok_seq.ctypes.data,
                                           cu_tok.ctypes.data, info.ctypes.data))
            assert int(info[0]) == sum(len(r[0]["centres"]) for r in per_view) and int(base_rows[0]) == n0
            # stage rows of view ix are [ix*P, (ix+1)*P) for a single episode: every member list must stay inside its view's block
            for ix in range(V):
                m = members[cu_m[view_start[ix]]:cu_m[view_start[ix + 1]]]
                assert m.min() >= ix * P and m.max() < (ix + 1) * P and len(m) == P
            for ix in range(V):
                raw, last_knn, last_merge = per_view[ix]
                Gn = len(raw["centres"])
                assert int(view_start[ix + 1] - view_start[ix]) == Gn
                n_ref = np.zeros(Gn, np.int32)
                L.check(lib.d3d_ffh_begin_view_refs(h, ix, n_ref.ctypes.data))
                res = np.zeros((Gn, 12), np.float32)
                res[:, 0:3] = raw["centres"]
                k_raw = raw["d2"].shape[1]
                res[:, 3:3 + k_raw] = raw["d2"]
                idx2 = np.full((Gn, 2), -1, np.int32); idx2[:, :k_raw] = raw["idx"]
                res[:, 5:7] = idx2.view(np.float32)
                if last_knn is not None and last_merge[1].shape[1] > 0:
                    lg = last_merge[1]
                    pad = np.zeros((Gn, 2, 2), np.float32); pad[:, :lg.shape[1]] = lg
                    res[:, 7:11] = pad.reshape(Gn, 4)
                sizes = np.zeros(10, np.int32); after = np.zeros(3, np.int64)
                L.check(lib.d3d_ffh_finish_view(h, res.ctypes.data, sizes.ctypes.data, after.ctypes.data))
                if last_knn is not None:
                    last = Feature_Fields._last(type("F", (), {"_h": h})(), 0)
                    assert np.array_equal(last["knn"][1], last_knn[1]) and np.array_equal(last["merge"], last_merge[0].astype(bool)), (t, ix)
            snap, cnt = _planner_snapshot(lib, h, 0, len(ep.patch_pos))
            if g is not None:
                assert RC.snapshots_equal(g["snap"], snap) == [], f"step {t}"
            assert RC.snapshots_equal(orc.snapshot(0), snap) == [], f"step {t} (oracle)"
    finally:
        lib.d3d_ffh_destroy(h)




@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ff_traj_*.npz"))), ids=os.path.basename)
def test_step_planner_reproduces_golden_state(path):
    """The whole-step entry points the engine uses, on the golden trajectories produced by the reference."""
    from oracle.make_golden import load_ff_fixture
    cfg, steps, gold = load_ff_fixture(path)
    _drive_step_planner(cfg, steps, gold)


@pytest.mark.parametrize("cfg", [
    dict(seed=21, n_steps=5, num_views=2, n_seg=48, seg_kind="voronoi", merge_bias=2.0, weight_seed=3),   # merge-heavy: slot reuse after culls
    dict(seed=22, n_steps=4, num_views=3, n_seg=16, seg_kind="voronoi", merge_bias=0.3, weight_seed=1, q1_fix=True),
    dict(seed=23, n_steps=6, num_views=1, n_seg=36, seg_kind="blocks", merge_bias=-2.0, weight_seed=2),   # never merges: instance slots only grow / die
], ids=lambda c: f"seed{c['seed']}")
def test_step_planner_matches_oracle_on_fresh_trajectories(cfg):
    """Beyond the three golden fixtures: seeded trajectories with other view counts, segmentations and merge rates (oracle vs planner)."""
    from oracle import ref_compare as RC
    steps = RC.make_inputs(cfg["seed"], cfg["n_steps"], cfg["num_views"], cfg["n_seg"], cfg["seg_kind"])
    _drive_step_planner(cfg, steps, None)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "ff_traj_*.npz"))), ids=os.path.basename)
def test_planner_reproduces_golden_state(path):
    from dynam3d_b200 import _lib as L
    from oracle import geometry as G
    from oracle import ref_compare as RC
    from oracle.ff_oracle import FeatureFieldsOracle
    from oracle.make_golden import load_ff_fixture
    lib = L.lib()
    cfg, steps, gold = load_ff_fixture(path)
    V, P = cfg["num_views"], 576
    orc = FeatureFieldsOracle(RC.ff_params(cfg["weight_seed"], cfg["merge_bias"]), batch_size=1, rnd=None)
    h = lib.d3d_ffh_create(1, 2, 2.0)
    try:
        for t, (st, g) in enumerate(zip(steps, gold)):
            d576 = G.depth_patch_grid(st["depth"], 1, V, q1_fix=cfg.get("q1_fix", False))
            full = G.preprocess_depth(st["depth"], (0.0, 10.0)).reshape(1, V, 256, 256)
            pos, head = [st["position"]], [st["heading"]]
            ep = orc.eps[0]
            # ---- cull: feed the oracle's mask (union over views == sequential application) ----
            n0 = len(ep.patch_pos)
            before = ep.patch_pos[:, 0].copy() == -10000.0 if n0 else np.zeros(0, bool)
            orc.delete_old_features_from_camera_frustum(full, pos, head, num_of_views=V)
            if n0:
                mask = ((ep.patch_pos[:, 0] == -10000.0) & ~before).astype(np.uint8)
                di, dz = np.zeros(4096, np.int64), np.zeros(4096, np.int64)
                nd = (ctypes.c_int * 2)()
                L.check(lib.d3d_ffh_cull(h, 0, mask.ctypes.data, n0, di.ctypes.data, ctypes.addressof(nd), dz.ctypes.data, ctypes.addressof(nd) + 4))
            else:
                lib.d3d_ffh_set_tree(h)
            # ---- update, view by view ----
            for ix in range(V):
                orc._update_view(ep, np.asarray(d576[0][ix], np.float32), st["grid"][0][ix].astype(np.float16), st["segm"][ix].reshape(-1),
                                 st["position"], float(st["heading"]), ix)
                raw = ep.last_raw
                Gn = len(raw["centres"])
                xyz = np.ascontiguousarray(ep.patch_pos[-P:]).astype(np.float32)
                # the oracle appended the view before culling nothing in between: the last P rows are this view's patches
                segm = np.ascontiguousarray(st["segm"][ix].reshape(1, P), dtype=np.int64)
                cap = P
                base_rows = np.zeros(1, np.int64); n_seg = np.zeros(1, np.int32); seq_owner = np.zeros(cap, np.int32)
                members = np.zeros(cap, np.int32); cu_m = np.zeros(cap + 1, np.int32); tok_src = np.zeros(2 * cap, np.int32)
                tok_seq = np.zeros(2 * cap, np.int32); cu_tok = np.zeros(cap + 1, np.int32); n_ref = np.zeros(cap, np.int32); info = np.zeros(2, np.int32)
                stage_off = np.zeros(1, np.int64)
                L.check(lib.d3d_ffh_begin_view(h, xyz.ctypes.data, segm.ctypes.data, P, stage_off.ctypes.data, base_rows.ctypes.data, n_seg.ctypes.data,
                                               seq_owner.ctypes.data, members.ctypes.data, cu_m.ctypes.data, tok_src.ctypes.data, tok_seq.ctypes.data,
                                               cu_tok.ctypes.data, n_ref.ctypes.data, info.ctypes.data))
                assert int(info[0]) == Gn
                res = np.zeros((Gn, 12), np.float32)
                res[:, 0:3] = raw["centres"]
                k_raw = raw["d2"].shape[1]
                res[:, 3:3 + k_raw] = raw["d2"]
                idx2 = np.full((Gn, 2), -1, np.int32); idx2[:, :k_raw] = raw["idx"]
                res[:, 5:7] = idx2.view(np.float32)
                if ep.last_knn is not None and ep.last_merge[1].shape[1] > 0:
                    lg = ep.last_merge[1]  # [G, K', 2]
                    pad = np.zeros((Gn, 2, 2), np.float32); pad[:, :lg.shape[1]] = lg
                    res[:, 7:11] = pad.reshape(Gn, 4)
                sizes = np.zeros(10, np.int32); after = np.zeros(3, np.int64)
                L.check(lib.d3d_ffh_finish_view(h, res.ctypes.data, sizes.ctypes.data, after.ctypes.data))
                if ep.last_knn is not None:
                    from dynam3d_b200.feature_fields import Feature_Fields
                    last = Feature_Fields._last(type("F", (), {"_h": h})(), 0)
                    assert np.array_equal(last["knn"][1], ep.last_knn[1]) and np.array_equal(last["knn"][0], ep.last_knn[0]), (t, ix)
                    assert np.array_equal(last["merge"], ep.last_merge[0].astype(bool)), (t, ix)
            snap, cnt = _planner_snapshot(lib, h, 0, len(ep.patch_pos))
            assert RC.snapshots_equal(g["snap"], snap) == [], f"step {t}"
            assert RC.snapshots_equal(orc.snapshot(0), snap) == [], f"step {t} (oracle)"
    finally:
        lib.d3d_ffh_destroy(h)
