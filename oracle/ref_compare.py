"""TEST INFRASTRUCTURE ONLY -- runs the unmodified reference `Feature_Fields` (via ref_shim, CPU) and the
restatement (`ff_oracle.FeatureFieldsOracle`) side by side on a seeded trajectory and compares state.

Only usable where /root/reference exists (this container).  Used by tests/test_oracle_vs_reference.py
and oracle/make_golden.py.
"""
import numpy as np
import torch

from dynam3d_b200 import synth
from . import geometry as G
from . import ref_shim
from .ff_oracle import FeatureFieldsOracle


def tune_discriminator(params, merge_bias=0.0):
    """Random init never merges; bias the 'merge' logit so a usable fraction of proposals merge."""
    params = dict(params)
    b = params["instance_merge_discriminator.3.bias"].clone()
    b[1] = b[1] + merge_bias
    params["instance_merge_discriminator.3.bias"] = b
    return params


def reference_snapshot(ff, b=0):
    return {
        "n_patches": len(ff.global_patch_position[b]),
        "p2i": dict(ff.global_patch_to_instance_dict[b]),
        "i2p": {k: np.asarray(v).copy() for k, v in ff.global_instance_to_patch_dict[b].items()},
        "i2p_order": list(ff.global_instance_to_patch_dict[b].keys()),
        "n_inst_slots": len(ff.global_instance_position[b]),
        "zone_key_to_id": dict(ff.global_zone_key_to_id[b]),
        "z2i": {k: np.asarray(v).copy() for k, v in ff.global_zone_to_instance_dict[b].items()},
        "z2i_order": list(ff.global_zone_to_instance_dict[b].keys()),
        "n_zone_slots": len(ff.global_zone_position[b]),
        "patch_tomb": (np.asarray(ff.global_patch_position[b])[:, 0] == -10000.0).copy(),
    }


def snapshots_equal(a, b):
    problems = []
    for k in ("n_patches", "n_inst_slots", "n_zone_slots", "i2p_order", "z2i_order"):
        if a[k] != b[k]:
            problems.append(f"{k}: {a[k]} != {b[k]}")
    if a["p2i"] != b["p2i"]:
        problems.append("p2i differs")
    for name in ("i2p", "z2i"):
        if set(a[name]) != set(b[name]):
            problems.append(f"{name} keys differ")
        else:
            for k in a[name]:
                if not np.array_equal(np.asarray(a[name][k]), np.asarray(b[name][k])):
                    problems.append(f"{name}[{k}] differs")
    ka = {tuple(float(x) for x in k): v for k, v in a["zone_key_to_id"].items()}
    kb = {tuple(float(x) for x in k): v for k, v in b["zone_key_to_id"].items()}
    if ka != kb:
        problems.append("zone_key_to_id differs")
    if not np.array_equal(a["patch_tomb"], b["patch_tomb"]):
        problems.append("patch tombstones differ")
    return problems


def make_inputs(seed, n_steps, num_views, n_seg, seg_kind, depth_size=256):
    """Deterministic step inputs: synthetic episode + hash-generated CLIP grid features (fp16)."""
    steps = synth.make_episode(seed, n_steps=n_steps, num_views=num_views, n_seg=n_seg, seg_kind=seg_kind, depth_size=depth_size)
    for t, st in enumerate(steps):
        st["grid"] = synth.hash_uniform((1, num_views, 576, 768), seed * 100 + t, 0.9).numpy().astype(np.float16)
    return steps


def ff_params(weight_seed, merge_bias):
    sd = synth.policy_state_dict(weight_seed, merge_bias)
    return {k[len("feature_fields."):]: v for k, v in sd.items() if k.startswith("feature_fields.")}


def run_pair(seed=1, n_steps=4, num_views=1, n_seg=16, seg_kind="blocks", merge_bias=0.0, q1_fix=False, weight_seed=0, steps=None,
             with_reference=True):
    """Returns (reference ff, oracle, per-step records).  Each record: snapshots + exported env tokens of both."""
    params = ff_params(weight_seed, merge_bias)
    ff = None
    if with_reference:
        ff = ref_shim.make_reference_feature_fields(batch_size=1, seed=weight_seed)
        ff.load_state_dict(params, strict=True)
        ff.reset(1)
        ff.initialize_camera_setting(90.0, 90.0)
    orc = FeatureFieldsOracle(params, batch_size=1, rnd=None)
    orc.initialize_camera_setting(90.0, 90.0)
    if steps is None:
        steps = make_inputs(seed, n_steps, num_views, n_seg, seg_kind)
    records = []
    for st in steps:
        V = num_views
        grid = st["grid"]
        obs_depth = st["depth"]  # [V,H,W,1]
        d576 = G.depth_patch_grid(obs_depth, 1, V, q1_fix=q1_fix)  # [1,V,576]
        full = G.preprocess_depth(obs_depth, (0.0, 10.0)).reshape(1, V, obs_depth.shape[1], obs_depth.shape[2])
        pos = [np.asarray(st["position"], np.float32)]
        head = [float(st["heading"])]
        rec = {}
        if with_reference:
            segm = torch.from_numpy(np.asarray(st["segm"])).view(V, 1, 24, 24)
            ref_shim.attach_segmentation(ff, lambda img, s=segm: s)
            with torch.no_grad():
                ff.delete_old_features_from_camera_frustum(torch.from_numpy(full), pos, head, num_of_views=V)
                ff.update_feature_fields(d576.copy(), grid.astype(np.float32), batch_image=np.zeros((1, V, 4, 4, 3), np.uint8),
                                         batch_position=pos, batch_heading=head, num_of_views=V)
                rec["env_ref"] = ff.get_environment_features(pos, head)
            rec["ref"] = reference_snapshot(ff)
        orc.delete_old_features_from_camera_frustum(full, pos, head, num_of_views=V)
        orc.update_feature_fields(d576, grid, np.asarray(st["segm"])[None], pos, head, num_of_views=V)
        rec.update({"orc": orc.snapshot(), "env_orc": orc.get_environment_features(pos, head), "knn": orc.eps[0].last_knn,
                    "merge": getattr(orc.eps[0], "last_merge", None)})
        records.append(rec)
    return ff, orc, records
