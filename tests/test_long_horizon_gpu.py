"""BASELINE.json configs[4]: long-horizon rollout -- 256 one-view steps, the dynamic token memory growing to ~4 k instance slots and
147 k stored patches on one GPU -- against oracle/ff_oracle.py (pinned to the reference class): identical discrete state at steps
{1, 32, 128, 256}, bit-equal exported agent-frame positions, and every term that grows with the memory exercised (frustum cull over
all stored patches FF:329-396, K-NN over all instance slots FF:604-610, re-encode of all member patches of merged instances FF:662-688,
episode pools growing past their initial commitment without `reserve()`)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _drive(ff, orc, steps, n_seg, seed, checkpoints, B=1):
    """Steps engine and oracle side by side (FF only: hash-generated CLIP grid features); compares at `checkpoints` (1-based steps)."""
    from dynam3d_b200 import ops, synth
    from oracle import geometry as G
    from oracle.ref_compare import snapshots_equal
    eps = [synth.make_episode(seed + b, n_steps=steps, num_views=1, rgb_size=8, depth_size=256, n_seg=n_seg, seg_kind="voronoi") for b in range(B)]
    n_merge = 0
    for t in range(steps):
        depth = np.concatenate([eps[b][t]["depth"] for b in range(B)], 0)
        segm = np.stack([eps[b][t]["segm"] for b in range(B)], 0)
        pos, head = [eps[b][t]["position"] for b in range(B)], [eps[b][t]["heading"] for b in range(B)]
        grid = synth.hash_uniform((B, 1, 576, 768), seed * 1000 + t, 0.9).numpy().astype(np.float16)
        # oracle (CPU)
        d576 = G.depth_patch_grid(depth, B, 1, q1_fix=True)
        full = G.preprocess_depth(depth, (0.0, 10.0)).reshape(B, 1, 256, 256)
        with torch.no_grad():
            orc.delete_old_features_from_camera_frustum(full, pos, head, num_of_views=1)
            orc.update_feature_fields(d576, grid, segm, pos, head, num_of_views=1)
        # engine
        dd = torch.from_numpy(depth).cuda().reshape(B, 256, 256).contiguous()
        e576 = ops.depth_patch_grid(dd, B, 1, 24, 24, literal_q1=False)
        efull = ops.depth_preprocess(dd, 0.0, 10.0).view(B, 1, 256, 256)
        ff.delete_old_features_from_camera_frustum(efull, pos, head, num_of_views=1)
        ff.update_feature_fields(e576.view(B, 1, 576), torch.from_numpy(grid).cuda(), batch_position=pos, batch_heading=head, num_of_views=1,
                                 batch_patch_segm=segm)
        last = ff._last(0)
        if last:
            n_merge += int(last["merge"].any(axis=1).sum())
        if (t + 1) in checkpoints:
            want = orc.get_environment_features(pos, head)
            got = ff.get_environment_features(pos, head)
            for b in range(B):
                assert snapshots_equal(orc.snapshot(b), ff.snapshot(b)) == [], f"state differs at step {t + 1}"
                for key in ("batch_instance_relative_position", "batch_zone_relative_position"):
                    w, g = np.asarray(want[key][b], np.float32).reshape(-1, 3), got[key][b].cpu().numpy()
                    assert w.shape == g.shape and np.array_equal(w, g, equal_nan=True), f"{key} differs at step {t + 1}"
                wf, gf = np.asarray(want["batch_instance_fts"][b], np.float32), got["batch_instance_fts"][b].cpu().numpy()
                assert wf.shape == gf.shape
                if len(wf):
                    assert np.abs(wf - gf).max() < 5e-2  # fp16-operand pooled encoder vs the matched oracle (values ~ +-3)
    return n_merge


def _pair(merge_bias, B=1):
    from dynam3d_b200 import synth
    from dynam3d_b200.feature_fields import Feature_Fields
    from oracle import nn_ops as NN
    from oracle.ff_oracle import FeatureFieldsOracle
    pol = synth.policy_state_dict(7, merge_bias=merge_bias)
    ff_sd = {k[len("feature_fields."):]: v for k, v in pol.items() if k.startswith("feature_fields.")}
    ff = Feature_Fields(batch_size=B, device="cuda", q7_fix=True)
    ff.load_state_dict(ff_sd)
    orc = FeatureFieldsOracle(ff_sd, batch_size=B, rnd=NN.round_fp16, q7_fix=True)
    return ff, orc


def test_256_steps_to_4k_instances_identical_state():
    """Merges off (every segment founds an instance): 256 steps x 17 segments -> >= 4 000 instance slots, 147 456 stored patches.
    No reserve(): the pools outgrow their initial commitment (16 384 patches, 2 048 instances) several times inside the rollout."""
    ff, orc = _pair(merge_bias=-10.0)
    ff.reset(1)
    base0 = ff.eps[0].patch_fts.t.data_ptr()
    _drive(ff, orc, 256, 17, 5000, checkpoints={1, 32, 128, 256})
    ep = ff.eps[0]
    assert ep.n_patch == 256 * 576 and ep.n_inst >= 4000, (ep.n_patch, ep.n_inst)
    assert ep.patch_fts.t.data_ptr() == base0 and ep.patch_fts.t.shape[0] >= ep.n_patch  # grown in place: same base address


def test_merge_heavy_rollout_identical_state():
    """Discriminator biased towards merging: instances accumulate hundreds of member patches that are re-encoded at every merge."""
    ff, orc = _pair(merge_bias=0.6)
    ff.reset(1)
    n_merge = _drive(ff, orc, 96, 16, 5100, checkpoints={1, 32, 64, 96})
    assert n_merge > 100, n_merge
    sizes = [len(v) for v in ff.global_instance_to_patch_dict[0].values()]
    assert max(sizes) > 150, max(sizes)


def test_rollout_delete_reset_rollout_uses_fresh_pools():
    """The trainer's sequence reset -> steps -> delete_feature_fields -> reset -> steps (ss_trainer_Dynam3D.py:621/683/803): the view
    runtime's pool-address table must follow the new episodes' pools (advisor finding, round 1)."""
    ff, orc = _pair(merge_bias=0.3, B=2)
    ff.reset(2)
    _drive(ff, orc, 40, 16, 5200, checkpoints={40}, B=2)  # 40 x 576 = 23 040 patches > the initial 16 384: the pools grew
    ff.delete_feature_fields()
    ff.reset(2)
    orc.reset(2)
    _drive(ff, orc, 6, 16, 5300, checkpoints={1, 6}, B=2)
    ff.pop(0)
    orc.pop(0)
    assert ff.batch_size == 1 and ff.snapshot(0)["n_patches"] == 6 * 576
