"""SURVEY.md 8(f) rank 3 -- candidate-waypoint predictor (BinaryDistPredictor_TRM + heat-map NMS, POL:188-292) on the GPU vs the CPU oracle
(pinned against the reference's own classes) and vs the reference-generated vectors of tests/golden/waypoint.npz."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _predictor(seed):
    from dynam3d_b200 import synth
    from dynam3d_b200.waypoint import WaypointPredictor
    p = WaypointPredictor("cuda")
    p.load_state_dict(synth.waypoint_state_dict(seed), strict=False)
    return p


def test_predictor_matches_reference_vectors_and_oracle():
    from dynam3d_b200 import synth
    from oracle import waypoint_oracle as WO
    from oracle.make_golden import WAYPOINT_SEEDS
    g = np.load(os.path.join(GOLD, "waypoint.npz"))
    p = _predictor(WAYPOINT_SEEDS[0])
    x = synth.waypoint_depth_embedding(WAYPOINT_SEEDS[1], WAYPOINT_SEEDS[2])
    lg = p(x.cuda())
    err = float((lg.cpu() - torch.from_numpy(g["logits"])).abs().max())
    print(f"waypoint logits vs reference vectors: max abs err {err:.2e} on a range of {float(np.abs(g['logits']).max()):.1f}")
    assert err < 5e-5 * float(np.abs(g["logits"]).max())   # fp32-class (2-term activations x 3-term weights on the tensor cores): ~2e-5 relative
    prob, nms = p.heatmap_nms(lg)
    assert np.array_equal(nms.cpu().numpy() != 0, g["nms"] != 0)              # the reference's candidate cells, exactly
    assert float((nms.cpu() - torch.from_numpy(g["nms"])).abs().max()) < 1e-5
    # and the derived per-episode lists against the oracle's restatement of POL:253-270
    cands = p.candidates(lg)
    for b in range(lg.shape[0]):
        want = WO.candidates_from_map(g["nms"][b])
        assert np.array_equal(cands[b]["angle_idxes"], want["angle_idxes"]) and np.array_equal(cands[b]["distance_idxes"], want["distance_idxes"])
        assert cands[b]["cand_angles"] == want["cand_angles"] and cands[b]["cand_distances"] == want["cand_distances"]
        assert np.array_equal(cands[b]["cand_img_idxes"], want["cand_img_idxes"]) and np.array_equal(cands[b]["cand_angle_fts"], want["cand_angle_fts"])


@pytest.mark.parametrize("seed,episodes", [(8, 1), (9, 8)])
def test_predictor_vs_oracle_other_seeds(seed, episodes):
    from dynam3d_b200 import synth
    from oracle import waypoint_oracle as WO
    sd = synth.waypoint_state_dict(seed)
    p = _predictor(seed)
    x = synth.waypoint_depth_embedding(seed + 1, episodes)
    lg = p(x.cuda())
    want = WO.predictor_logits({k: v.numpy() for k, v in sd.items()}, x.numpy())
    assert float(np.abs(lg.cpu().numpy() - want).max()) < 5e-5 * max(1.0, float(np.abs(want).max()))
    _, nms = p.heatmap_nms(lg)
    _, want_map = WO.heatmap_nms(want)
    assert np.array_equal(nms.cpu().numpy() != 0, want_map != 0)


def test_heatmap_nms_kernel_literal_quirks():
    """d3d_wp_heatmap_nms on hand-made and random maps vs the oracle (itself checked against the reference's `nms`): float row coordinate of the
    suppression box, circular CLASS axis, wrap rows, an exhausted map returning index 0, ties -> first maximum."""
    from dynam3d_b200.waypoint import WaypointPredictor
    from oracle import waypoint_oracle as WO
    p = WaypointPredictor("cuda")
    rng = np.random.default_rng(0)
    maps = [rng.normal(0, 3, (4, 120, 12)).astype(np.float32)]
    lg = np.full((3, 120, 12), -30.0, dtype=np.float32)
    lg[0, 40, 3], lg[0, 44, 9] = 5.0, 4.0            # second peak suppressed through the class-axis wrap
    lg[1, 0, 0], lg[1, 119, 11], lg[1, 60, 6] = 6.0, 6.0, 6.0   # exact ties; peaks on the wrap rows
    lg[2, 10, 5] = 40.0                              # one peak holds all the mass: the map is exhausted after round 1
    maps.append(lg)
    for m in maps:
        prob, nms = p.heatmap_nms(torch.from_numpy(m).cuda())
        wprob, wmap = WO.heatmap_nms(m)
        assert float(np.abs(prob.cpu().numpy() - wprob).max()) < 1e-6
        assert np.array_equal(nms.cpu().numpy() != 0, wmap != 0)
        assert float(np.abs(nms.cpu().numpy() - wmap).max()) < 1e-6


def test_neighbor_attention_vs_torch():
    import ctypes, math
    from dynam3d_b200 import _lib as L
    from dynam3d_b200.waypoint import neighbor_mask
    B, H, Dh = 3, 12, 64
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * 12, 3 * H * Dh, generator=g).cuda()
    mask = torch.from_numpy(neighbor_mask()).cuda()
    out = torch.empty(B * 12, H * Dh, device="cuda")
    L.check(L.lib().d3d_wp_neighbor_attention(L.ptr(qkv), L.ptr(mask), B, 12, H, Dh, ctypes.c_float(1.0 / math.sqrt(Dh)), L.ptr(out), L.stream_ptr()))
    q, k, v = (t.view(B, 12, H, Dh).transpose(1, 2) for t in qkv.split(H * Dh, -1))
    sc = q @ k.transpose(-1, -2) / math.sqrt(Dh) + mask
    want = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * 12, H * Dh)
    assert float((out - want).abs().max()) < 1e-5


def test_get_candidate_waypoints_reference_signature():
    """`Dynam3D_VLN.get_candidate_waypoints(waypoint_predictor, observations)` (POL:188-292) with a stand-in depth encoder attached: view
    re-ordering, flip back, pooling and the candidate lists vs a direct restatement with the oracle."""
    from dynam3d_b200 import synth
    from dynam3d_b200.policy import Dynam3D_VLN
    from oracle import waypoint_oracle as WO
    from oracle.make_golden import WAYPOINT_SEEDS
    B = 2
    net = Dynam3D_VLN()
    proj = synth.hash_uniform((128 * 16, 64), 77, 0.3).cuda()

    def depth_encoder(obs):  # stand-in for VlnResnetDepthEncoder: [B*12, 16, 16, 1] -> [B*12, 128, 4, 4], any deterministic map will do
        d = obs["depth"].cuda().float().reshape(obs["depth"].shape[0], -1)[:, :64]
        return torch.relu(d @ proj.T).reshape(-1, 128, 4, 4)

    net.depth_encoder = depth_encoder
    obs = {("depth" if i == 0 else f"depth_{i}"): synth.hash_uniform((B, 16, 16, 1), 100 + i, 1.0) + 1.0 for i in range(12)}
    p = _predictor(WAYPOINT_SEEDS[0])
    out = net.get_candidate_waypoints(p, obs)
    # restatement: clockwise order for the predictor
    dep = torch.zeros(B * 12, 16, 16, 1)
    for a, k in enumerate(obs):
        for bi in range(B):
            dep[(12 - a) % 12 + bi * 12] = obs[k][bi]
    emb = depth_encoder({"depth": dep})
    sd = {k: v.numpy() for k, v in synth.waypoint_state_dict(WAYPOINT_SEEDS[0]).items()}
    _, want_map = WO.heatmap_nms(WO.predictor_logits(sd, emb.cpu().numpy()))
    e5 = emb.reshape(B, 12, 128, 4, 4)
    feats = torch.cat((e5[:, 0:1], torch.flip(e5[:, 1:], [1])), 1).mean(dim=(3, 4))
    assert torch.allclose(out["pano_depth"], feats, atol=1e-6) and out["pano_angle_fts"].shape == (12, 4)
    for j in range(B):
        w = WO.candidates_from_map(want_map[j])
        assert out["cand_angles"][j] == w["cand_angles"] and out["cand_distances"][j] == w["cand_distances"]
        assert np.array_equal(out["cand_img_idxes"][j], w["cand_img_idxes"])
        assert torch.allclose(out["cand_depth"][j], feats[j, torch.from_numpy(w["cand_img_idxes"]).cuda()], atol=1e-6)
        assert np.array_equal(out["cand_angle_fts"][j].numpy(), w["cand_angle_fts"])


def test_waypoint_timing_print():
    """Not an assertion on speed: prints the device time of predictor + NMS for 8 episodes (the bench's batch) for DESIGN.md."""
    from dynam3d_b200 import synth
    p = _predictor(3)
    x = synth.waypoint_depth_embedding(4, 8).cuda()
    for _ in range(3):
        lg = p(x)
        p.heatmap_nms(lg)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        lg = p(x)
        p.heatmap_nms(lg)
    b.record(); torch.cuda.synchronize()
    print(f"waypoint predictor + heat-map NMS, 8 episodes x 12 views: {a.elapsed_time(b) / 10:.3f} ms per call")
