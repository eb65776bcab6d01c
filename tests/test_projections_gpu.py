"""a14 in isolation (POL:83-111, 432-435): the five policy projection MLPs (Linear -> LayerNorm -> GELU -> Linear) on the C ABI vs the oracle's
`mlp_ln_gelu` with the same rounding points, and in the precise (split-operand) arithmetic vs the pure-fp32 oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precise", [False, True])
def test_policy_projections_match_oracle(precise):
    from dynam3d_b200 import ops, synth
    from dynam3d_b200.policy import Dynam3D_VLN
    from oracle import nn_ops as NN
    sd = synth.policy_state_dict(21)
    net = Dynam3D_VLN(precise=precise)
    net.load_policy_state_dict(sd)
    PW = net._policy_weights()
    P = {k: v.float() for k, v in sd.items()}
    rnd = None if precise else NN.round_fp16
    rng = np.random.default_rng(5)
    n = 203
    fts = torch.from_numpy(rng.standard_normal((n, 768)).astype(np.float32))
    rel = torch.from_numpy((rng.standard_normal((n, 3)) * 2.0).astype(np.float32))
    for pos_name, proj_name, pos_key, proj_key in (("instance_position_embedding", "instance_projector", "inst_pos", "inst_proj"),
                                                   ("zone_position_embedding", "zone_projector", "zone_pos", "zone_proj")):
        got = net._project_tokens(fts.cuda(), rel.cuda(), PW[pos_key], PW[proj_key]).cpu()
        pe = NN.mlp_ln_gelu(rel, P, pos_name, rnd)
        want = NN.mlp_ln_gelu(torch.cat([fts, pe], -1), P, proj_name, rnd)
        e = (got - want).abs().max().item()
        print(f"{proj_name} ({'precise' if precise else 'production'}): max abs err {e:.2e} (|y| max {want.abs().max().item():.2f})")
        assert got.shape == (n, 3072) and e < (1e-4 if precise else 3e-3)
    # patch_position_embedding on the 6-d patch info rows [rel_x, rel_y, rel_z, sin(dir), cos(dir), scale]
    info5 = torch.from_numpy(rng.standard_normal((5, 2, 576)).astype(np.float32)).cuda()
    rows = torch.empty((2 * 576, 8), device="cuda", dtype=torch.float32 if precise else torch.float16)
    ops.patch_info_rows(info5.contiguous(), rows)
    got = net._mlp(rows, PW["patch_pos"]).cpu()
    i5 = info5.cpu()
    feat6 = torch.stack([i5[0], i5[1], i5[2], torch.sin(i5[3]), torch.cos(i5[3]), i5[4]], -1).reshape(-1, 6)
    want = NN.mlp_ln_gelu(feat6, P, "patch_position_embedding", rnd)
    e = (got - want).abs().max().item()
    print(f"patch_position_embedding ({'precise' if precise else 'production'}): max abs err {e:.2e}")
    assert e < (1e-4 if precise else 3e-3)
    assert net._project_tokens(torch.zeros((0, 768), device="cuda"), torch.zeros((0, 3), device="cuda"), PW["inst_pos"], PW["inst_proj"]).shape == (0, 3072)
