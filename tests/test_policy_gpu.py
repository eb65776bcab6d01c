"""Full navigation step (a1-a16) on the C ABI vs the CPU oracle restatement of Dynam3D_VLN.forward (oracle/policy_oracle.py).

Discrete outputs (token counts, 3D-memory state, argmax action token) must be identical.  Logits are floating point: both sides
round GEMM operands to fp16 at the same points (the reference runs under fp16 autocast); the residual difference is the
16-bit rounding-flip noise floor measured in DESIGN.md (a 1e-7 relative perturbation of the oracle itself moves the logits by
~4e-3 at 4 layers), so the tolerance here is 8e-3 absolute on logits of magnitude ~4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(seed, clip_layers, lm_layers, B, q1_fix):
    from dynam3d_b200 import synth
    from dynam3d_b200.policy import Dynam3D_VLN
    from oracle import nn_ops as NN
    from oracle.policy_oracle import PolicyOracle
    pol_sd = synth.policy_state_dict(seed, merge_bias=0.0)
    clip_sd = synth.vit_state_dict(seed, layers=clip_layers)
    llava_sd = synth.llava_state_dict(seed, clip_layers=clip_layers, lm_layers=lm_layers, lm_round_to=torch.float16)
    net = Dynam3D_VLN(q1_fix=q1_fix)
    net.load_policy_state_dict(pol_sd)
    net.rgb_encoder.load_openai_state_dict(clip_sd)
    net.llava.load_state_dict(llava_sd, max_images=B)
    net.feature_fields.reset(B)
    orc = PolicyOracle(pol_sd, clip_sd, llava_sd, clip_layers=clip_layers, lm_layers=lm_layers, batch_size=B, rnd=NN.round_fp16, q1_fix=q1_fix)
    return net, orc


@pytest.mark.parametrize("cfg", [dict(seed=2, V=1, B=1, steps=2, q1_fix=False), dict(seed=4, V=3, B=2, steps=2, q1_fix=True)])
def test_full_step_matches_oracle(cfg):
    from dynam3d_b200 import synth
    from oracle.ref_compare import snapshots_equal
    B, V = cfg["B"], cfg["V"]
    net, orc = _build(cfg["seed"], 2, 2, B, cfg["q1_fix"])
    tok = synth.ToyTokenizer()
    eps = [synth.make_episode(cfg["seed"] * 10 + b, n_steps=cfg["steps"], num_views=V, rgb_size=224, n_seg=16, seg_kind="voronoi") for b in range(B)]
    instr = [synth.make_instruction(cfg["seed"] + b) for b in range(B)]
    for t in range(cfg["steps"]):
        obs = {"rgb": np.concatenate([eps[b][t]["rgb"] for b in range(B)], 0), "depth": np.concatenate([eps[b][t]["depth"] for b in range(B)], 0),
               "patch_segm": np.stack([eps[b][t]["segm"] for b in range(B)], 0)}
        pos = [eps[b][t]["position"] for b in range(B)]
        head = [eps[b][t]["heading"] for b in range(B)]
        # the prompt needs the token counts, which depend on the 3D memory: run the engine's encoder stage first
        t_obs = {"rgb": torch.from_numpy(obs["rgb"]), "depth": torch.from_numpy(obs["depth"]), "patch_segm": obs["patch_segm"]}
        patch, inst, zone = net.encode_step(t_obs, pos, head, num_of_views=V)
        ids = [tok(net.build_prompt(576 + inst[b].shape[0] + zone[b].shape[0], instr[b], ["none\n"] * 4)) for b in range(B)]
        want = orc.step_logits(obs, pos, head, ids, num_of_views=V)
        # engine LM on the tokens it just built (same code path as forward_logits after encode_step)
        lm = net.llava.lm
        seqs, lens = [], []
        for b in range(B):
            n_img = 576 + inst[b].shape[0] + zone[b].shape[0]
            idt = torch.tensor(ids[b], dtype=torch.int32, device="cuda")
            eh = torch.empty((2, 3072), device="cuda"); et = torch.empty((len(ids[b]) - n_img - 2, 3072), device="cuda")
            lm.embed(idt[:2].contiguous(), eh); lm.embed(idt[n_img + 2:].contiguous(), et)
            seqs += [eh, patch[b], inst[b], zone[b], et]
            lens.append(2 + n_img + et.shape[0])
        assert lens == orc.last_lens
        X = torch.cat(seqs, 0)
        cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
        p = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
        got = lm.prefill(X, cu, p, B, max(lens), (cu[1:] - 1).int().contiguous()).cpu()
        for b in range(B):
            assert snapshots_equal(orc.ff.snapshot(b), net.feature_fields.snapshot(b)) == []
        err = (got - want).abs().max().item()
        print(f"step {t}: S={lens} |logit|max={want.abs().max().item():.2f} err={err:.2e}")
        assert err <= 8e-3
        assert torch.equal(got.argmax(-1), want.argmax(-1))


def test_forward_eval_branch_returns_text_and_updates_history():
    """POL:463-469 through the public `forward`: tokenizer in, greedy decode with the KV cache, text cut at "<|end|>", history rotated."""
    from dynam3d_b200 import synth
    B, V = 2, 1
    net, _ = _build(6, 2, 2, B, q1_fix=True)
    net.tokenize = synth.ToyTokenizer()
    vocab = {32007: "<|end|>"}
    net.detokenize = lambda ids: "".join(vocab.get(i, chr(97 + i % 26)) for i in ids)
    net.max_new_tokens = 6
    # two independent history lists (the reference aliases them, Q10; the test wants to see both episodes)
    net.feature_fields.history_actions = [["none\n"] * 4 for _ in range(B)]
    eps = [synth.make_episode(60 + b, n_steps=1, num_views=V, rgb_size=224, n_seg=16, seg_kind="voronoi") for b in range(B)]
    obs = {"rgb": torch.from_numpy(np.concatenate([eps[b][0]["rgb"] for b in range(B)], 0)),
           "depth": torch.from_numpy(np.concatenate([eps[b][0]["depth"] for b in range(B)], 0)),
           "patch_segm": np.stack([eps[b][0]["segm"] for b in range(B)], 0)}
    pos, head = [eps[b][0]["position"] for b in range(B)], [eps[b][0]["heading"] for b in range(B)]
    instr = [synth.make_instruction(b) for b in range(B)]
    # reference run of the same step for the ids (fresh memory), then the text path
    logits, ids = net.generate_ids(obs, instr, pos, head, num_of_views=V)
    assert logits.shape == (B, 32064) and all(1 <= len(s) <= 6 for s in ids)
    assert [s[0] for s in ids] == logits.argmax(-1).cpu().tolist()  # the first generated id is the prefill's arg-max
    net.feature_fields.reset(B)
    net.feature_fields.history_actions = [["none\n"] * 4 for _ in range(B)]
    net.eos_token_ids = (ids[0][2],)  # make the third token of episode 0 the end token
    vocab[ids[0][2]] = "<|end|>"
    texts = net(obs, instr, pos, head, num_of_views=V)
    assert isinstance(texts, list) and len(texts) == B and all(isinstance(t, str) for t in texts)
    k = ids[0].index(ids[0][2])  # greedy decoding of random weights may repeat ids: the text ends at the FIRST occurrence
    assert len(texts[0]) == k and "<|end|>" not in texts[0]
    for b in range(B):
        h = net.feature_fields.history_actions[b]
        assert len(h) == 4 and h[-1] == texts[b] + "\n" and h[0] == "none\n"
    assert net.convert_text_to_action(["stop"]) == [-100]


def test_chunked_prefill_overlap_gives_identical_logits():
    """`chunked_prefill`: the prompt prefix (2 text + 510 patch tokens) is prefilled on the side stream while the 3D memory updates, the rest
    afterwards over the strided KV cache -> the same logits bit for bit as the one-pass prefill."""
    from dynam3d_b200 import synth
    B, V = 2, 2
    outs = []
    for chunked in (False, True):
        net, _ = _build(8, 2, 2, B, q1_fix=True)
        net.tokenize = synth.ToyTokenizer()
        net.chunked_prefill, net.overlap_sms = chunked, (132, 16)
        eps = [synth.make_episode(80 + b, n_steps=2, num_views=V, rgb_size=224, n_seg=16, seg_kind="voronoi") for b in range(B)]
        instr = [synth.make_instruction(b) for b in range(B)]
        for t in range(2):
            obs = {"rgb": torch.from_numpy(np.concatenate([eps[b][t]["rgb"] for b in range(B)], 0)),
                   "depth": torch.from_numpy(np.concatenate([eps[b][t]["depth"] for b in range(B)], 0)),
                   "patch_segm": np.stack([eps[b][t]["segm"] for b in range(B)], 0)}
            lg = net.forward_logits(obs, instr, [eps[b][t]["position"] for b in range(B)], [eps[b][t]["heading"] for b in range(B)], num_of_views=V)
            outs.append(lg.cpu())
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])
