"""The one-call entries (csrc/forward_host.cu: d3d_vit_forward, d3d_phi3_prefill) against the per-kernel layer loops of the Python engines
(which are themselves compared with the oracle in test_vit_gpu.py / test_lm_gpu.py): same kernels in the same order -> bit-identical;
the trimmed final Phi-3 layer (last-token rows only, skinny GEMM) within fp32 accumulation-order noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _loop(fn):
    """Run fn() through the per-kernel Python loop (the engines take it whenever the per-stage profile is recording)."""
    from dynam3d_b200 import ops
    ops.STAGE_PROFILE = []
    try:
        return fn()
    finally:
        ops.STAGE_PROFILE = None


@pytest.mark.parametrize("size", [336, 224])
def test_vit_forward_entry_equals_layer_loop(size):
    from dynam3d_b200 import synth
    from dynam3d_b200.clip_vit import ViTEngine, ViTWeights
    sd = synth.vit_state_dict(12, layers=3)
    eng = ViTEngine(ViTWeights.from_openai_state_dict(sd), n_head=16, resolution=336, max_images=3)
    img = torch.from_numpy(np.random.default_rng(size).integers(0, 256, size=(3, size, size, 3), dtype=np.uint8)).cuda()
    cls1, patch1 = eng.forward(img)
    cls1, patch1 = cls1.clone(), patch1.clone()
    cls2, patch2 = _loop(lambda: eng.forward(img))
    assert torch.equal(cls1, cls2) and torch.equal(patch1, patch2)
    hid1 = eng.forward(img, n_layers_run=2, project=False).clone()
    hid2 = _loop(lambda: eng.forward(img, n_layers_run=2, project=False))
    assert torch.equal(hid1, hid2)


def test_phi3_prefill_entry_equals_layer_loop():
    from dynam3d_b200 import synth
    from dynam3d_b200.phi3 import LMEngine, LMWeights
    hidden, layers, heads, ffn, vocab, lens = 3072, 2, 32, 8192, 32064, [300, 411, 97]
    sd = synth.lm_state_dict(8, hidden, layers, ffn, vocab, device="cuda", round_to=torch.float16)
    eng = LMEngine(LMWeights.from_state_dict(sd, dtype=torch.float16), n_heads=heads, max_tokens=sum(lens))
    emb = synth.hash_uniform((sum(lens), hidden), 108, 1.0, device="cuda")
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
    last = (cu[1:] - 1).to(torch.int32).contiguous()
    want = _loop(lambda: eng.prefill(emb.clone(), cu, pos, len(lens), max(lens), last))
    eng.trim_last_layer = False
    got = eng.prefill(emb.clone(), cu, pos, len(lens), max(lens), last)
    assert torch.equal(got, want)
    eng.trim_last_layer = True
    got_t = eng.prefill(emb.clone(), cu, pos, len(lens), max(lens), last)
    e = (got_t - want).abs().max().item()
    print(f"trimmed final layer vs full: max abs logit diff {e:.2e} (|logit| max {want.abs().max().item():.2f})")
    assert e < 2e-3 and torch.equal(got_t.argmax(-1), want.argmax(-1))  # skinny-GEMM accumulation order -> a few fp16 roundings of the MLP operand flip
    # with a KV cache (generate): the cache rows are the same packed QKV matrices
    want_kv = _loop(lambda: eng.prefill(emb.clone(), cu, pos, len(lens), max(lens), last, kv_rows=8))
    kv_loop = eng.kv[:, : sum(lens)].clone()
    got_kv = eng.prefill(emb.clone(), cu, pos, len(lens), max(lens), last, kv_rows=8)
    assert torch.equal(eng.kv[:, : sum(lens)], kv_loop) and (got_kv - want_kv).abs().max().item() < 2e-3


def test_chunked_prefill_equals_one_pass():
    """d3d_phi3_prefill_chunk: the first 512 tokens of every sequence in one pass, the rest (attending to the cached prefix) in a second pass over a
    strided KV cache -> bit-identical logits to the one-pass prefill (same kernels per row / tile)."""
    from dynam3d_b200 import synth
    from dynam3d_b200.phi3 import LMEngine, LMWeights
    hidden, layers, heads, ffn, vocab, lens = 3072, 2, 32, 8192, 32064, [745, 612, 701]
    B, P, stride = len(lens), 512, 1024
    sd = synth.lm_state_dict(9, hidden, layers, ffn, vocab, device="cuda", round_to=torch.float16)
    eng = LMEngine(LMWeights.from_state_dict(sd, dtype=torch.float16), n_heads=heads, max_tokens=sum(lens))
    emb = synth.hash_uniform((sum(lens), hidden), 109, 1.0, device="cuda")
    cu_h = np.concatenate([[0], np.cumsum(lens)])
    cu = torch.tensor(cu_h, dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
    last = (cu[1:] - 1).to(torch.int32).contiguous()
    for trim in (False, True):
        eng.trim_last_layer = trim
        want = eng.prefill(emb.clone(), cu, pos, B, max(lens), last)
        eng.chunk_cache(B, stride)
        i32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.int32, device="cuda")
        starts = i32([b * stride for b in range(B)])
        # pass 1: positions [0, 512) of every sequence
        Xp = torch.cat([emb[cu_h[b]:cu_h[b] + P] for b in range(B)], 0).contiguous()
        rows1 = i32(np.concatenate([b * stride + np.arange(P) for b in range(B)]))
        pos1 = i32(np.tile(np.arange(P), B))
        assert eng.prefill_chunk(Xp, starts, i32([P] * B), rows1, pos1, B, P, 0, P // 128) is None
        # pass 2: positions [512, len)
        Xs = torch.cat([emb[cu_h[b] + P:cu_h[b + 1]] for b in range(B)], 0).contiguous()
        rows2 = i32(np.concatenate([b * stride + np.arange(P, lens[b]) for b in range(B)]))
        pos2 = i32(np.concatenate([np.arange(P, lens[b]) for b in range(B)]))
        last2 = i32(np.cumsum([n - P for n in lens]) - 1)
        got = eng.prefill_chunk(Xs, starts, i32(lens), rows2, pos2, B, max(lens), P // 128, 1 << 20, last_rows=last2)
        assert torch.equal(got, want), (trim, (got - want).abs().max().item())
