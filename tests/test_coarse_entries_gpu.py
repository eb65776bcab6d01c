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
