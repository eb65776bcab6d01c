"""Greedy decode with the KV cache (POL:463-469) vs the cache-free CPU oracle, plus its kernels in isolation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias, act, residual):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 4:
        y = torch.nn.functional.silu(y[:, 0::2]) * y[:, 1::2]
    if residual is not None:
        y = y + residual
    return y


@pytest.mark.parametrize("M,N,K,dtype,use_bias,act,use_res,out_dtype", [
    (8, 9216, 3072, torch.float16, False, 0, False, torch.float16),
    (8, 16384, 3072, torch.float16, False, 4, False, torch.float16),
    (8, 3072, 8192, torch.float16, False, 0, True, torch.float32),
    (3, 32064, 3072, torch.bfloat16, False, 0, False, torch.float32),
    (16, 72, 64, torch.float16, True, 0, False, torch.float32),
    (1, 23, 96, torch.bfloat16, True, 0, True, torch.float32),
    (11, 40, 3072, torch.float16, False, 4, False, torch.float32),
])
@pytest.mark.parametrize("variant", [0, 10])  # 0 = register-staged kernels (heuristic), 10 = cp.async.bulk ring
def test_skinny_gemm_matches_torch(M, N, K, dtype, use_bias, act, use_res, out_dtype, variant):
    """HBM-bound decode GEMM (1..16 rows) vs a plain PyTorch fp32 reference; only the summation order differs (tolerances as test_gemm_gpu)."""
    import ctypes
    from dynam3d_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1 if use_bias else None
    n_out = N // 2 if act == 4 else N
    ldc = (n_out + 7) // 8 * 8
    res = torch.randn(M, ldc, device="cuda", generator=g) if use_res else None
    out = torch.zeros((M, ldc), device="cuda", dtype=out_dtype)
    args = L.GemmArgs(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), ldc, M, N, K, L.kind_of(dtype), L.kind_of(out_dtype), L.ptr(bias),
                      act, L.ptr(res), ldc if use_res else 0)
    try:
        L.lib().d3d_gemm_skinny_set_config(variant)
        L.check(L.lib().d3d_gemm_skinny(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
        torch.cuda.synchronize()
    finally:
        L.lib().d3d_gemm_skinny_set_config(0)
    ref = _ref(a, w, bias, act, res[:, :n_out] if use_res else None)
    tol = 2e-3 if out_dtype == torch.float32 else 8e-3
    err = (out[:, :n_out].float() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), err
    assert float(out[:, n_out:].abs().max()) == 0.0 if ldc > n_out else True  # nothing written past N


def test_argmax_first_maximum():
    from dynam3d_b200 import _lib as L
    x = torch.randn(5, 32064, device="cuda")
    x[0, 100] = x[0, 31000] = 50.0   # tie -> lowest index
    x[1, 32063] = 60.0
    x[2, 0] = 70.0
    out = torch.empty(5, device="cuda", dtype=torch.int32)
    L.check(L.lib().d3d_argmax_rows(L.ptr(x), x.stride(0), 5, 32064, L.ptr(out), L.stream_ptr()))
    assert out.cpu().tolist() == x.argmax(-1).cpu().tolist() and out[0].item() == 100


@pytest.mark.parametrize("H,Dh,lens,dtype", [(32, 96, [735, 745, 16, 1, 33], torch.float16), (8, 64, [40, 3], torch.bfloat16), (4, 128, [17, 300], torch.float16)])
def test_decode_attention_kernel(H, Dh, lens, dtype):
    """The decode attention in isolation: (1) vs fp32 torch over [prefill rows || earlier decode rows || own row]; (2) with the RoPE of the
    step's rows fused the output AND the cache rows are bit-identical to d3d_rope_apply followed by the plain call."""
    import ctypes
    from dynam3d_b200 import _lib as L, ops
    n_seq, step = len(lens), 2
    T = sum(lens)
    rows = T + (step + 1) * n_seq
    g = torch.Generator().manual_seed(H * 1000 + Dh)
    qkv = (torch.randn(rows, 3 * H * Dh, generator=g) * 0.6).to(dtype).cuda()
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
    pos = torch.tensor([n + step for n in lens], device="cuda", dtype=torch.int32)
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, Dh, 2, dtype=torch.float32) / Dh))).cuda()
    tab = ops.rope_table(pos, inv_freq, Dh)
    scale = ctypes.c_float(Dh ** -0.5)
    kind = L.kind_of(dtype)
    own = slice(T + step * n_seq, T + (step + 1) * n_seq)
    # reference path: rotate the step's rows, then the plain kernel
    ref_cache = qkv.clone()
    ops.rope_apply(ref_cache[own], tab, H, Dh)
    ref = torch.empty(n_seq, H * Dh, device="cuda", dtype=dtype)
    L.check(L.lib().d3d_decode_attention(L.ptr(ref_cache), ref_cache.stride(0), L.ptr(cu), n_seq, T, step, H, Dh, kind, scale, L.ptr(ref), ref.stride(0),
                                         L.stream_ptr()))
    # fused path on the un-rotated rows
    cache = qkv.clone()
    out = torch.empty_like(ref)
    L.check(L.lib().d3d_decode_attention_rope(L.ptr(cache), cache.stride(0), L.ptr(cu), n_seq, T, step, H, Dh, kind, scale, L.ptr(tab), L.ptr(out),
                                              out.stride(0), L.stream_ptr()))
    assert torch.equal(out, ref)
    HD = H * Dh
    assert torch.equal(cache[own, HD:], ref_cache[own, HD:])          # K rotated in place, V untouched
    assert torch.equal(cache[:T + step * n_seq], qkv[:T + step * n_seq])  # nothing else written
    # fp32 torch reference
    c32 = ref_cache.float()
    for b, n in enumerate(lens):
        idx = list(range(int(cu[b]), int(cu[b]) + n)) + [T + s * n_seq + b for s in range(step + 1)]
        q = c32[T + step * n_seq + b, :HD].view(H, Dh)
        k = c32[idx, HD:2 * HD].view(-1, H, Dh)
        v = c32[idx, 2 * HD:].view(-1, H, Dh)
        p = torch.softmax(torch.einsum("hd,jhd->hj", q, k) * Dh ** -0.5, -1)
        want = torch.einsum("hj,jhd->hd", p, v).reshape(-1)
        tol = 3e-3 if dtype == torch.float16 else 2e-2   # 16-bit P and output rounding
        assert (ref[b].float() - want).abs().max().item() < tol, b


@pytest.mark.parametrize("cfg", [dict(hidden=768, layers=3, heads=8, ffn=1536, vocab=2048, lens=[37, 70, 5], dtype=torch.float16, n_new=6),
                                 dict(hidden=768, layers=2, heads=8, ffn=1536, vocab=2048, lens=[33], dtype=torch.bfloat16, n_new=4),
                                 dict(hidden=3072, layers=2, heads=32, ffn=8192, vocab=32064, lens=[90, 41], dtype=torch.float16, n_new=4)])
def test_decode_matches_cache_free_oracle(cfg):
    """Teacher-forced comparison: the engine decodes greedily with its KV cache; the oracle re-runs the full prefill over
    [prompt || the engine's tokens] for every step (the arithmetic of HF generate without a cache).  Per-step logits must agree within the
    16-bit-operand tolerance of the prefill tests, and every token the engine chose must be an arg-max of the oracle up to that tolerance."""
    from dynam3d_b200 import synth
    from dynam3d_b200.phi3 import LMEngine, LMWeights
    from oracle import nn_ops as NN
    dtype, lens, n_new = cfg["dtype"], cfg["lens"], cfg["n_new"]
    rnd = NN.round_fp16 if dtype == torch.float16 else NN.round_bf16
    sd = synth.lm_state_dict(9, cfg["hidden"], cfg["layers"], cfg["ffn"], cfg["vocab"], round_to=dtype)
    emb = synth.hash_uniform((sum(lens), cfg["hidden"]), 77, 1.0)
    eng = LMEngine(LMWeights.from_state_dict(sd, dtype=dtype), n_heads=cfg["heads"], max_tokens=sum(lens))
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
    last = (cu[1:] - 1).to(torch.int32).contiguous()
    steps = []
    logits0, toks = eng.generate(emb.cuda().clone(), cu, pos, len(lens), max(lens), last, max_new_tokens=n_new, eos_ids=(), step_logits=steps)
    assert all(len(t) == n_new for t in toks)
    got = torch.stack([logits0.cpu()] + [s.cpu() for s in steps], 0)  # [n_new, B, vocab]: entry s chose toks[b][s]
    want = NN.lm_teacher_forced_logits(emb, lens, [t[:n_new - 1] for t in toks], sd, cfg["layers"], cfg["heads"], rnd=rnd)
    tol = (8e-3 if dtype == torch.float16 else 4e-2)
    err = (got - want).abs().max().item()
    print(f"decode {cfg['hidden']}x{cfg['layers']} {dtype}: max |logit| {want.abs().max().item():.2f}, err over {n_new} steps {err:.2e}")
    assert err <= tol
    for s in range(n_new):
        for b in range(len(lens)):
            assert want[s, b, toks[b][s]] >= want[s, b].max() - 2 * tol, (s, b)


def test_decode_launch_modes_agree():
    """Programmatic dependent launch on / off and the two decode-attention kernels: same tokens; the PDL switch alone is bit-identical
    (it only changes WHEN kernels start), the attention kernels differ by 16-bit rounding of P."""
    from dynam3d_b200 import synth, _lib as L
    from dynam3d_b200.phi3 import LMEngine, LMWeights
    lens = [70, 33, 5]
    sd = synth.lm_state_dict(11, 768, 3, 1536, 2048, round_to=torch.float16)
    emb = synth.hash_uniform((sum(lens), 768), 78, 1.0).cuda()
    eng = LMEngine(LMWeights.from_state_dict(sd, dtype=torch.float16), n_heads=8, max_tokens=sum(lens))
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
    last = (cu[1:] - 1).to(torch.int32).contiguous()
    res = {}
    try:
        for mode in (1, 0, 3):
            L.check(L.lib().d3d_lm_decode_set_pdl(mode))
            steps = []
            _, toks = eng.generate(emb.clone(), cu, pos, len(lens), max(lens), last, max_new_tokens=6, eos_ids=(), step_logits=steps)
            res[mode] = (toks, torch.stack(steps).cpu())
    finally:
        L.check(L.lib().d3d_lm_decode_set_pdl(1))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])
    assert (res[3][1][0] - res[1][1][0]).abs().max().item() < 5e-3  # first decode step (later ones depend on the tokens chosen)


def test_generate_stops_at_eos():
    from dynam3d_b200 import synth
    from dynam3d_b200.phi3 import LMEngine, LMWeights
    sd = synth.lm_state_dict(3, 768, 2, 1536, 2048, round_to=torch.float16)
    lens = [20, 31]
    emb = synth.hash_uniform((sum(lens), 768), 5, 1.0)
    eng = LMEngine(LMWeights.from_state_dict(sd, dtype=torch.float16), n_heads=8, max_tokens=sum(lens))
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
    last = (cu[1:] - 1).to(torch.int32).contiguous()
    _, free = eng.generate(emb.cuda().clone(), cu, pos, 2, max(lens), last, max_new_tokens=8)
    eos = {free[0][2], free[1][4]}  # make the 3rd token of sequence 0 and the 5th of sequence 1 the EOS ids
    _, cut = eng.generate(emb.cuda().clone(), cu, pos, 2, max(lens), last, max_new_tokens=8, eos_ids=eos)
    for b in range(2):
        want = []
        for t in free[b]:
            want.append(t)
            if t in eos:  # HF generate keeps the EOS id as the last output token
                break
        assert cut[b] == want and len(cut[b]) < 8
