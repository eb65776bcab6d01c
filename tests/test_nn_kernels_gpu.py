"""Norm / RoPE / attention / preprocessing kernels vs plain PyTorch fp32 references of the same ops (floating point)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,D", [(577, 1024), (37, 768), (600, 3072), (5, 4096)])
def test_layernorm_rmsnorm(T, D):
    from dynam3d_b200 import ops
    x = torch.randn(T, D, device="cuda") * 3 + 0.5
    g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
    o32 = torch.empty_like(x); o16 = torch.empty(T, D, device="cuda", dtype=torch.float16)
    ops.layernorm(x, g, b, 1e-5, out32=o32, out16=o16)
    ref = F.layer_norm(x, (D,), g, b, 1e-5)
    assert (o32 - ref).abs().max().item() < 2e-5 * max(1, ref.abs().max().item())  # fp32, different reduction order
    assert torch.equal(o16, o32.half())
    ops.layernorm(x, g, b, 1e-12, out32=o32, act=2)
    assert (o32 - F.gelu(F.layer_norm(x, (D,), g, b, 1e-12))).abs().max().item() < 5e-5
    idx = torch.randint(0, T, (11,), device="cuda", dtype=torch.int32)
    og = torch.empty(11, D, device="cuda")
    ops.layernorm(x, g, b, 1e-5, out32=og, row_index=idx)
    assert (og - ref[idx.long()]).abs().max().item() < 2e-5 * max(1, ref.abs().max().item())
    ob = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
    ops.rmsnorm(x, g, 1e-5, out32=o32, out16=ob)
    ref = g * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5))
    assert (o32 - ref).abs().max().item() < 2e-5 * max(1, ref.abs().max().item())
    assert torch.equal(ob, o32.bfloat16())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_rope_matches_hf_formula(dtype):
    from dynam3d_b200 import ops
    T, H, Dh = 70, 4, 96
    qkv = torch.randn(T, 3 * H * Dh, device="cuda").to(dtype)
    pos = torch.arange(T, device="cuda", dtype=torch.int32) % 50
    inv_freq = (1.0 / (10000.0 ** (torch.arange(0, Dh, 2).float() / Dh))).cuda()
    ref = qkv.float().clone()
    ref_in = qkv.clone()
    freqs = pos.float()[:, None] * inv_freq[None]
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]
    for part in (0, 1):
        x = ref[:, part * H * Dh:(part + 1) * H * Dh].view(T, H, Dh)
        rot = torch.cat([-x[..., Dh // 2:], x[..., :Dh // 2]], -1)
        ref[:, part * H * Dh:(part + 1) * H * Dh] = (x * cos + rot * sin).reshape(T, H * Dh)
    ops.rope(qkv, pos, inv_freq, H, Dh)
    tol = 4e-3 if dtype == torch.float16 else 3e-2  # one 16-bit rounding of values up to ~4
    assert (qkv.float() - ref).abs().max().item() < tol
    assert torch.equal(qkv[:, 2 * H * Dh:].float(), ref[:, 2 * H * Dh:])  # v untouched
    # the table-based vectorised form used by the LM engine is bit-identical
    qkv2 = ref_in.clone()
    ops.rope_apply(qkv2, ops.rope_table(pos, inv_freq, Dh), H, Dh)
    assert torch.equal(qkv2, qkv)


@pytest.mark.parametrize("impl", ["simt", "mma", "tc"])
@pytest.mark.parametrize("lens,H,Dh,causal,dtype", [
    ([577, 577], 16, 64, False, torch.float16),
    ([1, 2, 33, 64, 65, 300, 130], 12, 64, False, torch.float16),
    ([600, 75], 4, 96, True, torch.bfloat16),
    ([700, 130, 128, 5], 3, 64, True, torch.float16),
    ([577] * 6, 16, 64, False, torch.bfloat16),
    ([128], 2, 96, True, torch.float16),
    ([735, 745, 300, 129], 8, 96, True, torch.float16),   # Phi-3 prefill shape (head_dim 96, causal, ragged)
    ([200, 65], 2, 96, False, torch.bfloat16),
])
def test_attention_matches_torch(lens, H, Dh, causal, dtype, impl):
    from dynam3d_b200 import ops
    if impl == "tc" and Dh not in (64, 96):
        pytest.skip("tcgen05 attention is built for head_dim 64 and 96")
    T = sum(lens)
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 0.7).to(dtype)
    out = torch.zeros(T, H * Dh, device="cuda", dtype=dtype)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
    ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=causal, impl=impl)
    s = 0
    for n in lens:
        q, k, v = [t.view(n, H, Dh).transpose(0, 1).float() for t in qkv[s:s + n].split(H * Dh, -1)]
        ref = F.scaled_dot_product_attention(q[None], k[None], v[None], is_causal=causal)[0].transpose(0, 1).reshape(n, H * Dh)
        tol = 3e-3 if dtype == torch.float16 else 2e-2  # 16-bit output rounding (+ 16-bit P on the tensor-core path)
        assert (out[s:s + n].float() - ref).abs().max().item() < tol, (n, impl)
        s += n


@pytest.mark.parametrize("halves", [1, 2])
@pytest.mark.parametrize("lens,H,Dh,causal,dtype", [
    ([577, 577, 64, 65, 1], 16, 64, False, torch.float16),
    ([700, 130, 128, 5], 3, 64, True, torch.bfloat16),
    ([735, 745, 300, 129, 63], 8, 96, True, torch.float16),
    ([200, 65], 2, 96, False, torch.bfloat16),
])
def test_attention_tc_tile_shapes(lens, H, Dh, causal, dtype, halves):
    """Both tile shapes of the tcgen05 attention (64- and 128-key tiles) at both head dims vs fp32 torch."""
    from dynam3d_b200 import ops, _lib as L
    T = sum(lens)
    g = torch.Generator().manual_seed(T + Dh)
    qkv = (torch.randn(T, 3 * H * Dh, generator=g) * 0.7).to(dtype).cuda()
    out = torch.zeros(T, H * Dh, device="cuda", dtype=dtype)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
    L.check(L.lib().d3d_attention_tc_set_halves(halves, halves))
    try:
        ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=causal, impl="tc")
        torch.cuda.synchronize()
    finally:
        L.check(L.lib().d3d_attention_tc_set_halves(2, 1))  # the defaults
    s = 0
    for n in lens:
        q, k, v = [t.view(n, H, Dh).transpose(0, 1).float() for t in qkv[s:s + n].split(H * Dh, -1)]
        ref = F.scaled_dot_product_attention(q[None], k[None], v[None], is_causal=causal)[0].transpose(0, 1).reshape(n, H * Dh)
        tol = 3e-3 if dtype == torch.float16 else 2e-2
        assert (out[s:s + n].float() - ref).abs().max().item() < tol, (n, halves)
        s += n


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("lens", [[1, 2, 33, 64, 65, 300, 130, 16, 17, 48], [37] * 1500 + [100, 3], [5, 64, 1, 31]])
def test_attention_mixed_lengths(lens, dtype):
    """The pooling passes' dispatch (short sequences on the warp-per-(sequence, head) kernel, long ones on the 128-row kernel) vs fp32 torch,
    and identical within 16-bit rounding to the single-kernel path it replaces."""
    from dynam3d_b200 import ops
    H, Dh = 12, 64
    T = sum(lens)
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 0.7).to(dtype)
    out = torch.full((T, H * Dh), float("nan"), device="cuda", dtype=dtype)
    one = torch.zeros_like(out)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
    ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, impl="mixed")
    ops.attention(qkv, one, cu, len(lens), max(lens), H, Dh, impl="mma" if max(lens) >= 64 else "simt")
    assert torch.isfinite(out.float()).all()
    tol = 3e-3 if dtype == torch.float16 else 2e-2
    assert (out.float() - one.float()).abs().max().item() < tol
    s = 0
    for n in lens[:12]:
        q, k, v = [t.view(n, H, Dh).transpose(0, 1).float() for t in qkv[s:s + n].split(H * Dh, -1)]
        ref = F.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1).reshape(n, H * Dh)
        assert (out[s:s + n].float() - ref).abs().max().item() < tol, n
        s += n


@pytest.mark.parametrize("size", [336, 224])
def test_preprocess_im2col(size):
    from dynam3d_b200 import ops
    g = torch.Generator().manual_seed(size)
    img = torch.randint(0, 256, (3, size, size, 3), generator=g, dtype=torch.uint8)
    cols = ops.preprocess_im2col(img.cuda()).cpu().float()
    x = img.permute(0, 3, 1, 2).float()
    if size != 336:  # torchvision 0.14 Resize on a uint8 tensor: float bicubic, then round + clamp back to uint8 (ENC:268)
        x = F.interpolate(x, size=(336, 336), mode="bicubic", align_corners=False).round().clamp(0, 255)
    x = x / 255.0
    mean = torch.tensor(ops.CLIP_MEAN).view(1, 3, 1, 1); std = torch.tensor(ops.CLIP_STD).view(1, 3, 1, 1)
    x = ((x - mean) / std).half().float()
    ref = F.unfold(x, 14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert cols.shape == (3 * 576, 592) and torch.all(cols[:, 588:] == 0)
    diff = (cols[:, :588] - ref).abs()
    # a resized pixel may land on the other side of a .5 rounding boundary (1/255/std ~ 0.015): allow <= 1e-4 of pixels
    bad = (diff > 2e-3).float().mean().item()
    assert bad <= (1e-4 if size != 336 else 0.0), bad
    assert diff.max().item() < 0.02


def test_bicubic_resize_is_bit_identical_to_torch_cuda():
    """a1 byte work: the resized uint8 pixels equal torch's own CUDA bicubic kernel (what torchvision's Resize, ENC:268, runs for observations on
    the GPU) on EVERY pixel; torch's CPU kernel orders the same formula differently (<= 1e-4 of pixels flip across a .5 boundary, checked above)."""
    from dynam3d_b200 import ops
    g = torch.Generator().manual_seed(7)
    img = torch.randint(0, 256, (4, 224, 224, 3), generator=g, dtype=torch.uint8)
    img[1] = (torch.arange(224 * 224 * 3) % 251).view(224, 224, 3).to(torch.uint8)  # smooth ramps: many exact .5 candidates
    cols = ops.preprocess_im2col(img.cuda(), out_dtype=torch.float16).float()
    x = F.interpolate(img.cuda().permute(0, 3, 1, 2).float(), size=(336, 336), mode="bicubic", align_corners=False).round().clamp(0, 255)
    mean = torch.tensor(ops.CLIP_MEAN, device="cuda").view(1, 3, 1, 1); std = torch.tensor(ops.CLIP_STD, device="cuda").view(1, 3, 1, 1)
    ref = F.unfold(((x / 255.0 - mean) / std).half().float(), 14, stride=14).transpose(1, 2).reshape(-1, 588)
    # one uint8 step is 1/255/std ~ 0.015 in normalised units: any pixel flip would show as a difference >> the fp16 rounding of equal pixels
    n_flip = int(((cols[:, :588] - ref).abs() > 5e-3).sum())
    assert n_flip == 0, n_flip
