"""CLIP ViT tower on the C ABI vs the precision-matched CPU oracle (oracle/nn_ops.vit_forward, fp16 operand rounding)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(width, layers, heads, res, out_dim, n_img, in_size, pretrain=False, seed=4):
    from dynam3d_b200 import synth
    from dynam3d_b200.clip_vit import ViTEngine, ViTWeights
    from oracle import nn_ops as NN
    sd = synth.vit_state_dict(seed, width=width, layers=layers, resolution=res, out_dim=out_dim)
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(n_img, in_size, in_size, 3), dtype=np.uint8)
    eng = ViTEngine(ViTWeights.from_openai_state_dict(sd), n_head=heads, resolution=res, max_images=n_img)
    cls, patch = eng.forward(torch.from_numpy(img).cuda(), ln_post_on_patches=not pretrain)
    cls, patch = cls.float().cpu(), patch.float().cpu()
    x = NN.clip_preprocess(img, res, rnd=NN.round_fp16)
    want_cls, want_patch = NN.vit_forward(x, sd, layers, heads, rnd=NN.round_fp16, ln_post_on_patches=not pretrain)
    x32 = NN.clip_preprocess(img, res)
    f32_cls, f32_patch = NN.vit_forward(x32, sd, layers, heads, rnd=None, ln_post_on_patches=not pretrain)
    e_p = (patch - want_patch).abs().max().item()
    e_c = (cls - want_cls).abs().max().item()
    e_f = (patch - f32_patch).abs().max().item()
    scale = want_patch.abs().max().item()
    print(f"vit L={layers} W={width}: |patch|max={scale:.3f} err_vs_matched_oracle patch={e_p:.2e} cls={e_c:.2e}; vs_fp32={e_f:.2e}")
    return e_p, e_c, e_f, scale


def test_vit_small_config():
    e_p, e_c, e_f, scale = _run(256, 3, 4, 112, 128, 3, 80)
    # matched oracle: identical rounding points; remaining differences = fp32 summation order + 16-bit output store (2^-11 rel)
    assert e_p <= 2e-3 * max(1.0, scale) and e_c <= 2e-3 * max(1.0, scale)


def test_vit_small_config_pretrain_variant():
    e_p, e_c, e_f, scale = _run(256, 2, 4, 112, 128, 2, 112, pretrain=True)
    assert e_p <= 2e-3 * max(1.0, scale) and e_c <= 2e-3 * max(1.0, scale)


def test_vit_l14_336_full_size_one_view():
    e_p, e_c, e_f, scale = _run(1024, 24, 16, 336, 768, 1, 224)
    assert e_p <= 4e-3 * max(1.0, scale) and e_c <= 4e-3 * max(1.0, scale)
    assert e_f <= 3e-2 * max(1.0, scale)  # distance to the pure-fp32 restatement = the fp16 operand rounding of the reference path
