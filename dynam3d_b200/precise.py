"""The "precise" pipeline: the same kernels, fp32 activations between them, split fp16x2 tensor-core operands.

Purpose: parity evidence.  The production path uses fp16 GEMM operands like the reference's fp16 autocast, whose rounding-flip
noise floor (~4e-3 on the action logits) is above the north star's 1e-3 tolerance.  This mode computes every
GEMM as A_hi W_hi + A_lo W_hi (+ A_hi W_lo when the weight is not fp16-representable) on the unchanged tcgen05 kernel by
concatenating the split operands along K, keeps q/k/v and all other activations in fp32, and is compared against the oracle's
pure-fp32 path (= the arithmetic of the reference on CPU).  About 2-3x slower than the production mode; never used by bench.py.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib as L
from . import ops

_WCACHE = {}


def split_weight(w, device=None):
    """[N,K] weight (any float dtype) -> ([N, terms*K] fp16 = [W_hi | W_hi (| W_lo)], terms)."""
    key = (w.data_ptr(), tuple(w.shape), w.dtype)
    hit = _WCACHE.get(key)
    if hit is not None:
        return hit
    w32 = w.detach().to(device=device or w.device, dtype=torch.float32)
    hi = w32.to(torch.float16)
    lo = (w32 - hi.to(torch.float32)).to(torch.float16)
    if bool((lo == 0).all()):
        out = (torch.cat([hi, hi], 1).contiguous(), 2)
    else:
        out = (torch.cat([hi, hi, lo], 1).contiguous(), 3)
    _WCACHE[key] = out + (w,)  # keep `w` alive so the data_ptr key stays valid
    return _WCACHE[key]


def linear(x32, w, bias=None, act=L.ACT_NONE, residual=None, out=None, w_prepared=None):
    """fp32 [T,K] @ W^T with split operands; fp32 output."""
    wx, terms = (w_prepared or split_weight(w))[:2]
    T, K = x32.shape
    assert x32.dtype == torch.float32 and x32.stride(1) == 1 and wx.shape[1] == terms * K, (wx.shape, terms, K)
    a = torch.empty((T, terms * K), device=x32.device, dtype=torch.float16)
    L.check(L.lib().d3d_split16(L.ptr(x32), x32.stride(0), L.ptr(a), a.stride(0), T, K, terms, L.stream_ptr()))
    if out is None:
        n_out = wx.shape[0] // 2 if act == L.ACT_SWIGLU else wx.shape[0]
        out = torch.empty((T, (n_out + 3) // 4 * 4), device=x32.device, dtype=torch.float32)[:, :n_out]  # 16-byte aligned rows
    return ops.gemm(a, wx, out=out, bias=bias, act=act, residual=residual)


SPLIT_TC = True  # split attention of sequences >= 256 tokens (head_dim 64 / 96) on tcgen05 (csrc/attention_tc.cu, SPLIT); False = the mma.sync kernel
SPLIT_ATTENTION = True  # sequences of >= 64 tokens on the tensor cores with split fp16x2 operands (csrc/attention_split.cu); False = fp32 CUDA cores


def attention(qkv32, cu, n_seq, max_len, H, Dh, causal):
    T, W = qkv32.shape[0], 3 * H * Dh
    out = torch.empty((T, H * Dh), device=qkv32.device, dtype=torch.float32)
    split = SPLIT_ATTENTION and max_len >= 64
    work = 4.0 * (T ** 2 / max(n_seq, 1)) * Dh * H * (0.5 if causal else 1.0) * (3.0 if split else 1.0)
    with ops._Rec("attention_precise", "tensor", work):
        if split:  # split the layer's QKV matrix ONCE into [hi | lo] fp16 halves, then the pipelined tensor-core kernel
            hl = torch.empty((T, 2 * W), device=qkv32.device, dtype=torch.float16)
            L.check(L.lib().d3d_split16(L.ptr(qkv32), qkv32.stride(0), L.ptr(hl), hl.stride(0), T, W, 2, L.stream_ptr()))
            if SPLIT_TC and Dh in (64, 96) and max_len >= 256:
                L.check(L.lib().d3d_attention_split_tc(L.ptr(hl), hl.stride(0), T, W, L.ptr(out), out.stride(0), L.ptr(cu), n_seq, max_len, H, Dh,
                                                       int(bool(causal)), 1.0 / math.sqrt(Dh), L.stream_ptr()))
            else:
                L.check(L.lib().d3d_attention_split(L.ptr(hl), hl.stride(0), W, L.ptr(out), out.stride(0), L.ptr(cu), n_seq, max_len, H, Dh,
                                                    int(bool(causal)), 1.0 / math.sqrt(Dh), L.stream_ptr()))
        else:
            L.check(L.lib().d3d_attention_f32(L.ptr(qkv32), qkv32.stride(0), L.ptr(out), out.stride(0), L.ptr(cu), n_seq, max_len, H, Dh,
                                              int(bool(causal)), 1.0 / math.sqrt(Dh), L.stream_ptr()))
    return out


def mlp_ln_gelu(x32, m):
    """nn.Sequential(Linear, LayerNorm, GELU, Linear) with fp32 master weights m = {w0, b0, g, b, w3, b3}; x32 [T, kpad]."""
    h = linear(x32, m["w0"], m["b0"])
    ops.layernorm(h, m["g"], m["b"], 1e-5, out32=h, act=L.ACT_GELU)
    return linear(h, m["w3"], m["b3"])


def encoder(X, cu, n_seq, max_len, e):
    """2-layer post-norm TransformerEncoder + final LayerNorm on token 0 of every sequence (fp32 master weights)."""
    D = X.shape[1]
    for l in e["layers"]:
        qkv = linear(X, l["w_in"], l["b_in"])
        att = attention(qkv, cu, n_seq, max_len, D // 64, 64, False)
        linear(att, l["w_out"], l["b_out"], residual=X, out=X)
        ops.layernorm(X, l["n1"][0], l["n1"][1], 1e-5, out32=X)
        h = linear(X, l["w1"], l["b1"], act=L.ACT_GELU)
        linear(h, l["w2"], l["b2"], residual=X, out=X)
        ops.layernorm(X, l["n2"][0], l["n2"][1], 1e-5, out32=X)
    out = torch.empty((n_seq, D), device=X.device, dtype=torch.float32)
    ops.layernorm(X, e["norm"][0], e["norm"][1], e["eps"], out32=out, row_index=cu[:n_seq])
    return out


def vit_forward(eng, img_u8, n_layers_run=None, ln_post_on_patches=True, project=True, fp16_pixels=False):
    """Precise counterpart of ViTEngine.forward (weights of the tower are exactly fp16, so 2-term splits).
    fp16_pixels: the normalised pixels are rounded to fp16 first -- POL:438 casts the HF processor's pixel_values with `.to(device, torch.float16)`
    whatever precision the model runs in."""
    w = eng.w
    ops.STAGE_TAG = eng.tag
    N = img_u8.shape[0]
    T = N * w.tokens
    dev = img_u8.device
    k = 3 * w.patch * w.patch
    g = eng.R // w.patch
    if fp16_pixels:
        cols = ops.preprocess_im2col(img_u8, eng.R, w.patch, torch.float16).float()
    else:
        cols = torch.empty((N * g * g, w.kpad), device=dev, dtype=torch.float32)
        m, mp = ops._hp_f32(ops.CLIP_MEAN)
        s, sp = ops._hp_f32(ops.CLIP_STD)
        L.check(L.lib().d3d_preprocess_im2col(L.ptr(img_u8), N, img_u8.shape[1], img_u8.shape[2], eng.R, w.patch, mp, sp, L.ptr(cols), w.kpad,
                                              L.D3D_OUT_F32, L.stream_ptr()))
    conv = linear(cols, w.conv_w)
    X = torch.empty((T, w.width), device=dev, dtype=torch.float32)
    ops.vit_embed_ln(conv, w.cls, w.pos, w.ln_pre[0], w.ln_pre[1], 1e-5, N, w.tokens, X)
    cu = (torch.arange(N + 1, device=dev, dtype=torch.int32) * w.tokens).contiguous()
    Dh = w.width // eng.H
    h = torch.empty_like(X)
    run = len(w.layers) if n_layers_run is None else n_layers_run
    for l in range(run):
        p = w.layers[l]
        ops.layernorm(X, p["ln1"][0], p["ln1"][1], 1e-5, out32=h)
        qkv = linear(h, p["w_qkv"], p["b_qkv"])
        att = attention(qkv, cu, N, w.tokens, eng.H, Dh, False)
        linear(att, p["w_o"], p["b_o"], residual=X, out=X)
        ops.layernorm(X, p["ln2"][0], p["ln2"][1], 1e-5, out32=h)
        f = linear(h, p["w_fc"], p["b_fc"], act=L.ACT_QUICK_GELU)
        linear(f, p["w_pr"], p["b_pr"], residual=X, out=X)
    if not project:
        return X.view(N, w.tokens, w.width)
    if ln_post_on_patches:
        ops.layernorm(X, w.ln_post[0], w.ln_post[1], 1e-5, out32=h)
    else:
        h.copy_(X)
        cls_rows = (torch.arange(N, device=dev, dtype=torch.int32) * w.tokens).contiguous()
        tmp = torch.empty((N, w.width), device=dev, dtype=torch.float32)
        ops.layernorm(X, w.ln_post[0], w.ln_post[1], 1e-5, out32=tmp, row_index=cls_rows)
        h.view(N, w.tokens, w.width)[:, 0].copy_(tmp)
    out = linear(h, w.proj).view(N, w.tokens, w.out_dim)
    return out[:, 0], out[:, 1:]


def lm_prefill(eng, X, cu, positions, n_seq, max_len, last_rows):
    """Precise counterpart of LMEngine.prefill (X fp32 [T, hidden] is overwritten)."""
    w = eng.w
    ops.STAGE_TAG = "lm"
    T = X.shape[0]
    h = torch.empty_like(X)
    for p in w.layers:
        ops.rmsnorm(X, p["rms1"], eng.eps, out32=h)
        qkv = linear(h, p["w_qkv"])
        L.check(L.lib().d3d_rope(L.ptr(qkv), qkv.stride(0), L.ptr(positions), L.ptr(eng.inv_freq), T, eng.H, eng.Dh, L.D3D_OUT_F32, L.stream_ptr()))
        att = attention(qkv, cu, n_seq, max_len, eng.H, eng.Dh, True)
        linear(att, p["w_o"], residual=X, out=X)
        ops.rmsnorm(X, p["rms2"], eng.eps, out32=h)
        f = linear(h, p["w_gu"], act=L.ACT_SWIGLU)
        linear(f, p["w_down"], residual=X, out=X)
    last = torch.empty((n_seq, w.hidden), device=X.device, dtype=torch.float32)
    ops.rmsnorm(X, w.norm, eng.eps, out32=last, row_index=last_rows)
    return linear(last, w.lm_head)
