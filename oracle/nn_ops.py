"""TEST INFRASTRUCTURE ONLY -- fp32 PyTorch (CPU) restatement of the neural blocks on Dynam3D's hot path.

`rnd` is the 16-bit rounding hook that mirrors where the CUDA engine stores fp16/bf16
(the A operand of every tensor-core GEMM, fp16 QKV / attention outputs).  With `rnd=None` this is
a plain fp32 restatement of the reference modules (used to pin the oracle against the reference
run on CPU); with `rnd=round_fp16` it is the precision-matched oracle the `-m gpu` parity tests use.
The reference itself runs these blocks under `torch.cuda.amp.autocast()` (ss_trainer_Dynam3D.py:385),
i.e. with fp16 GEMM operands, so 16-bit operand rounding is the reference path's own precision.

FF    = Dynam3D_VLN/vlnce_baselines/models/feature_fields.py
CLIPM = Dynam3D_VLN/vlnce_baselines/models/encoders/clip/model.py
"""
import math

import torch
import torch.nn.functional as F


def round_fp16(t):
    return t.to(torch.float16).to(torch.float32)


def round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _id(t):
    return t


def linear(x, w, b=None, rnd=None):
    """y = rnd(x) @ rnd(w)^T + b, fp32 accumulate."""
    r = rnd or _id
    y = r(x) @ r(w.to(torch.float32)).t()
    if b is not None:
        y = y + b.to(torch.float32)
    return y


def layer_norm(x, w, b, eps):
    return F.layer_norm(x.to(torch.float32), (x.shape[-1],), w.to(torch.float32), b.to(torch.float32), eps)


def gelu(x):
    return F.gelu(x)  # exact (erf) GELU, nn.GELU() default / activation="gelu"


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)  # CLIPM:162-164


def mlp_ln_gelu(x, P, prefix, rnd=None):
    """nn.Sequential(Linear, LayerNorm, GELU, Linear) (FF:139-143,148-152,157-161; POL:83-111)."""
    h = linear(x, P[prefix + ".0.weight"], P[prefix + ".0.bias"], rnd)
    h = layer_norm(h, P[prefix + ".1.weight"], P[prefix + ".1.bias"], 1e-5)
    h = gelu(h)
    return linear(h, P[prefix + ".3.weight"], P[prefix + ".3.bias"], rnd)


def mha_varlen(x, w_in, b_in, w_out, b_out, n_head, seq_lens, rnd=None, causal=False):
    """Multi-head self-attention over packed sequences x [T, d] (block-diagonal by `seq_lens`)."""
    r = rnd or _id
    T, d = x.shape
    hd = d // n_head
    qkv = r(linear(x, w_in, b_in, rnd))  # engine stores QKV in 16 bit
    q, k, v = qkv.split(d, dim=-1)
    out = torch.empty(T, d, dtype=torch.float32)
    s = 0
    scale = 1.0 / math.sqrt(hd)
    for n in seq_lens:
        qs = q[s:s + n].view(n, n_head, hd).transpose(0, 1)
        ks = k[s:s + n].view(n, n_head, hd).transpose(0, 1)
        vs = v[s:s + n].view(n, n_head, hd).transpose(0, 1)
        att = (qs @ ks.transpose(1, 2)) * scale
        if causal:
            att = att.masked_fill(torch.ones(n, n, dtype=torch.bool).triu(1), float("-inf"))
        att = torch.softmax(att, dim=-1)
        out[s:s + n] = (att @ vs).transpose(0, 1).reshape(n, d)
        s += n
    out = r(out)  # attention output stored in 16 bit (A operand of out-proj)
    return linear(out, w_out, b_out, rnd)


def post_norm_encoder(x, P, prefix, seq_lens, n_head=12, n_layers=2, final_eps=1e-12, rnd=None):
    """nn.TransformerEncoder(TransformerEncoderLayer(d, nhead, 4d, activation='gelu', batch_first=True),
    num_layers=2, norm=LayerNorm(eps=1e-12)) in eval mode (FF:134-137,146,155): post-norm layers,
    layer-norm eps 1e-5, final norm eps 1e-12; x is a packed [T, d] batch of independent sequences."""
    for l in range(n_layers):
        p = f"{prefix}.layers.{l}."
        sa = mha_varlen(x, P[p + "self_attn.in_proj_weight"], P[p + "self_attn.in_proj_bias"],
                        P[p + "self_attn.out_proj.weight"], P[p + "self_attn.out_proj.bias"], n_head, seq_lens, rnd)
        x = layer_norm(x + sa, P[p + "norm1.weight"], P[p + "norm1.bias"], 1e-5)
        h = gelu(linear(x, P[p + "linear1.weight"], P[p + "linear1.bias"], rnd))
        h = linear(h, P[p + "linear2.weight"], P[p + "linear2.bias"], rnd)
        x = layer_norm(x + h, P[p + "norm2.weight"], P[p + "norm2.bias"], 1e-5)
    return layer_norm(x, P[prefix + ".norm.weight"], P[prefix + ".norm.bias"], final_eps)


# ------------------------------------------------------------------------------------------------
# CLIP ViT (CLIPM:153-238) -- restated functionally on an OpenAI-layout state dict
# ------------------------------------------------------------------------------------------------
def vit_forward(images, P, n_layers, n_head, patch=14, rnd=None, ln_post_on_patches=True, n_layers_run=None,
                return_hidden=False):
    """images [N,3,R,R] fp32 (already normalised).  Returns (cls [N,out], patch [N,g*g,out]).

    `ln_post_on_patches=False` is the Pretrain variant (Q8, src_3dff/models/encoders/clip/model.py:231-236).
    `n_layers_run` / `return_hidden` give the hidden state after k blocks (LLaVA's vision_feature_layer=-2).
    """
    r = rnd or _id
    N = images.shape[0]
    w = P["conv1.weight"].to(torch.float32)  # [width,3,p,p]
    width = w.shape[0]
    g = images.shape[-1] // patch
    cols = F.unfold(images.to(torch.float32), kernel_size=patch, stride=patch).transpose(1, 2)  # [N, g*g, 3*p*p]
    x = r(cols) @ r(w.reshape(width, -1)).t()
    cls = P["class_embedding"].to(torch.float32).expand(N, 1, width)
    x = torch.cat([cls, x], dim=1) + P["positional_embedding"].to(torch.float32)
    x = layer_norm(x, P["ln_pre.weight"], P["ln_pre.bias"], 1e-5)
    T = g * g + 1
    x = x.reshape(N * T, width)
    run = n_layers if n_layers_run is None else n_layers_run
    for l in range(run):
        p = f"transformer.resblocks.{l}."
        h = layer_norm(x, P[p + "ln_1.weight"], P[p + "ln_1.bias"], 1e-5)
        x = x + mha_varlen(h, P[p + "attn.in_proj_weight"], P[p + "attn.in_proj_bias"],
                           P[p + "attn.out_proj.weight"], P[p + "attn.out_proj.bias"], n_head, [T] * N, rnd)
        h = layer_norm(x, P[p + "ln_2.weight"], P[p + "ln_2.bias"], 1e-5)
        h = r(quick_gelu(linear(h, P[p + "mlp.c_fc.weight"], P[p + "mlp.c_fc.bias"], rnd)))
        x = x + linear(h, P[p + "mlp.c_proj.weight"], P[p + "mlp.c_proj.bias"], rnd)
    x = x.reshape(N, T, width)
    if return_hidden:
        return x
    proj = P["proj"].to(torch.float32)  # [width, out]
    x_cls = linear(layer_norm(x[:, 0], P["ln_post.weight"], P["ln_post.bias"], 1e-5), proj.t(), None, rnd)
    xp = x[:, 1:]
    if ln_post_on_patches:
        xp = layer_norm(xp, P["ln_post.weight"], P["ln_post.bias"], 1e-5)
    x_patch = linear(xp, proj.t(), None, rnd)
    return x_cls, x_patch


def clip_preprocess(img_u8, R=336, rnd=None):
    """CLIPEncoder.forward preprocessing (ENC:267-284) on a uint8 NHWC batch -> normalised fp32 [N,3,R,R].
    torchvision 0.14 (the reference's pin) resizes uint8 tensors in float32 WITHOUT antialias and rounds back to uint8."""
    r = rnd or _id
    x = torch.as_tensor(img_u8).permute(0, 3, 1, 2).to(torch.float32)
    if x.shape[-1] != R or x.shape[-2] != R:
        x = F.interpolate(x, size=(R, R), mode="bicubic", align_corners=False).round().clamp(0, 255)
    x = x / 255.0
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    return r((x - mean) / std)  # `image.type(self.dtype)`: fp16 (CLIPM:338-339)


# ------------------------------------------------------------------------------------------------
# llava-phi-3-mini language model prefill (HF LlamaForCausalLM / Phi3ForCausalLM math; call site POL:463)
# pinned against transformers' LlamaForCausalLM on a small config in tests/test_oracle_lm.py
# ------------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    x = x.to(torch.float32)
    return w.to(torch.float32) * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def rope_tables(positions, head_dim, theta=10000.0):
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    freqs = positions.to(torch.float32)[:, None] * inv_freq[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def apply_rope(x, cos, sin):
    """x [T, H, Dh]; HF rotate_half convention."""
    half = x.shape[-1] // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], dim=-1)
    return x * cos[:, None, :] + rot * sin[:, None, :]


def lm_prefill(embeds, seq_lens, P, n_layers, n_heads, eps=1e-5, theta=10000.0, rnd=None, prefix="model."):
    """embeds [T, D] fp32 packed sequences -> logits of each sequence's LAST token [B, vocab] (Llama-layout names:
    {prefix}layers.N.self_attn.{q,k,v,o}_proj.weight, mlp.{gate,up,down}_proj.weight, input_layernorm, post_attention_layernorm,
    {prefix}norm.weight, lm_head.weight)."""
    r = rnd or _id
    x = embeds.to(torch.float32)
    T, D = x.shape
    hd = D // n_heads
    pos = torch.cat([torch.arange(n) for n in seq_lens])
    cos, sin = rope_tables(pos, hd, theta)
    scale = 1.0 / math.sqrt(hd)
    for l in range(n_layers):
        p = f"{prefix}layers.{l}."
        h = rms_norm(x, P[p + "input_layernorm.weight"], eps)
        q = r(linear(h, P[p + "self_attn.q_proj.weight"], None, rnd)).view(T, n_heads, hd)
        k = r(linear(h, P[p + "self_attn.k_proj.weight"], None, rnd)).view(T, n_heads, hd)
        v = r(linear(h, P[p + "self_attn.v_proj.weight"], None, rnd)).view(T, n_heads, hd)
        q, k = r(apply_rope(q, cos, sin)), r(apply_rope(k, cos, sin))
        att = torch.empty(T, n_heads, hd)
        s = 0
        for n in seq_lens:
            qs, ks, vs = (t[s:s + n].transpose(0, 1) for t in (q, k, v))
            a = (qs @ ks.transpose(1, 2)) * scale
            a = a.masked_fill(torch.ones(n, n, dtype=torch.bool).triu(1), float("-inf"))
            att[s:s + n] = (torch.softmax(a, dim=-1) @ vs).transpose(0, 1)
            s += n
        x = x + linear(r(att.reshape(T, D)), P[p + "self_attn.o_proj.weight"], None, rnd)
        h = rms_norm(x, P[p + "post_attention_layernorm.weight"], eps)
        g = linear(h, P[p + "mlp.gate_proj.weight"], None, rnd)
        u = linear(h, P[p + "mlp.up_proj.weight"], None, rnd)
        x = x + linear(r(F.silu(g) * u), P[p + "mlp.down_proj.weight"], None, rnd)
    last = torch.tensor([sum(seq_lens[:i + 1]) - 1 for i in range(len(seq_lens))])
    h = rms_norm(x[last], P[prefix + "norm.weight"], eps)
    return linear(h, P["lm_head.weight"], None, rnd)


def lm_teacher_forced_logits(embeds, seq_lens, tokens, P, n_layers, n_heads, rnd=None, prefix="model.", **kw):
    """Next-token logits after feeding `tokens[b][:s]` for every s (no cache: each step re-runs the full prefill over
    [prompt || generated], the arithmetic HF `generate` performs with use_cache=False; POL:463).  tokens: list of B id lists (same length n).
    Returns [n + 1, B, vocab]: entry s = logits that choose token s."""
    r = rnd or _id
    table = P[prefix + "embed_tokens.weight"].to(torch.float32)
    parts, s0 = [], 0
    for n in seq_lens:
        parts.append(embeds[s0:s0 + n].to(torch.float32))
        s0 += n
    n_new = len(tokens[0])
    out = []
    for s in range(n_new + 1):
        cur = [torch.cat([parts[b], r(table[torch.tensor(tokens[b][:s], dtype=torch.long)])], 0) if s else parts[b] for b in range(len(seq_lens))]
        out.append(lm_prefill(torch.cat(cur, 0), [c.shape[0] for c in cur], P, n_layers, n_heads, rnd=rnd, prefix=prefix, **kw))
    return torch.stack(out, 0)


def lm_greedy_decode(embeds, seq_lens, P, n_layers, n_heads, max_new_tokens=20, eos_ids=(), rnd=None, prefix="model.", **kw):
    """HF greedy search restated without a cache; returns B id lists (ending with the EOS id when one was produced).  Only meaningful where the arg-max margins exceed the
    arithmetic noise of the implementation under test (tests use lm_teacher_forced_logits for the numeric comparison)."""
    B = len(seq_lens)
    outs, done = [[] for _ in range(B)], [False] * B
    for _ in range(max_new_tokens):
        n = max(len(o) for o in outs)
        padded = [o + [0] * (n - len(o)) for o in outs]  # finished sequences keep a dummy continuation (their logits are ignored)
        lg = lm_teacher_forced_logits(embeds, seq_lens, padded, P, n_layers, n_heads, rnd=rnd, prefix=prefix, **kw)[-1] if n else \
            lm_prefill(embeds, seq_lens, P, n_layers, n_heads, rnd=rnd, prefix=prefix, **kw)
        for b in range(B):
            if done[b]:
                continue
            t = int(lg[b].argmax())
            outs[b].append(t)  # HF keeps the EOS id as the last token
            if t in eos_ids:
                done[b] = True
        if all(done):
            break
    return outs


# ------------------------------------------------------------------------------------------------
# HF CLIPImageProcessor (transformers 4.46 "slow" / PIL path) -- the LLaVA tower's input path: POL:438 `llava_processor(images=rgb)`.
# Third-party arithmetic (Pillow's ImagingResample, 8 bits per channel; transformers' rescale / normalize): restated from the published
# algorithm and PINNED in tests/test_image_processor.py against PIL.Image.resize and transformers' own PIL-backed CLIPImageProcessorPil.
# ------------------------------------------------------------------------------------------------
def _pil_bicubic_filter(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_resample_tables(in_size, out_size, support=2.0, filt=_pil_bicubic_filter, precision_bits=22):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` (src/libImaging/Resample.c): per output index the first source index, the tap
    count and the fixed-point taps.  Returns (bounds int32 [out,2], kk int32 [out,ksize])."""
    import numpy as np
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size  # box = (0, 0, w, h) as C floats
    filterscale = max(scale, 1.0)
    sup = support * filterscale
    ksize = int(math.ceil(sup)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - sup + 0.5), 0)   # C (int) cast: truncation
        xmax = min(int(center + sup + 0.5), in_size) - xmin
        k = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(k[i] for i in range(xmax)) if xmax else 0.0
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            kk[xx, x] = int(-0.5 + w * (1 << precision_bits)) if w < 0 else int(0.5 + w * (1 << precision_bits))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_bicubic_resize_u8(img_u8, out_h, out_w):
    """Image.resize((out_w, out_h), BICUBIC) of uint8 [N,H,W,C] images: horizontal pass, uint8 intermediate, vertical pass (ImagingResample)."""
    import numpy as np
    img = np.asarray(img_u8)
    N, H, W, C = img.shape
    PB = 22

    def one_pass(src, axis, out_size):
        bounds, kk = pil_resample_tables(src.shape[axis], out_size)
        src = np.moveaxis(src, axis, 0).astype(np.int64)  # [in, ...]
        out = np.empty((out_size,) + src.shape[1:], np.uint8)
        for xx in range(out_size):
            x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
            acc = np.full(src.shape[1:], 1 << (PB - 1), np.int64)
            for t in range(n):
                acc += src[x0 + t] * int(kk[xx, t])
            out[xx] = np.clip(acc >> PB, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)
    x = img
    if W != out_w:
        x = one_pass(x, 2, out_w)
    if H != out_h:
        x = one_pass(x, 1, out_h)
    return x


def hf_clip_image_process(img_u8, R=336, rnd=None):
    """CLIPImageProcessor(size={'shortest_edge': R}, crop_size=R, resample=BICUBIC) on square uint8 NHWC images, followed by the cast of
    POL:438 (`.to(device, torch.float16)`): PIL resize -> * (1/255) in float64 -> float32 -> (x - mean) / std in float32 -> fp16.
    Returns normalised [N,3,R,R] fp32 (fp16-rounded when rnd is given)."""
    import numpy as np
    r = rnd or _id
    img = np.asarray(img_u8)
    assert img.shape[1] == img.shape[2], "square inputs (shortest-edge resize + centre crop is the identity crop then)"
    if img.shape[1] != R:
        img = pil_bicubic_resize_u8(img, R, R)
    x = (img.astype(np.float64) * (1 / 255)).astype(np.float32)
    mean = np.array([0.48145466, 0.4578275, 0.40821073], np.float32)
    std = np.array([0.26862954, 0.26130258, 0.27577711], np.float32)
    x = ((x - mean) / std).transpose(0, 3, 1, 2)
    return r(torch.from_numpy(np.ascontiguousarray(x)))
