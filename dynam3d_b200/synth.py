"""Seeded synthetic inputs for the Dynam3D per-step hot path.

Used by the tests, `bench.py` and `__graft_entry__.smoke()`; numpy only (no CUDA, no oracle).

* RGB: uniform uint8.
* Depth: z-depth ray-cast of an axis-aligned room + a few boxes from the agent pose, divided by 10 m,
  quantised to 1/4096 (so values are exact in fp32 on any host), ~2 % zeros (exercises the column-max fill).
* Pose: habitat axes (x right, y up, z back), y = 1.25 m; trajectory = 0.25 m forward + turns of 15..60 deg.
* Segmentation (FastSAM stand-in): dense int64 labels on the 24x24 patch grid.
* Weights: `hash_uniform` -- a counter-based generator written with exact integer tensor ops so that the
  CPU oracle and the GPU engine get bit-identical parameters without shipping checkpoints.
"""
import math

import numpy as np

ROOM = np.array([[-4.0, 4.0], [-3.0, 3.0], [0.0, 3.0]])  # internal frame X, Y(forward@0), Z(up)


def _boxes(rng, n=3):
    out = []
    for _ in range(n):
        c = np.array([rng.uniform(-3, 3), rng.uniform(-2, 2), 0.0])
        s = np.array([rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.5, 2.0)])
        out.append(np.stack([c - s * [1, 1, 0], c + s], 1))
    return out


def _ray_box_exit(o, d, box):
    """distance along d (per ray) at which a ray starting INSIDE `box` leaves it."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (box[:, 0][None] - o[None]) / d
        t2 = (box[:, 1][None] - o[None]) / d
    t = np.where(d > 0, t2, t1)
    t = np.where(d == 0, np.inf, t)
    return t.min(axis=-1)


def _ray_box_hit(o, d, box):
    """entry distance of rays starting OUTSIDE `box` (inf when missed)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (box[:, 0][None] - o[None]) / d
        t2 = (box[:, 1][None] - o[None]) / d
    tmin = np.nanmax(np.minimum(t1, t2), axis=-1)
    tmax = np.nanmin(np.maximum(t1, t2), axis=-1)
    return np.where((tmax >= tmin) & (tmin > 0), tmin, np.inf)


def render_depth(position_hab, heading, size=256, hfov=90.0, boxes=(), zero_frac=0.02, rng=None):
    """[size,size,1] fp32 depth in [0,1] (z-depth / 10 m), habitat camera at `position_hab` looking along `heading`."""
    px, py, pz = position_hab[0], -position_hab[2], position_hab[1]
    o = np.array([px, py, pz], dtype=np.float64)
    fwd = np.array([-math.sin(heading), math.cos(heading), 0.0])
    right = np.array([math.cos(heading), math.sin(heading), 0.0])
    up = np.array([0.0, 0.0, 1.0])
    t = math.tan(math.radians(hfov) / 2)
    c = ((np.arange(size) + 0.5) / (size / 2) - 1.0) * t
    r = (1.0 - (np.arange(size) + 0.5) / (size / 2)) * t
    d = fwd[None, None] + c[None, :, None] * right[None, None] + r[:, None, None] * up[None, None]
    d = d.reshape(-1, 3)
    depth = _ray_box_exit(o, d, ROOM)
    for bx in boxes:
        depth = np.minimum(depth, _ray_box_hit(o, d, bx))
    depth = np.clip(depth / 10.0, 0.0, 1.0)
    depth = np.round(depth * 4096.0) / 4096.0
    depth = depth.reshape(size, size).astype(np.float32)
    if rng is not None and zero_frac > 0:
        depth[rng.random((size, size)) < zero_frac] = 0.0
    return depth[..., None]


def make_segmentation(rng, n_seg, kind="blocks", grid=24):
    """Dense labels 0..G-1 on the patch grid, [grid, grid] int64."""
    if n_seg == 1:
        return np.zeros((grid, grid), np.int64)
    if kind == "blocks":
        shapes = {16: (6, 6), 48: (3, 4), 4: (12, 12), 36: (4, 4), 64: (3, 3), 144: (2, 2), 576: (1, 1)}
        bh, bw = shapes[n_seg]
        lab = (np.arange(grid)[:, None] // bh) * (grid // bw) + (np.arange(grid)[None, :] // bw)
        return lab.astype(np.int64)
    # voronoi: irregular segment sizes
    seeds = rng.uniform(0, grid, size=(n_seg, 2))
    yy, xx = np.meshgrid(np.arange(grid) + 0.5, np.arange(grid) + 0.5, indexing="ij")
    d = (yy[..., None] - seeds[:, 0]) ** 2 + (xx[..., None] - seeds[:, 1]) ** 2
    lab = d.argmin(-1)
    _, dense = np.unique(lab, return_inverse=True)
    return dense.reshape(grid, grid).astype(np.int64)


def make_trajectory(rng, n_steps, start=(0.0, 1.25, 0.0)):
    """positions [n,3] fp32 (habitat axes) and headings [n] python floats."""
    pos = np.array(start, dtype=np.float64)
    heading = float(rng.uniform(0, 2 * math.pi))
    P, Hd = [], []
    for _ in range(n_steps):
        P.append(pos.astype(np.float32))
        Hd.append(heading)
        heading = (heading + math.radians(float(rng.choice([-60, -45, -30, -15, 15, 30, 45, 60])))) % (2 * math.pi)
        step = np.array([-math.sin(heading) * 0.25, 0.0, -math.cos(heading) * 0.25])
        nxt = pos + step
        if abs(nxt[0]) < 3.5 and abs(nxt[2]) < 2.5:
            pos = nxt
    return np.stack(P, 0), Hd


def make_episode(seed, n_steps=1, num_views=1, rgb_size=224, depth_size=256, n_seg=16, seg_kind="blocks", pano_depth=True):
    """One synthetic episode.  Returns a list of per-step dicts:
    rgb u8 [V,rgb,rgb,3]; depth f32 [V,depth,depth,1]; segm i64 [V,24,24]; position f32 [3]; heading float."""
    rng = np.random.default_rng(seed)
    boxes = _boxes(rng)
    pos, head = make_trajectory(rng, n_steps)
    steps = []
    for t in range(n_steps):
        rgb = rng.integers(0, 256, size=(num_views, rgb_size, rgb_size, 3), dtype=np.uint8)
        depth = np.stack([render_depth(pos[t], head[t] + (ix * (-math.pi / 6) if pano_depth else 0.0), depth_size, boxes=boxes, rng=rng)
                          for ix in range(num_views)], 0)
        segm = np.stack([make_segmentation(rng, n_seg, seg_kind) for _ in range(num_views)], 0)
        steps.append({"rgb": rgb, "depth": depth, "segm": segm, "position": pos[t].copy(), "heading": float(head[t])})
    return steps


# ------------------------------------------------------------------------------------------------
# counter-based deterministic weights (torch, device agnostic, bit-identical on CPU and CUDA)
# ------------------------------------------------------------------------------------------------
def hash_uniform(shape, seed, scale=1.0, device="cpu", dtype=None):
    """uniform(-scale, scale) from a 64-bit integer mix of (seed, index); exact integer math, so CPU == CUDA."""
    import torch
    n = int(np.prod(shape)) if len(shape) else 1
    x = torch.arange(n, dtype=torch.int64, device=device)
    x = x + (int(seed) * 0x9E3779B97F4A7C15 % (1 << 63))
    x = (x ^ (x >> 30)) * 0x3F4A7C159E3779B9
    x = (x ^ (x >> 27)) * 0x14057B7EF767814F
    x = x ^ (x >> 31)
    u = ((x >> 11) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)  # 24 bits -> [0,1), exact in fp32
    w = (u * 2.0 - 1.0) * float(scale)
    w = w.reshape(shape)
    return w if dtype is None else w.to(dtype)


def _u(shape, seed, std, device="cpu"):
    return hash_uniform(shape, seed, scale=std * math.sqrt(3.0), device=device)


def vit_state_dict(seed, width=1024, layers=24, patch=14, resolution=336, out_dim=768, device="cpu", half_round=True):
    """OpenAI-CLIP-layout visual state dict with CLIP's init scales; weights rounded to fp16-representable values
    (the reference stores the tower in fp16, CLIPM:389-410)."""
    import torch
    r = (lambda t: t.to(torch.float16).to(torch.float32)) if half_round else (lambda t: t)
    s = seed * 1000
    attn_std, proj_std, fc_std = width ** -0.5, (width ** -0.5) * ((2 * layers) ** -0.5), (2 * width) ** -0.5
    tokens = (resolution // patch) ** 2 + 1
    sd = {
        "conv1.weight": r(_u((width, 3, patch, patch), s + 1, (3 * patch * patch) ** -0.5, device)),
        "class_embedding": _u((width,), s + 2, width ** -0.5, device),
        "positional_embedding": _u((tokens, width), s + 3, width ** -0.5, device),
        "ln_pre.weight": 1.0 + _u((width,), s + 4, 0.05, device), "ln_pre.bias": _u((width,), s + 5, 0.05, device),
        "ln_post.weight": 1.0 + _u((width,), s + 6, 0.05, device), "ln_post.bias": _u((width,), s + 7, 0.05, device),
        "proj": r(_u((width, out_dim), s + 8, width ** -0.5, device)),
    }
    for l in range(layers):
        p, b = f"transformer.resblocks.{l}.", s + 100 + 20 * l
        sd[p + "ln_1.weight"] = 1.0 + _u((width,), b + 1, 0.05, device)
        sd[p + "ln_1.bias"] = _u((width,), b + 2, 0.05, device)
        sd[p + "attn.in_proj_weight"] = r(_u((3 * width, width), b + 3, attn_std, device))
        sd[p + "attn.in_proj_bias"] = r(_u((3 * width,), b + 4, 0.02, device))
        sd[p + "attn.out_proj.weight"] = r(_u((width, width), b + 5, proj_std, device))
        sd[p + "attn.out_proj.bias"] = r(_u((width,), b + 6, 0.02, device))
        sd[p + "ln_2.weight"] = 1.0 + _u((width,), b + 7, 0.05, device)
        sd[p + "ln_2.bias"] = _u((width,), b + 8, 0.05, device)
        sd[p + "mlp.c_fc.weight"] = r(_u((4 * width, width), b + 9, fc_std, device))
        sd[p + "mlp.c_fc.bias"] = r(_u((4 * width,), b + 10, 0.02, device))
        sd[p + "mlp.c_proj.weight"] = r(_u((width, 4 * width), b + 11, proj_std, device))
        sd[p + "mlp.c_proj.bias"] = r(_u((width,), b + 12, 0.02, device))
    return sd


def lm_state_dict(seed, hidden=3072, layers=32, ffn=8192, vocab=32064, device="cpu", round_to=None):
    """Llama-layout state dict of the llava-phi-3-mini language model shape (random init, std 0.02), rounded to the
    16-bit storage type (`round_to` = torch.float16 / torch.bfloat16) like the checkpoint (POL:125 loads bf16)."""
    import torch
    r = (lambda t: t.to(round_to).to(torch.float32)) if round_to is not None else (lambda t: t)
    s = seed * 100000
    sd = {"model.embed_tokens.weight": r(_u((vocab, hidden), s + 1, 0.02, device)),
          "model.norm.weight": 1.0 + _u((hidden,), s + 2, 0.05, device),
          "lm_head.weight": r(_u((vocab, hidden), s + 3, 0.02, device))}
    for l in range(layers):
        p, b = f"model.layers.{l}.", s + 100 + 20 * l
        sd[p + "input_layernorm.weight"] = 1.0 + _u((hidden,), b + 1, 0.05, device)
        sd[p + "post_attention_layernorm.weight"] = 1.0 + _u((hidden,), b + 2, 0.05, device)
        for j, n in enumerate(("q", "k", "v", "o")):
            sd[p + f"self_attn.{n}_proj.weight"] = r(_u((hidden, hidden), b + 3 + j, 0.02, device))
        sd[p + "mlp.gate_proj.weight"] = r(_u((ffn, hidden), b + 8, 0.02, device))
        sd[p + "mlp.up_proj.weight"] = r(_u((ffn, hidden), b + 9, 0.02, device))
        sd[p + "mlp.down_proj.weight"] = r(_u((hidden, ffn), b + 10, 0.02, device))
    return sd


def llava_state_dict(seed, clip_layers=24, lm_layers=32, device="cpu", lm_round_to=None, clip_width=1024, lm_hidden=3072, lm_ffn=8192,
                     vocab=32064):
    """HF-llava-style state dict (vision_tower.vision_model.*, multi_modal_projector.*, language_model.*) with random init."""
    import torch
    ov = vit_state_dict(seed + 1, width=clip_width, layers=clip_layers, device=device)
    sd = {}
    p = "vision_tower.vision_model."
    sd[p + "embeddings.patch_embedding.weight"] = ov["conv1.weight"]
    sd[p + "embeddings.class_embedding"] = ov["class_embedding"]
    sd[p + "embeddings.position_embedding.weight"] = ov["positional_embedding"]
    sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"] = ov["ln_pre.weight"], ov["ln_pre.bias"]
    sd[p + "post_layernorm.weight"], sd[p + "post_layernorm.bias"] = ov["ln_post.weight"], ov["ln_post.bias"]
    for l in range(clip_layers):
        s, d = f"transformer.resblocks.{l}.", p + f"encoder.layers.{l}."
        sd[d + "layer_norm1.weight"], sd[d + "layer_norm1.bias"] = ov[s + "ln_1.weight"], ov[s + "ln_1.bias"]
        sd[d + "layer_norm2.weight"], sd[d + "layer_norm2.bias"] = ov[s + "ln_2.weight"], ov[s + "ln_2.bias"]
        w, b = ov[s + "attn.in_proj_weight"], ov[s + "attn.in_proj_bias"]
        for i, n in enumerate("qkv"):
            sd[d + f"self_attn.{n}_proj.weight"] = w[i * clip_width:(i + 1) * clip_width].clone()
            sd[d + f"self_attn.{n}_proj.bias"] = b[i * clip_width:(i + 1) * clip_width].clone()
        sd[d + "self_attn.out_proj.weight"], sd[d + "self_attn.out_proj.bias"] = ov[s + "attn.out_proj.weight"], ov[s + "attn.out_proj.bias"]
        sd[d + "mlp.fc1.weight"], sd[d + "mlp.fc1.bias"] = ov[s + "mlp.c_fc.weight"], ov[s + "mlp.c_fc.bias"]
        sd[d + "mlp.fc2.weight"], sd[d + "mlp.fc2.bias"] = ov[s + "mlp.c_proj.weight"], ov[s + "mlp.c_proj.bias"]
    r16 = lambda t: t.to(torch.float16).to(torch.float32)
    s0 = seed * 7919
    sd["multi_modal_projector.linear_1.weight"] = r16(_u((lm_hidden, clip_width), s0 + 1, clip_width ** -0.5, device))
    sd["multi_modal_projector.linear_1.bias"] = _u((lm_hidden,), s0 + 2, 0.02, device)
    sd["multi_modal_projector.linear_2.weight"] = r16(_u((lm_hidden, lm_hidden), s0 + 3, lm_hidden ** -0.5, device))
    sd["multi_modal_projector.linear_2.bias"] = _u((lm_hidden,), s0 + 4, 0.02, device)
    for k, v in lm_state_dict(seed + 2, lm_hidden, lm_layers, lm_ffn, vocab, device, lm_round_to).items():
        sd["language_model." + k] = v
    return sd


def policy_state_dict(seed, merge_bias=0.0):
    """Reference-named policy parameters (feature_fields.* + the five projection MLPs, POL:79-111), PyTorch-like init scales."""
    import torch
    import torch.nn as nn
    w = 768
    enc_layer = nn.TransformerEncoderLayer(d_model=w, nhead=12, dim_feedforward=4 * w, batch_first=True)

    def mlp(i, h, o):
        return nn.Sequential(nn.Linear(i, h), nn.LayerNorm(h), nn.GELU(), nn.Linear(h, o))
    shapes = {}
    ff = nn.ModuleDict({
        "patch_to_instance_position_embedding": mlp(7, w, w),
        "aggregate_patch_to_instance_encoder": nn.TransformerEncoder(enc_layer, 2, norm=nn.LayerNorm(w, eps=1e-12), enable_nested_tensor=False),
        "instance_to_zone_position_embedding": mlp(4, w, w),
        "aggregate_instance_to_zone_encoder": nn.TransformerEncoder(enc_layer, 2, norm=nn.LayerNorm(w, eps=1e-12), enable_nested_tensor=False),
        "instance_merge_discriminator": mlp(2 * w + 3, 4 * w, 2)})
    for k, v in ff.state_dict().items():
        shapes["feature_fields." + k] = tuple(v.shape)
    shapes["feature_fields.aggregate_patch_to_instance_embedding"] = (1, w)
    shapes["feature_fields.aggregate_instance_to_zone_embedding"] = (1, w)
    for name, m in (("patch_position_embedding", mlp(6, 4 * w, 4 * w)), ("instance_position_embedding", mlp(3, w, w)),
                    ("zone_position_embedding", mlp(3, w, w)), ("instance_projector", mlp(2 * w, 4 * w, 4 * w)),
                    ("zone_projector", mlp(2 * w, 4 * w, 4 * w))):
        for k, v in m.state_dict().items():
            shapes[name + "." + k] = tuple(v.shape)
    sd = {}
    for i, (k, shp) in enumerate(sorted(shapes.items())):
        if len(shp) >= 2:
            sd[k] = hash_uniform(shp, seed * 1000 + i, scale=shp[-1] ** -0.5)
        elif ("norm" in k and k.endswith("weight")) or k.endswith(".1.weight"):
            sd[k] = 1.0 + hash_uniform(shp, seed * 1000 + i, scale=0.05)
        else:
            sd[k] = hash_uniform(shp, seed * 1000 + i, scale=0.05)
    sd["feature_fields.instance_merge_discriminator.3.bias"] = sd["feature_fields.instance_merge_discriminator.3.bias"] + torch.tensor([0.0, merge_bias])
    return sd


class ToyTokenizer:
    """Stand-in for the llava-phi-3 tokenizer (not available offline): special tokens are single ids, other text is one id
    per character.  Only the SHAPE of the prompt matters for the hot path (ids index a random embedding table)."""
    SPECIAL = {"<|user|>": 32010, "<|end|>": 32007, "<|assistant|>": 32001, "<image>": 32038}

    _SPLIT = None

    def __call__(self, text):
        import re
        if ToyTokenizer._SPLIT is None:  # runs of "<image>" are one match: the prompt carries hundreds of them (POL:436)
            ToyTokenizer._SPLIT = re.compile("((?:<image>)+|" + "|".join(re.escape(t) for t in self.SPECIAL if t != "<image>") + ")")
        ids = [1]  # BOS like LlamaTokenizer
        for part in ToyTokenizer._SPLIT.split(text):  # leftmost special token wins at every position, as a character scan would find it
            if not part:
                continue
            if part in self.SPECIAL:
                ids.append(self.SPECIAL[part])
            elif part.startswith("<image>"):
                ids.extend([self.SPECIAL["<image>"]] * (len(part) // 7))
            else:
                ids.extend([3 + (ord(c) * 131) % 31990 for c in part])
        return ids


def make_instruction(seed, n_chars=64):
    rng = np.random.default_rng(seed)
    words = ["walk", "past", "the", "table", "turn", "left", "right", "at", "door", "stop", "near", "sofa", "exit", "hall"]
    s = ""
    while len(s) < n_chars:
        s += words[int(rng.integers(len(words)))] + " "
    return s[:n_chars]


def nerf_state_dict(seed, width=768, K=4, layers=4):
    """Reference-named parameters of the Pretrain renderer (PFF:221-254): two Linear+LayerNorm blocks and the flat tinycudann vectors."""
    import torch
    s = seed * 3000
    pad = lambda n: (n + 15) // 16 * 16
    n_enc = width * width * (layers // 2) + pad(width + 1) * width
    n_dec = width * width * (layers - layers // 2) + pad(width) * width
    return {
        "patch_to_nerf_position_embedding.0.weight": hash_uniform((width, 6), s + 1, 6 ** -0.5),
        "patch_to_nerf_position_embedding.0.bias": hash_uniform((width,), s + 2, 0.05),
        "patch_to_nerf_position_embedding.1.weight": 1.0 + hash_uniform((width,), s + 3, 0.05),
        "patch_to_nerf_position_embedding.1.bias": hash_uniform((width,), s + 4, 0.05),
        "aggregate_patch_to_nerf_encoder.0.weight": hash_uniform((width, width * K), s + 5, (width * K) ** -0.5),
        "aggregate_patch_to_nerf_encoder.0.bias": hash_uniform((width,), s + 6, 0.05),
        "aggregate_patch_to_nerf_encoder.1.weight": 1.0 + hash_uniform((width,), s + 7, 0.05),
        "aggregate_patch_to_nerf_encoder.1.bias": hash_uniform((width,), s + 8, 0.05),
        "nerf_encoder.params": hash_uniform((n_enc,), s + 9, (3.0 / width) ** 0.5).to(torch.float16).to(torch.float32),
        "nerf_decoder.params": hash_uniform((n_dec,), s + 10, (3.0 / width) ** 0.5).to(torch.float16).to(torch.float32),
    }


def waypoint_state_dict(seed, device="cpu"):
    """State dict of the reference's BinaryDistPredictor_TRM (TRM_net.py: key names and shapes) with seeded uniform weights, fp32.  LayerNorm
    weights near 1, biases small; scales chosen so that the heat-map logits spread over a few units (a peaked softmax like a trained model)."""
    H, I = 768, 3072
    s = seed * 7919
    sd = {}
    n = [0]

    def u(shape, std):
        n[0] += 1
        return _u(shape, s + n[0], std, device)

    sd["visual_fc_depth.1.weight"], sd["visual_fc_depth.1.bias"] = u((H, 2048), 2048 ** -0.5), u((H,), 0.02)
    sd["visual_merge.0.weight"], sd["visual_merge.0.bias"] = u((H, 2 * H), (2 * H) ** -0.5), u((H,), 0.02)
    for l in range(2):
        p = f"waypoint_TRM.bert.encoder.layer.{l}."
        for nm in ("query", "key", "value"):
            sd[p + f"attention.self.{nm}.weight"], sd[p + f"attention.self.{nm}.bias"] = u((H, H), 1.5 * H ** -0.5), u((H,), 0.02)
        sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"] = u((H, H), H ** -0.5), u((H,), 0.02)
        sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"] = 1.0 + u((H,), 0.05), u((H,), 0.02)
        sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"] = u((I, H), H ** -0.5), u((I,), 0.02)
        sd[p + "output.dense.weight"], sd[p + "output.dense.bias"] = u((H, I), I ** -0.5), u((H,), 0.02)
        sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"] = 1.0 + u((H,), 0.05), u((H,), 0.02)
    sd["mergefeats_LayerNorm.weight"], sd["mergefeats_LayerNorm.bias"] = 1.0 + u((H,), 0.05), u((H,), 0.02)
    sd["vis_classifier.0.weight"], sd["vis_classifier.0.bias"] = u((H, H), 2.0 * H ** -0.5), u((H,), 0.02)
    sd["vis_classifier.2.weight"], sd["vis_classifier.2.bias"] = u((120, H), 3.0 * H ** -0.5), u((120,), 0.1)
    return sd


def waypoint_depth_embedding(seed, episodes, device="cpu"):
    """Stand-in for the depth encoder's output (POL:207): [episodes * 12, 128, 4, 4], non-negative like a post-ReLU feature map."""
    import torch
    return torch.relu(hash_uniform((episodes * 12, 128, 4, 4), seed * 31 + 5, 1.5, device) + 0.25)
