"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline): a CPU restatement (numpy, fp32) of the reference's
candidate-waypoint predictor and heat-map post-processing -- SURVEY.md 8(f) rank 3.

PINNED: `tests/test_oracle_vs_reference.py::test_waypoint_oracle_matches_reference` runs the reference's own `BinaryDistPredictor_TRM`
(Dynam3D_VLN/vlnce_baselines/waypoint_pred/TRM_net.py) and `nms` (waypoint_pred/utils.py) in the build container through
`oracle/ref_shim.load_reference_waypoint_predictor()` on the same weights / inputs; `tests/golden/waypoint.npz` carries the reference outputs to
the GPU box.

File tags: TRM = waypoint_pred/TRM_net.py, WBERT = waypoint_pred/transformer/waypoint_bert.py, WU = waypoint_pred/utils.py,
POL = models/Policy_Dynam3D_VLN.py, MU = models/utils.py.  The depth encoder (VlnResnetDepthEncoder, POL:137-143) is NOT restated: its
output [B*12, 128, 4, 4] is the input here.
"""
import math

import numpy as np

NUM_ANGLES, NUM_IMGS, NUM_CLASSES, HEATMAP_OFFSET = 120, 12, 12, 5  # TRM:15-20, POL:192-195
HIDDEN, HEADS, LAYERS, LN_EPS = 768, 12, 2, 1e-12                  # TRM:38-45 (BertConfig defaults: 12 heads, layer_norm_eps 1e-12)
F32 = np.float32


def attention_mask(num_imgs=NUM_IMGS, neighbor=1):
    """WU:90-102: view i may attend to itself and its `neighbor` circular neighbours on each side; returned as 0 / 1 [num_imgs, num_imgs]."""
    mask = np.zeros((num_imgs, num_imgs))
    t = np.zeros(num_imgs)
    t[:neighbor + 1] = 1
    if neighbor != 0:
        t[-neighbor:] = 1
    for ri in range(num_imgs):
        mask[ri] = t
        t = np.roll(t, 1)
    return mask


def additive_mask(num_imgs=NUM_IMGS, neighbor=1):
    """WBERT:184-185: (1 - mask) * -10000, added to the scaled scores."""
    return ((1.0 - attention_mask(num_imgs, neighbor)) * -10000.0).astype(F32)


def _linear(x, w, b):
    return (x @ w.T.astype(F32) + b.astype(F32)).astype(F32)


def _layernorm(x, g, b, eps=LN_EPS):
    """TRM:91-106 (TF-style: epsilon inside the square root) == torch LayerNorm."""
    u = x.mean(-1, keepdims=True, dtype=F32)
    s = ((x - u) ** 2).mean(-1, keepdims=True, dtype=F32)
    return (g * ((x - u) / np.sqrt(s + F32(eps))) + b).astype(F32)


def _gelu(x):
    """BERT `gelu` (modeling_bert.py: x * 0.5 * (1 + erf(x / sqrt(2))))."""
    from math import erf
    v = np.vectorize(erf, otypes=[np.float64])(x.astype(np.float64) / math.sqrt(2.0))
    return (x.astype(np.float64) * 0.5 * (1.0 + v)).astype(F32)


def predictor_logits(sd, depth_embedding):
    """TRM:66-88 (depth path; rgb_feats is None in POL:211): sd = state dict of BinaryDistPredictor_TRM as numpy fp32, depth_embedding
    [B*12, 128, 4, 4] -> heat-map logits [B, 120, 12] (already rolled by HEATMAP_OFFSET)."""
    g = lambda k: np.asarray(sd[k], dtype=F32)
    x = np.asarray(depth_embedding, dtype=F32).reshape(depth_embedding.shape[0], -1)          # nn.Flatten (TRM:28)
    bsi = x.shape[0] // NUM_IMGS
    h = np.maximum(_linear(x, g("visual_fc_depth.1.weight"), g("visual_fc_depth.1.bias")), 0)  # TRM:27-31
    add = additive_mask()
    dh = HIDDEN // HEADS
    for l in range(LAYERS):
        p = f"waypoint_TRM.bert.encoder.layer.{l}."
        q = _linear(h, g(p + "attention.self.query.weight"), g(p + "attention.self.query.bias"))
        k = _linear(h, g(p + "attention.self.key.weight"), g(p + "attention.self.key.bias"))
        v = _linear(h, g(p + "attention.self.value.weight"), g(p + "attention.self.value.bias"))
        ctx = np.empty_like(q)
        for e in range(bsi):                                                                     # WBERT:62-84
            r = slice(e * NUM_IMGS, (e + 1) * NUM_IMGS)
            qe, ke, ve = (t[r].reshape(NUM_IMGS, HEADS, dh).transpose(1, 0, 2) for t in (q, k, v))
            sc = (qe @ ke.transpose(0, 2, 1)) / F32(math.sqrt(dh)) + add[None]
            sc = sc - sc.max(-1, keepdims=True)
            pr = np.exp(sc, dtype=F32)
            pr = pr / pr.sum(-1, keepdims=True, dtype=F32)
            ctx[r] = (pr @ ve).transpose(1, 0, 2).reshape(NUM_IMGS, HIDDEN)
        a = _linear(ctx, g(p + "attention.output.dense.weight"), g(p + "attention.output.dense.bias"))
        h1 = _layernorm(a + h, g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias"))      # BertSelfOutput
        inter = _gelu(_linear(h1, g(p + "intermediate.dense.weight"), g(p + "intermediate.dense.bias")))                  # BertIntermediate
        o = _linear(inter, g(p + "output.dense.weight"), g(p + "output.dense.bias"))
        h = _layernorm(o + h1, g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"))                          # BertOutput
    c = np.maximum(_linear(h, g("vis_classifier.0.weight"), g("vis_classifier.0.bias")), 0)      # TRM:60-64
    lg = _linear(c, g("vis_classifier.2.weight"), g("vis_classifier.2.bias"))                    # [B*12, 120]
    lg = lg.reshape(bsi, NUM_ANGLES, NUM_CLASSES)                                                # TRM:80-81
    return np.concatenate([lg[:, HEATMAP_OFFSET:], lg[:, :HEATMAP_OFFSET]], 1)                   # TRM:84-86


def heatmap_nms(logits, max_predictions=5, sigma=(7.0, 5.0)):
    """POL:226-247 + WU:37-66 on [B, 120, 12] logits -> (prob [B,120,12], nms map [B,120,12]).  Literal details kept: the map is wrapped by one
    angle row on both sides; the suppression box is centred at (x = ix % W, y = ix / W as a FLOAT: true division, WU:55); `circular_x`
    wraps the CLASS axis (WU:23-24); an exhausted map keeps returning index 0 (torch.max: first maximum)."""
    lg = np.asarray(logits, dtype=F32)
    B = lg.shape[0]
    flat = lg.reshape(B, -1)
    ex = np.exp(flat - flat.max(1, keepdims=True), dtype=F32)
    prob = (ex / ex.sum(1, keepdims=True, dtype=F32)).astype(F32).reshape(B, NUM_ANGLES, NUM_CLASSES)
    wrap = np.concatenate([prob[:, -1:], prob, prob[:, :1]], 1)                                  # [B, 122, 12]
    H, W = wrap.shape[1], wrap.shape[2]
    out = np.zeros_like(wrap)
    supp = wrap.copy()
    xs = np.arange(W, dtype=F32)[None, :]
    ys = np.arange(H, dtype=F32)[:, None]
    for _ in range(max_predictions):
        for b in range(B):
            ix = int(np.argmax(supp[b].reshape(-1)))                                             # first maximum
            out[b].reshape(-1)[ix] = wrap[b].reshape(-1)[ix]
            y_mu, x_mu = F32(ix) / F32(W), F32(ix % W)
            y_diff = ys - y_mu
            x_diff = xs - x_mu
            x_diff = np.minimum(np.abs(x_diff), np.abs(x_diff + F32(W)))
            g = ((np.abs(x_diff) <= F32(sigma[0])) & (np.abs(y_diff) <= F32(sigma[1]))).astype(F32)
            supp[b] *= (1 - g)
    out[out < 0] = 0
    return prob, out[:, 1:-1, :]


def angle_feature(headings):
    """MU:49-57."""
    h = np.asarray(headings, dtype=F32)
    return np.stack([np.sin(h), np.cos(h), np.sin(np.zeros_like(h)), np.cos(np.zeros_like(h))], 0).astype(F32).T


def candidates_from_map(nms_map):
    """POL:253-270 for one episode's [120, 12] NMS map -> dict(angle_idxes, distance_idxes, cand_angles (counter-clockwise, rad),
    cand_distances (m), cand_img_idxes, cand_angle_fts [K, 4])."""
    nz = np.argwhere(np.asarray(nms_map) != 0)          # row-major like torch.nonzero
    angle_idxes, distance_idxes = nz[:, 0], nz[:, 1]
    angle_rad_c = angle_idxes.astype(F32) / F32(120) * F32(2 * math.pi)
    angle_rad_cc = F32(2 * math.pi) - angle_idxes.astype(F32) / F32(120) * F32(2 * math.pi)
    img_idxes = 12 - (angle_idxes + 5) // 10
    img_idxes[img_idxes == 12] = 0
    return {"angle_idxes": angle_idxes, "distance_idxes": distance_idxes, "cand_angles": angle_rad_cc.tolist(),
            "cand_distances": ((distance_idxes + 1) * 0.25).tolist(), "cand_img_idxes": img_idxes, "cand_angle_fts": angle_feature(angle_rad_c)}
