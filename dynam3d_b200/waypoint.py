"""Candidate-waypoint predictor on the GPU: the reference's `BinaryDistPredictor_TRM` (TRM_net.py:9-88: Linear+ReLU on
the depth embedding of each of the 12 views, a 2-layer BERT whose attention sees only the neighbouring views, a 2-layer classifier -> 120 x 12
heat map) and the heat-map post-processing of `get_candidate_waypoints` (POL:226-270: softmax, wrap, NMS, indices -> angles / distances / views).

The depth encoder (VlnResnetDepthEncoder, POL:137-143, ENC:15-109) is NOT built: its output `[B*12, 128, 4, 4]` is the input, like FastSAM's
masks are for the 3D memory.  All arithmetic is fp32-class (split fp16 operands on the tensor cores), because the result is an ARG-MAX over
the heat map: the candidates must be the reference's, not within a tolerance of them.

Weights live under the reference's key names (`visual_fc_depth.1.*`, `waypoint_TRM.bert.encoder.layer.{0,1}.*`, `vis_classifier.{0,2}.*`), so
`predictor.load_state_dict(torch.load(WAYPOINT_PREDICTOR_CKPT)["predictor"]["state_dict"])`-style loading works unchanged.
"""
import math

import numpy as np
import torch

from . import _lib as L
from . import ops
from . import precise as PR
from .weight_store import WeightStore

NUM_ANGLES, NUM_IMGS, NUM_CLASSES, HEATMAP_OFFSET = 120, 12, 12, 5   # TRM:15-20
HIDDEN, HEADS, LAYERS, LN_EPS = 768, 12, 2, 1e-12


def neighbor_mask(num_imgs=NUM_IMGS, neighbor=1):
    """utils.py:90-102 as the additive mask of WBERT:184-185 ((1 - mask) * -10000), fp32 [num_imgs, num_imgs]."""
    mask = np.zeros((num_imgs, num_imgs), dtype=np.float32)
    t = np.zeros(num_imgs, dtype=np.float32)
    t[:neighbor + 1] = 1
    if neighbor != 0:
        t[-neighbor:] = 1
    for ri in range(num_imgs):
        mask[ri] = t
        t = np.roll(t, 1)
    return (1.0 - mask) * -10000.0


class WaypointPredictor(WeightStore):
    """forward(depth_embedding [B*12, 128, 4, 4] or [B*12, 2048], cuda fp32) -> heat-map logits [B, 120, 12] (TRM:66-88);
    candidates(logits) -> the per-episode lists of POL:253-270."""

    def __init__(self, device="cuda"):
        super().__init__(device=device)
        self.device = torch.device(device)
        self._built_version = -1
        self._w = None

    def _weights(self):
        if self._w is not None and self._built_version == self.version:
            return self._w
        g = lambda k: self.get(k).detach().to(self.device, torch.float32).contiguous()
        w = {"fc": (g("visual_fc_depth.1.weight"), g("visual_fc_depth.1.bias")), "layers": [],
             "c0": (g("vis_classifier.0.weight"), g("vis_classifier.0.bias")), "c2": (g("vis_classifier.2.weight"), g("vis_classifier.2.bias")),
             "mask": torch.from_numpy(neighbor_mask()).to(self.device)}
        for l in range(LAYERS):
            p = f"waypoint_TRM.bert.encoder.layer.{l}."
            w["layers"].append({
                # query / key / value as ONE [3 * 768, 768] matrix: a single GEMM per layer (WBERT:57-59)
                "qkv": (torch.cat([g(p + f"attention.self.{n}.weight") for n in ("query", "key", "value")], 0).contiguous(),
                        torch.cat([g(p + f"attention.self.{n}.bias") for n in ("query", "key", "value")], 0).contiguous()),
                "ao": (g(p + "attention.output.dense.weight"), g(p + "attention.output.dense.bias")),
                "ln1": (g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias")),
                "i": (g(p + "intermediate.dense.weight"), g(p + "intermediate.dense.bias")),
                "o": (g(p + "output.dense.weight"), g(p + "output.dense.bias")),
                "ln2": (g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"))})
        self._w, self._built_version = w, self.version
        return w

    @staticmethod
    def _relu(x):
        L.check(L.lib().d3d_wp_relu(L.ptr(x), x.numel(), L.stream_ptr()))
        return x

    def forward(self, depth_embedding):
        L.lib()  # raises if the CUDA extension is not built: there is no CPU path
        w = self._weights()
        with L.stream_scope():
            x = depth_embedding.to(self.device, torch.float32).reshape(depth_embedding.shape[0], -1).contiguous()   # nn.Flatten (TRM:28)
            T = x.shape[0]
            assert T % NUM_IMGS == 0 and x.shape[1] == 2048, x.shape
            B = T // NUM_IMGS
            h = self._relu(PR.linear(x, w["fc"][0], w["fc"][1]))                                                    # TRM:27-31
            for lw in w["layers"]:
                qkv = PR.linear(h, lw["qkv"][0], lw["qkv"][1]).contiguous()
                ctx = torch.empty((T, HIDDEN), device=self.device, dtype=torch.float32)
                L.check(L.lib().d3d_wp_neighbor_attention(L.ptr(qkv), L.ptr(w["mask"]), B, NUM_IMGS, HEADS, HIDDEN // HEADS,
                                                          1.0 / math.sqrt(HIDDEN // HEADS), L.ptr(ctx), L.stream_ptr()))     # WBERT:62-84
                a = PR.linear(ctx, lw["ao"][0], lw["ao"][1], residual=h).contiguous()                               # BertSelfOutput: dense + residual
                ops.layernorm(a, lw["ln1"][0], lw["ln1"][1], LN_EPS, out32=a)
                inter = PR.linear(a, lw["i"][0], lw["i"][1], act=L.ACT_GELU)                                        # BertIntermediate (erf GELU)
                h = PR.linear(inter.contiguous(), lw["o"][0], lw["o"][1], residual=a).contiguous()                  # BertOutput
                ops.layernorm(h, lw["ln2"][0], lw["ln2"][1], LN_EPS, out32=h)
            c = self._relu(PR.linear(h, w["c0"][0], w["c0"][1]).contiguous())                                       # TRM:60-64
            lg = PR.linear(c, w["c2"][0], w["c2"][1]).contiguous().reshape(B, NUM_ANGLES, NUM_CLASSES)              # TRM:80-81
            return torch.cat([lg[:, HEATMAP_OFFSET:], lg[:, :HEATMAP_OFFSET]], 1).contiguous()                      # TRM:84-86

    def heatmap_nms(self, logits, max_predictions=5, sigma=(7.0, 5.0)):
        """POL:226-247 on the device: (softmax probabilities [B, 120, 12], NMS map [B, 120, 12])."""
        lg = logits.to(self.device, torch.float32).contiguous()
        B = lg.shape[0]
        prob, nms = torch.empty_like(lg), torch.empty_like(lg)
        with L.stream_scope():
            L.check(L.lib().d3d_wp_heatmap_nms(L.ptr(lg), B, NUM_ANGLES, NUM_CLASSES, int(max_predictions), float(sigma[0]), float(sigma[1]),
                                               L.ptr(prob), L.ptr(nms), L.stream_ptr()))
        return prob, nms

    def candidates(self, logits, max_predictions=5, sigma=(7.0, 5.0)):
        """POL:226-270: per episode the candidate angle / distance indices and what the policy derives from them (angles counter-clockwise in
        rad, distances in m, the view each candidate falls into, the clockwise angle features)."""
        _, nms = self.heatmap_nms(logits, max_predictions, sigma)
        maps = nms.cpu().numpy()          # one small D2H: [B, 120, 12] fp32
        out = []
        for j in range(maps.shape[0]):
            nz = np.argwhere(maps[j] != 0)
            angle_idxes, distance_idxes = nz[:, 0], nz[:, 1]
            rad_c = angle_idxes.astype(np.float32) / np.float32(120) * np.float32(2 * math.pi)                      # clockwise (POL:258)
            rad_cc = np.float32(2 * math.pi) - angle_idxes.astype(np.float32) / np.float32(120) * np.float32(2 * math.pi)
            img_idxes = 12 - (angle_idxes + 5) // 10                                                                # POL:264-265
            img_idxes[img_idxes == 12] = 0
            fts = np.stack([np.sin(rad_c), np.cos(rad_c), np.sin(np.zeros_like(rad_c)), np.cos(np.zeros_like(rad_c))], 0).astype(np.float32).T
            out.append({"angle_idxes": angle_idxes, "distance_idxes": distance_idxes, "cand_angles": rad_cc.tolist(),
                        "cand_distances": ((distance_idxes + 1) * 0.25).tolist(), "cand_img_idxes": img_idxes, "cand_angle_fts": fts})
        return out
