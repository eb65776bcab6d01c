"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Pretrain novel-view patch renderer, habitat mode
(Dynam3D_Pretrain/src_3dff/models/feature_fields.py = PFF: get_rays_habitat 408-422, raw2feature 446-474,
patch_to_nerf_encode 477-491, render_view_3d_patch 494-625).

Parity status: PINNED against the reference's own `render_view_3d_patch`, executed here through oracle/ref_shim.py (unmodified source, CPU,
fp16 autocast) -- tests/test_oracle_vs_reference.py::test_render_oracle_matches_reference_renderer and the stored outputs in
tests/golden/render.npz: rendered unit-norm features agree to 8e-5, the selected sample positions are bit-equal on every ray that has a
neighbour within the search radius.  The one piece that stays UNPINNED is `tinycudann` itself (pinned 2.0, installed from
NVlabs/tiny-cuda-nn HEAD, environment.yml:290; CUDA-only and absent here): `tcnn.Network(CutlassMLP)` is restated -- in the shim and
below -- from its published behaviour: bias-free fully-connected layers, fp16 weights / activations with fp32 accumulation,
LeakyReLU slope 0.01, output width padded to a multiple of 16, one flat `params` vector holding the row-major [out, in]
matrices in layer order.  Everything else follows the reference source line by line, including the in-place aliasing at
PFF:598-599 (y' is computed from the already rotated x').  torch.topk's tie order is unspecified (rays without any neighbour tie over all
501 samples; every input of such a ray is masked, so its feature does not depend on the choice); we take the lowest index.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import geometry as G
from . import nn_ops as NN

F32 = np.float32


def get_rays_habitat(near=0.0, far=10.0, n_samples=501, H=12, W=12, hfov=90.0, vfov=90.0):
    rel_y = np.expand_dims(np.linspace(near, far, n_samples), axis=0).repeat(H * W, axis=0)  # float64 [HW, S]
    half_H, half_W = H // 2, W // 2
    tan_xy = np.array(([[i / half_W + 1 / W] for i in range(-half_W, half_W)]) * H, np.float32) * math.tan(np.deg2rad(hfov) / 2.0)
    rel_direction = -np.arctan(tan_xy)  # [HW,1] float32
    rel_x = rel_y * tan_xy
    tan_z = np.array([[i / half_H - 1 / H for i in range(half_H, -half_H, -1)]] * W, np.float32).T.reshape((-1, 1)) * math.tan(np.deg2rad(vfov) / 2.0)
    rel_z = rel_y * tan_z
    return (rel_x, rel_y, rel_z), rel_direction, rel_y, tan_xy.astype(F32), tan_z.astype(F32)


def tcnn_mlp(x, params, n_in, n_out, n_neurons, n_hidden, out_act, rnd):
    """tiny-cuda-nn CutlassMLP restatement (see module docstring)."""
    r = rnd or (lambda t: t)
    pad_out = (n_out + 15) // 16 * 16
    off = 0
    h = r(x)
    dims = [(n_neurons, n_in)] + [(n_neurons, n_neurons)] * (n_hidden - 1) + [(pad_out, n_neurons)]
    for li, (o, i) in enumerate(dims):
        w = params[off: off + o * i].view(o, i).to(torch.float32)
        off += o * i
        h = h @ r(w).t()
        last = li == len(dims) - 1
        if not last or out_act == "LeakyReLU":
            h = F.leaky_relu(h, 0.01)
        h = r(h)
    return h[:, :n_out]


def render_view_3d_patch(P, patch_pos, patch_dir, patch_scale, patch_fts16, position_hab, heading, rnd=None, near=0.0, far=10.0,
                         n_samples=501, n_top=8, K=4, radius=1.0, H=12, W=12, width=768, layers=4):
    r = rnd or (lambda t: t)
    pos = G.habitat_to_internal(position_hab)
    cd = float(heading)
    (rel_x, rel_y, rel_z), rel_direction, rel_dist, _, _ = get_rays_habitat(near, far, n_samples, H, W)
    ray_x = rel_x * math.cos(cd) - rel_y * math.sin(cd) + pos[0]
    ray_y = rel_x * math.sin(cd) + rel_y * math.cos(cd) + pos[1]
    ray_z = rel_z + pos[2]
    ray_xyz = np.stack([ray_x, ray_y, ray_z], -1).astype(F32)  # [HW,S,3]
    R_ = H * W
    d2, idx = G.knn3d(patch_pos, ray_xyz.reshape(-1, 3), K)
    dist = np.sqrt(d2).astype(F32)
    idx = idx.astype(np.int64)
    idx[dist >= radius] = -1
    dist[dist >= radius] = radius
    dist = dist.reshape(R_, n_samples, K)
    tmp = dist[..., 0]
    for k in range(1, K):
        tmp = (tmp + dist[..., k]).astype(F32)
    dens = (F32(1.0) / tmp).astype(F32)
    topk = np.argsort(-dens, axis=1, kind="stable")[:, :n_top]  # largest first, lowest index on ties
    sample_xyz = np.take_along_axis(ray_xyz, topk[..., None], axis=1)  # [HW, n_top, 3]
    positions = sample_xyz[:, 0].copy()
    ray_dir = rel_direction[:, -1]  # [HW]
    d2b, idxb = G.knn3d(patch_pos, sample_xyz.reshape(-1, 3), K)
    idxb = idxb.astype(np.int64)
    idxb[np.sqrt(d2b).astype(F32) >= radius] = -1
    idxb = idxb.reshape(R_, n_top, K)
    pp = torch.from_numpy(np.asarray(patch_pos, F32))
    sx = torch.from_numpy(sample_xyz)
    xyzds = torch.zeros((R_, n_top, K, 6), dtype=torch.float32)
    it = torch.from_numpy(idxb)
    xyzds[..., :3] = pp[it] - sx.unsqueeze(-2)
    x = xyzds[..., 0]          # views, as in the reference (PFF:596-599): y' uses the ALREADY rotated x'
    y = xyzds[..., 1]
    xyzds[..., 0] = x * math.cos(-cd) - y * math.sin(-cd)
    xyzds[..., 1] = x * math.sin(-cd) + y * math.cos(-cd)
    xyzds[..., :3][it == -1] = far
    pdir = torch.from_numpy(np.asarray(patch_dir, F32)) - cd
    a = pdir[it] - torch.from_numpy(np.asarray(ray_dir, F32)).unsqueeze(-1).unsqueeze(-1)
    xyzds[..., 3] = torch.sin(a)
    xyzds[..., 4] = torch.cos(a)
    xyzds[..., 3:5][it == -1] = 0
    xyzds[..., 5] = torch.from_numpy(np.asarray(patch_scale, F32))[it]
    xyzds[..., 5:][it == -1] = 0
    emb = torch.from_numpy(np.asarray(patch_fts16).astype(np.float32))[it]
    emb[it == -1] = 0
    # patch_to_nerf_encode (PFF:477-491)
    emb = emb.reshape(-1, width * K)  # fp16 values already
    pe = NN.linear(xyzds.reshape(-1, 6), P["patch_to_nerf_position_embedding.0.weight"], P["patch_to_nerf_position_embedding.0.bias"], rnd)
    pe = NN.layer_norm(pe, P["patch_to_nerf_position_embedding.1.weight"], P["patch_to_nerf_position_embedding.1.bias"], 1e-12)
    pe16 = pe.reshape(-1, width * K).to(torch.float16)
    s16 = (emb.to(torch.float16) + pe16).to(torch.float32)  # fp16 + fp16 (one rounding)
    si = NN.linear(s16, P["aggregate_patch_to_nerf_encoder.0.weight"], P["aggregate_patch_to_nerf_encoder.0.bias"], rnd)
    si = NN.layer_norm(si, P["aggregate_patch_to_nerf_encoder.1.weight"], P["aggregate_patch_to_nerf_encoder.1.bias"], 1e-12)
    enc = tcnn_mlp(si, P["nerf_encoder.params"], width, width + 1, width, layers // 2, "LeakyReLU", rnd)
    density = enc[:, -1].reshape(R_, n_top)
    encoded = enc[:, :-1] + si
    out = tcnn_mlp(encoded, P["nerf_decoder.params"], width, width, width, layers - layers // 2, "None", rnd).reshape(R_, n_top, width)
    # raw2feature (PFF:446-474); rel_dist is cast to fp16 by the caller (PFF:620)
    rd = torch.from_numpy(rel_dist.astype(np.float16).astype(np.float32))
    sd = F.softplus(density)
    dists = torch.abs(rd[..., 1:] - rd[..., :-1])
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)
    tk = torch.from_numpy(topk.astype(np.int64))
    dfull = torch.zeros(rd.shape).scatter(1, tk, sd)
    alpha = 1.0 - torch.exp(-F.relu(dfull) * dists)
    weights = alpha * torch.cumprod(torch.cat([torch.ones((R_, 1)), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    sw = torch.gather(weights, 1, tk)
    fmap = torch.sum(sw[..., None] * out, -2)
    fmap = fmap / torch.clamp(torch.linalg.norm(fmap, dim=-1, keepdim=True), min=1e-7)
    depth = torch.sum(weights * rd, -1) / torch.clamp(torch.sum(weights, -1), min=1e-7)
    return {"feature_map": fmap.numpy(), "positions": positions, "depth_map": depth.numpy(), "topk": topk, "idx": idxb,
            "ray_xyz": ray_xyz, "density": density.numpy(), "xyzds": xyzds.reshape(-1, 6).numpy(), "si": si.numpy()}
