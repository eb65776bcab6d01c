"""Pretrain novel-view patch renderer on the C ABI: `Feature_Fields.render_view_3d_patch` of
Dynam3D_Pretrain/src_3dff/models/feature_fields.py (= PFF:494-625), habitat mode -- the "KNN gather + tinycudann MLP" path.

  rays (fp64 like numpy, PFF:408-422) -> exact K-NN (d3d_knn3d replaces torch_kdtree, K = 4) over the stored patches ->
  per-ray density proxy + 8 important samples -> second K-NN -> neighbour gather (fp16 features, 6-d geometry) ->
  Linear+LN / Linear+LN / tinycudann encoder + residual / decoder as tcgen05 GEMMs (LeakyReLU in the epilogue) -> volume rendering.

Parameters use the reference's names: patch_to_nerf_position_embedding.{0,1}.*, aggregate_patch_to_nerf_encoder.{0,1}.*,
nerf_encoder.params, nerf_decoder.params (flat tinycudann vectors: row-major [out, in] matrices in layer order, output padded to 16).
"""
import math

import numpy as np
import torch

from . import _lib as L
from . import ops


class NerfRenderer:
    def __init__(self, params, device="cuda", near=0.0, far=10.0, n_samples=501, n_top=8, K=4, radius=1.0, H=12, W=12, hfov=90.0, vfov=90.0,
                 width=768, layers=4):
        L.require_device()
        self.dev = torch.device(device)
        self.near, self.far, self.S, self.n_top, self.K, self.radius, self.H, self.W, self.D = near, far, n_samples, n_top, K, radius, H, W, width
        f32 = lambda t: t.detach().to(self.dev, torch.float32).contiguous()
        h16 = lambda t: t.detach().to(self.dev, torch.float32).to(torch.float16).contiguous()
        w0 = params["patch_to_nerf_position_embedding.0.weight"].detach().float().cpu()
        self.pe_w = h16(torch.cat([w0, torch.zeros(w0.shape[0], 2)], 1))  # K 6 -> 8
        self.pe_b = f32(params["patch_to_nerf_position_embedding.0.bias"])
        self.pe_ln = (f32(params["patch_to_nerf_position_embedding.1.weight"]), f32(params["patch_to_nerf_position_embedding.1.bias"]))
        self.ag_w = h16(params["aggregate_patch_to_nerf_encoder.0.weight"])
        self.ag_b = f32(params["aggregate_patch_to_nerf_encoder.0.bias"])
        self.ag_ln = (f32(params["aggregate_patch_to_nerf_encoder.1.weight"]), f32(params["aggregate_patch_to_nerf_encoder.1.bias"]))

        def mats(flat, n_in, n_out, n_hidden):
            pad = (n_out + 15) // 16 * 16
            dims = [(width, n_in)] + [(width, width)] * (n_hidden - 1) + [(pad, width)]
            out, off = [], 0
            for o, i in dims:
                out.append(h16(flat[off: off + o * i].view(o, i)))
                off += o * i
            return out
        self.enc = mats(params["nerf_encoder.params"].detach().float().cpu(), width, width + 1, layers // 2)
        self.dec = mats(params["nerf_decoder.params"].detach().float().cpu(), width, width, layers - layers // 2)
        # ray tables with the reference's expressions (PFF:408-422)
        half_H, half_W = H // 2, W // 2
        tan_xy = np.array(([[i / half_W + 1 / W] for i in range(-half_W, half_W)]) * H, np.float32) * math.tan(np.deg2rad(hfov) / 2.0)
        tan_z = np.array([[i / half_H - 1 / H for i in range(half_H, -half_H, -1)]] * W, np.float32).T.reshape((-1, 1)) * math.tan(np.deg2rad(vfov) / 2.0)
        rel_y = np.linspace(near, far, n_samples)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.rel_y = up(rel_y)
        self.tan_x = up(tan_xy.astype(np.float32).reshape(-1))
        self.tan_z = up(tan_z.astype(np.float32).reshape(-1))
        self.ray_dir = up((-np.arctan(tan_xy)).astype(np.float32).reshape(-1))
        self.rel_dist16 = up(rel_y.astype(np.float16).astype(np.float32))  # PFF:620 casts rel_dist to fp16

    def render(self, patch_pos, patch_dir, patch_scale, patch_fts16, position_hab, heading):
        """patch_* : the episode's stored patches (device: [N,3] f32, [N] f32, [N] f32, [N,768] f16).
        Returns (feature_map [H*W,768] f32, positions [H*W,3] f32, depth_map [H*W] f32, aux dict)."""
        lib, st = L.lib(), L.stream_ptr
        R, S, K, T8, D = self.H * self.W, self.S, self.K, self.n_top, self.D
        cd = float(heading)
        cam = (float(position_hab[0]), -float(position_hab[2]), float(position_hab[1]))
        dev = self.dev
        ray_xyz = torch.empty((R * S, 3), device=dev, dtype=torch.float32)
        L.check(lib.d3d_ray_points_habitat(L.ptr(self.rel_y), L.ptr(self.tan_x), L.ptr(self.tan_z), R, S, math.cos(cd), math.sin(cd), cam[0], cam[1],
                                           cam[2], L.ptr(ray_xyz), st()))
        d2, idx = ops.knn3d(patch_pos, ray_xyz, K)
        topk = torch.empty((R, T8), device=dev, dtype=torch.int32)
        L.check(lib.d3d_ray_topk(L.ptr(d2), L.ptr(idx), R, S, K, self.radius, T8, L.ptr(topk), st()))
        sample_xyz = torch.empty((R * T8, 3), device=dev, dtype=torch.float32)
        L.check(lib.d3d_gather_samples(L.ptr(ray_xyz), L.ptr(topk), R, S, T8, L.ptr(sample_xyz), st()))
        d2b, idxb = ops.knn3d(patch_pos, sample_xyz, K)
        n_pts = R * T8
        pos_rows = torch.empty((n_pts * K, 8), device=dev, dtype=torch.float16)
        feat_rows = torch.empty((n_pts, K * D), device=dev, dtype=torch.float16)
        L.check(lib.d3d_nerf_gather(L.ptr(d2b), L.ptr(idxb), L.ptr(sample_xyz), L.ptr(patch_pos), L.ptr(patch_dir), L.ptr(patch_scale),
                                    L.ptr(patch_fts16), L.ptr(self.ray_dir), n_pts, T8, K, D, self.radius, self.far, cd, math.cos(-cd), math.sin(-cd),
                                    L.ptr(pos_rows), L.D3D_F16, L.ptr(feat_rows), st()))
        # patch_to_nerf_encode (PFF:477-491)
        pe = ops.gemm(pos_rows, self.pe_w, bias=self.pe_b, out_dtype=torch.float32)
        ops.layernorm(pe, self.pe_ln[0], self.pe_ln[1], 1e-12, out32=pe)
        a16 = torch.empty((n_pts, K * D), device=dev, dtype=torch.float16)
        L.check(lib.d3d_add_half(L.ptr(pe), L.ptr(feat_rows), L.ptr(a16), n_pts * K * D, st()))
        si = ops.gemm(a16, self.ag_w, bias=self.ag_b, out_dtype=torch.float32)
        si16 = torch.empty((n_pts, D), device=dev, dtype=torch.float16)
        ops.layernorm(si, self.ag_ln[0], self.ag_ln[1], 1e-12, out32=si, out16=si16)
        h = si16
        for i, w in enumerate(self.enc):  # tinycudann encoder: LeakyReLU on every layer incl. the output
            h = ops.gemm(h, w, act=L.ACT_LEAKY_RELU, out_dtype=torch.float16)
        enc = h.float()
        density = enc[:, D].contiguous()
        encoded = torch.empty((n_pts, D), device=dev, dtype=torch.float16)
        ops.cast16((enc[:, :D] + si).contiguous(), encoded)
        h = encoded
        for i, w in enumerate(self.dec):  # decoder: LeakyReLU on hidden layers, linear output
            h = ops.gemm(h, w, act=L.ACT_LEAKY_RELU if i < len(self.dec) - 1 else L.ACT_NONE, out_dtype=torch.float16)
        feat = h.float().contiguous()
        fmap = torch.empty((R, D), device=dev, dtype=torch.float32)
        depth = torch.empty((R,), device=dev, dtype=torch.float32)
        L.check(lib.d3d_volume_render(L.ptr(feat), L.ptr(density), L.ptr(topk), L.ptr(self.rel_dist16), R, S, T8, D, L.ptr(fmap), L.ptr(depth), st()))
        positions = sample_xyz.view(R, T8, 3)[:, 0].contiguous()
        return fmap, positions, depth, {"topk": topk, "idx": idxb.view(R, T8, K), "ray_xyz": ray_xyz.view(R, S, 3), "density": density.view(R, T8),
                                        "pos_rows": pos_rows, "si": si}
