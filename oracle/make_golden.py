"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz by running the UNMODIFIED reference (via ref_shim) in this container.

    python -m oracle.make_golden

The fixtures let the CPU suite and the GPU box (where /root/reference does not exist) check the oracle AND the CUDA engine
against outputs of the reference itself:
  * ff_traj_<name>.npz   -- per-step discrete state of the reference `Feature_Fields` (patch->instance map, member lists,
                            dict orders, zone keys/ids, tombstones) + exported token tensors, with the exact inputs.
  * geometry.npz         -- project_depth_to_3d_habitat / get_patch_3d_info / frustum-mask vectors of the reference functions.
  * vit_small.npz        -- reference `VisionTransformer` (VLN and Pretrain variants) outputs on a seeded small config.
"""
import json
import os

import numpy as np
import torch

from dynam3d_b200 import synth
from . import geometry as G
from . import ref_compare as RC
from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

FF_CONFIGS = {
    "v1_seg16": dict(seed=3, n_steps=5, num_views=1, n_seg=16, seg_kind="voronoi", merge_bias=0.3, weight_seed=0),
    "v1_seg48": dict(seed=22, n_steps=6, num_views=1, n_seg=48, seg_kind="blocks", merge_bias=0.55, weight_seed=0),
    "v12_seg16": dict(seed=6, n_steps=2, num_views=12, n_seg=16, seg_kind="voronoi", merge_bias=0.3, weight_seed=0),
}


def pack_snapshot(prefix, snap, out):
    out[prefix + "n"] = np.array([snap["n_patches"], snap["n_inst_slots"], snap["n_zone_slots"]], np.int64)
    p2i = np.array(sorted(snap["p2i"].items()), np.int64).reshape(-1, 2)
    out[prefix + "p2i"] = p2i
    out[prefix + "i2p_order"] = np.array(snap["i2p_order"], np.int64)
    out[prefix + "i2p_len"] = np.array([len(snap["i2p"][k]) for k in snap["i2p_order"]], np.int64)
    out[prefix + "i2p_cat"] = np.concatenate([np.asarray(snap["i2p"][k], np.int64) for k in snap["i2p_order"]] + [np.zeros(0, np.int64)])
    out[prefix + "z2i_order"] = np.array(snap["z2i_order"], np.int64)
    out[prefix + "z2i_len"] = np.array([len(snap["z2i"][k]) for k in snap["z2i_order"]], np.int64)
    out[prefix + "z2i_cat"] = np.concatenate([np.asarray(snap["z2i"][k], np.int64) for k in snap["z2i_order"]] + [np.zeros(0, np.int64)])
    keys = sorted(snap["zone_key_to_id"].items(), key=lambda kv: kv[1])
    out[prefix + "zone_keys"] = np.array([list(k) for k, _ in keys], np.float32).reshape(-1, 3)
    out[prefix + "zone_ids"] = np.array([v for _, v in keys], np.int64)
    out[prefix + "tomb"] = np.packbits(snap["patch_tomb"])


def unpack_snapshot(prefix, z):
    n = z[prefix + "n"]
    i2p, off = {}, 0
    for k, ln in zip(z[prefix + "i2p_order"].tolist(), z[prefix + "i2p_len"].tolist()):
        i2p[k] = z[prefix + "i2p_cat"][off:off + ln]
        off += ln
    z2i, off = {}, 0
    for k, ln in zip(z[prefix + "z2i_order"].tolist(), z[prefix + "z2i_len"].tolist()):
        z2i[k] = z[prefix + "z2i_cat"][off:off + ln]
        off += ln
    return {"n_patches": int(n[0]), "n_inst_slots": int(n[1]), "n_zone_slots": int(n[2]),
            "p2i": {int(a): int(b) for a, b in z[prefix + "p2i"]}, "i2p": i2p, "i2p_order": z[prefix + "i2p_order"].tolist(),
            "z2i": z2i, "z2i_order": z[prefix + "z2i_order"].tolist(),
            "zone_key_to_id": {tuple(float(x) for x in k): int(v) for k, v in zip(z[prefix + "zone_keys"], z[prefix + "zone_ids"])},
            "patch_tomb": np.unpackbits(z[prefix + "tomb"])[: int(n[0])].astype(bool)}


def make_ff():
    for name, cfg in FF_CONFIGS.items():
        steps = RC.make_inputs(cfg["seed"], cfg["n_steps"], cfg["num_views"], cfg["n_seg"], cfg["seg_kind"])
        ff, orc, recs = RC.run_pair(**cfg, steps=steps)
        out = {"config": np.frombuffer(json.dumps(cfg).encode(), np.uint8)}
        for t, (r, st) in enumerate(zip(recs, steps)):
            assert RC.snapshots_equal(r["ref"], r["orc"]) == []
            # inputs (depth is quantised to 1/4096 by synth -> exact as uint16); grid features / weights are hash-generated
            q = np.round(st["depth"][..., 0] * 4096.0)
            assert np.array_equal((q / 4096.0).astype(np.float32), st["depth"][..., 0])
            out[f"s{t}_depth_u16"] = q.astype(np.uint16)
            out[f"s{t}_segm"] = np.asarray(st["segm"]).astype(np.uint16)
            out[f"s{t}_pose"] = np.array(list(st["position"]) + [st["heading"]], np.float64)
            pack_snapshot(f"s{t}_", r["ref"], out)
            e = r["env_ref"]
            out[f"s{t}_inst_rel"] = e["batch_instance_relative_position"][0].numpy()
            out[f"s{t}_zone_rel"] = e["batch_zone_relative_position"][0].numpy()
            out[f"s{t}_inst_fts_sum"] = e["batch_instance_fts"][0].numpy().sum(-1)  # per-token checksums keep the file small
            out[f"s{t}_zone_fts_sum"] = e["batch_zone_fts"][0].numpy().sum(-1)
            if r["knn"] is not None:
                out[f"s{t}_knn_idx"], out[f"s{t}_knn_d2"] = r["knn"][1], r["knn"][0]
                out[f"s{t}_merge"] = r["merge"][0].astype(np.uint8)
                lg = r["merge"][1]
                out[f"s{t}_margin"] = np.array([np.abs(lg[..., 1] - lg[..., 0]).min() if lg.size else 1.0])
        np.savez_compressed(os.path.join(OUT, f"ff_traj_{name}.npz"), **out)
        print("wrote", name, len(recs), "steps", "min margin", min(float(out[k][0]) for k in out if k.endswith("_margin")))


def load_ff_fixture(path):
    """-> (cfg, steps (inputs), golden per-step dicts)."""
    z = np.load(path)
    cfg = json.loads(bytes(z["config"]).decode())
    steps, gold = [], []
    for t in range(cfg["n_steps"]):
        pose = z[f"s{t}_pose"]
        st = {"depth": (z[f"s{t}_depth_u16"].astype(np.float32) / np.float32(4096.0))[..., None], "segm": z[f"s{t}_segm"].astype(np.int64),
              "position": pose[:3].astype(np.float32), "heading": float(pose[3]),
              "grid": synth.hash_uniform((1, cfg["num_views"], 576, 768), cfg["seed"] * 100 + t, 0.9).numpy().astype(np.float16)}
        steps.append(st)
        g = {"snap": unpack_snapshot(f"s{t}_", z)}
        for k in ("inst_rel", "zone_rel", "inst_fts_sum", "zone_fts_sum", "knn_idx", "knn_d2", "merge"):
            if f"s{t}_{k}" in z:
                g[k] = z[f"s{t}_{k}"]
        gold.append(g)
    return cfg, steps, gold


def make_geometry():
    mod = ref_shim.load_reference_feature_fields_module()
    ff = ref_shim.make_reference_feature_fields()
    rng = np.random.default_rng(0)
    depth = rng.uniform(0.1, 10, size=(4, 576)).astype(np.float32)
    headings = np.array([0.3, 2.0, 4.0, 6.2])
    out = {"depth": depth, "headings": headings}
    for i, h in enumerate(headings):
        rx, ry, rz, d, s = ff.project_depth_to_3d_habitat(depth[i:i + 1], float(h))
        out[f"unproj{i}"] = np.stack([rx[0], ry[0], rz[0], d, s]).astype(np.float32)
    info = ff.get_patch_3d_info(depth)
    out["info5"] = np.stack([a[..., 0].numpy() for a in info])
    pts = rng.uniform(-5, 5, size=(20000, 3)).astype(np.float32)
    dimg = (np.round(rng.uniform(0.5, 6, size=(256, 256)) * 64) / 64).astype(np.float32)
    cam = np.array([0.5, -0.25, 1.25], np.float32)
    m, dep, u, v = mod.get_frustum_mask_habitat(torch.from_numpy(pts), 256, 256, 90.0, 90.0, cam, 1.1, far=3.0)
    m = m & (dep < torch.from_numpy(dimg)[v % 256, u % 256] + 0.1)
    out.update(cull_pts=pts, cull_depth=dimg.astype(np.float16), cull_cam=cam, cull_heading=np.array([1.1]), cull_mask=np.packbits(m.numpy()))
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **out)
    print("wrote geometry")


def make_posed():
    """a4': outputs of the reference's own posed-dataset functions (project_depth_to_3d FF:50-60 through the open3d stand-in of
    ref_shim, the literal glue FF:536-546, get_heading_angle FF:250-259, get_frustum_mask FF:64-84 + z-test) on seeded inputs."""
    mod = ref_shim.load_reference_feature_fields_module()
    ff = ref_shim.make_reference_feature_fields()
    from oracle import geometry as G
    out = {}
    rng = np.random.default_rng(5)
    H, W = 120, 160
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(3):
        depth = (1500 + 900 * np.sin(xx / 19.0 + i) + 600 * np.cos(yy / 11.0) + rng.integers(0, 40, (H, W))).astype(np.uint16)
        depth[rng.integers(0, H, 20), rng.integers(0, W, 20)] = 0
        K = np.array([[140.0 + i, 0, W / 2 - 1.5], [0, 145.0, H / 2 + 0.75], [0, 0, 1]], np.float64)
        a = 0.4 * i + 0.1
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.0]])
        T = rng.uniform(-1, 1, (3, 1))
        pts, _ = mod.project_depth_to_3d(torch.from_numpy(depth.astype(np.int32)), K, 1000.0, 1000.0, 24, 24)
        points = pts.astype(np.float32)
        t = abs(np.tan(G.ray_direction0(140.0, 24, 3.0)))  # get_rays of the first image (Q14: the VLN get_rays itself cannot run)
        scale = points[:, -1] * float(t) * 2. / 24
        world = (R @ points.T + T).T
        out.update({f"depth{i}": depth, f"K{i}": K, f"R{i}": R, f"T{i}": T, f"xyz{i}": world.astype(np.float32),
                    f"dir{i}": ff.get_heading_angle(world).astype(np.float32), f"scale{i}": scale.astype(np.float32)})
    cloud = rng.uniform(-3, 3, (30000, 3)).astype(np.float32)
    M = np.eye(4)
    M[:3, :3], M[:3, 3:] = out["R1"].T, -out["R1"].T @ out["T1"]
    K4 = np.eye(4, dtype=np.float32)
    K4[:3, :3] = out["K1"].astype(np.float32)
    depth_m = (out["depth1"].astype(np.float32) / 1000.0).astype(np.float32)
    m, dep, u, v = mod.get_frustum_mask(torch.from_numpy(cloud), H, W, torch.from_numpy(K4), torch.from_numpy(M.astype(np.float32)))
    m = m & (dep < torch.from_numpy(depth_m)[v % H, u % W] + 0.1)
    out.update(cull_pts=cloud, cull_M=M.astype(np.float32), cull_mask=np.packbits(m.numpy()))
    np.savez_compressed(os.path.join(OUT, "posed.npz"), **out)
    print("wrote posed", int(m.sum()))


def make_vit():
    out = {}
    width, layers, heads, res, od = 256, 3, 4, 112, 128
    sd = synth.vit_state_dict(9, width=width, layers=layers, resolution=res, out_dim=od)
    x = synth.hash_uniform((2, 3, res, res), 77, 2.0)
    for pre in (False, True):
        mod = ref_shim.load_reference_clip_model_module(pretrain=pre)
        vit = mod.VisionTransformer(res, 14, width, layers, heads, od).eval()
        vit.load_state_dict(sd, strict=True)
        with torch.no_grad():
            c, p = vit(x)
        out["cls_pre" if pre else "cls"] = c.numpy()
        out["patch_pre" if pre else "patch"] = p.numpy()
    np.savez_compressed(os.path.join(OUT, "vit_small.npz"), **out)
    print("wrote vit_small")


def render_scene(seed=31):
    """The seeded patch cloud of the renderer fixtures (4 views of a synthetic room, a few tombstones, hash features) -- inputs are
    regenerated from seeds by the tests, only the reference's OUTPUTS are stored."""
    ep = synth.make_episode(seed, n_steps=1, num_views=4, n_seg=16)[0]
    d576 = G.depth_patch_grid(ep["depth"], 1, 4, q1_fix=True)[0]
    xyz, dr, sc = [], [], []
    for ix in range(4):
        a, b, c = G.unproject_view_world(d576[ix], ep["position"], ep["heading"], ix)
        xyz.append(a); dr.append(b); sc.append(c)
    xyz, dr, sc = np.concatenate(xyz), np.concatenate(dr), np.concatenate(sc)
    xyz[::97] = -10000.0
    fts = synth.hash_uniform((len(xyz), 768), 5, 0.9).numpy().astype(np.float16)
    return xyz, dr, sc, fts, ep["position"], ep["heading"] + 0.3, synth.nerf_state_dict(3)


def make_render():
    """a18: outputs of the reference's own `render_view_3d_patch` (PFF:494-625, habitat mode) run here through ref_shim (tinycudann replaced
    by the documented stand-in, CPU fp16 autocast): unit-norm rendered features [144,768] (stored fp16: the reference returns fp16) and the
    first important sample of every ray."""
    xyz, dr, sc, fts, pos, head, P = render_scene()
    ff = ref_shim.make_reference_pretrain_feature_fields()
    ff.load_state_dict(P, strict=False)
    f_ref, p_ref = ref_shim.reference_render_view(ff, xyz, dr, sc, fts, pos, head)
    np.savez_compressed(os.path.join(OUT, "render.npz"), feature_map=f_ref.astype(np.float16), positions=p_ref.astype(np.float32))
    print("wrote render", f_ref.shape, float(np.linalg.norm(f_ref, axis=-1).max()))


WAYPOINT_SEEDS = (3, 4, 3)  # (weights, depth embedding, episodes)


def make_waypoint():
    """8(f) rank 3: heat-map logits of the reference's own BinaryDistPredictor_TRM (TRM_net.py) and the NMS map of its `nms` (waypoint_pred/utils.py)
    applied as POL:226-247 does, on seeded weights (synth.waypoint_state_dict) and a seeded depth embedding: only the OUTPUTS are stored, the
    inputs are regenerated from the seeds."""
    import torch
    trm, utils = ref_shim.load_reference_waypoint_predictor()
    net = trm.BinaryDistPredictor_TRM(device="cpu").eval()
    net.load_state_dict(synth.waypoint_state_dict(WAYPOINT_SEEDS[0]), strict=True)
    x = synth.waypoint_depth_embedding(WAYPOINT_SEEDS[1], WAYPOINT_SEEDS[2])
    with torch.no_grad():
        lg = net(None, x)
        B = lg.shape[0]
        bx = torch.softmax(lg.reshape(B, -1), 1).reshape(B, 120, 12)
        wrap = torch.cat((bx[:, -1:], bx, bx[:, :1]), 1)
        nms_map = utils.nms(wrap.unsqueeze(1), max_predictions=5, sigma=(7.0, 5.0)).squeeze(1)[:, 1:-1, :]
    np.savez_compressed(os.path.join(OUT, "waypoint.npz"), logits=lg.numpy().astype(np.float32), nms=nms_map.numpy().astype(np.float32))
    print("wrote waypoint", tuple(lg.shape), [nms_map[b].nonzero().tolist() for b in range(B)])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    make_waypoint()
    make_render()
    make_geometry()
    make_posed()
    make_vit()
    make_ff()
