"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's dynamic 3D token memory
(`Feature_Fields`, Dynam3D_VLN/vlnce_baselines/models/feature_fields.py = FF), habitat branch.

Parity status: PINNED -- `oracle/make_golden.py` / `tests/test_oracle_vs_reference.py` run the unmodified
reference class (through oracle/ref_shim.py, CPU) next to this restatement on multi-step seeded
trajectories and require identical discrete state (patch ids, instance ids + member lists, zone keys/ids,
K-NN indices, merge decisions) and fp32-close features.

The quirks listed in SURVEY.md section 8 (Q2, Q3, Q5, Q6, Q7, Q9) are reproduced literally.
Differences by design (documented in DESIGN.md): centroids are accumulated in fp64 and rounded once to
fp32 (order independent, so a parallel reduction can match bit for bit); the reference's torch fp32
`mean` differs from that in the last ulp at most.
"""
import math

import numpy as np
import torch

from . import geometry as G
from . import nn_ops as NN

F32 = np.float32


def mean_f64(x):
    """Order-independent centroid: fp64 accumulate, one rounding to fp32.  Empty set -> NaN (Q5)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 3)
    if len(x) == 0:
        return np.full((3,), np.nan, dtype=F32)
    return (x.sum(axis=0) / len(x)).astype(F32)


def lowest_free_ids(used_keys, n):
    """FF:433-475: the n lowest non-negative integers not in `used_keys`."""
    if len(used_keys) == 0:
        return np.arange(n, dtype=np.int64)
    out = []
    for i in range(len(used_keys) + n):
        if len(out) == n:
            break
        if i in used_keys:
            continue
        out.append(i)
    return np.array(out, dtype=np.int64)


class EpisodeState:
    def __init__(self):
        self.patch_pos = np.zeros((0, 3), F32)
        self.patch_fts = np.zeros((0, 768), np.float16)
        self.patch_dir = np.zeros((0,), F32)
        self.patch_scale = np.zeros((0,), F32)
        self.p2i = {}
        self.i2p = {}
        self.inst_pos = np.zeros((0, 3), F32)
        self.inst_fts = np.zeros((0, 768), F32)
        self.zone_pos = np.zeros((0, 3), F32)
        self.zone_fts = np.zeros((0, 768), F32)
        self.zone_key_to_id = {}
        self.z2i = {}
        self.tree = False  # FF:179 `instance_tree[b] != []`
        # trace of the last processed view (for parity tests)
        self.last_knn = None


class FeatureFieldsOracle:
    """Same public surface as the reference class for the habitat branch (FF:119-862)."""

    def __init__(self, params, batch_size=1, rnd=None, hfov=90.0, vfov=90.0, q7_fix=False):
        self.P = {k: v.detach().to(torch.float32) for k, v in params.items()}
        self.rnd = rnd
        self.hfov, self.vfov = hfov, vfov
        self.far = 3.0
        self.num_proposal = 2
        self.zone_len = 2.0
        self.q7_fix = q7_fix
        self.reset(batch_size)

    # ---- FF:186-233 ----
    def reset(self, batch_size=1):
        self.batch_size = batch_size
        self.eps = [EpisodeState() for _ in range(batch_size)]

    def pop(self, index):
        self.batch_size -= 1
        self.eps.pop(index)

    def initialize_camera_setting(self, hfov, vfov):
        self.hfov, self.vfov = hfov, vfov

    # ---- neural blocks ----
    def _encode_patches(self, pos, fts16, direction, scale, centre):
        """FF:582-595 / 662-686: 7-d position feature -> MLP, + CLIP feature, prepend token, 2-layer encoder, token 0.
        Returns the packed sequence [n+1, 768] before the encoder (caller batches sequences)."""
        pos = torch.from_numpy(np.ascontiguousarray(pos))
        c = torch.from_numpy(np.ascontiguousarray(centre))
        rel = pos - c
        dist = torch.sqrt(torch.square(pos).sum(-1, keepdim=True))  # Q6: absolute position norm
        d = torch.from_numpy(np.ascontiguousarray(direction)).unsqueeze(-1)
        s = torch.from_numpy(np.ascontiguousarray(scale)).unsqueeze(-1)
        feat7 = torch.cat([rel, dist, torch.sin(d), torch.cos(d), s], dim=-1)
        emb = torch.from_numpy(np.ascontiguousarray(fts16)).to(torch.float32) + NN.mlp_ln_gelu(
            feat7, self.P, "patch_to_instance_position_embedding", self.rnd)
        return torch.cat([self.P["aggregate_patch_to_instance_embedding"], emb], dim=0)

    def _run_encoder(self, seqs, prefix):
        if len(seqs) == 0:
            return torch.zeros((0, 768))
        lens = [len(s) for s in seqs]
        x = torch.cat(seqs, dim=0)
        y = NN.post_norm_encoder(x, self.P, prefix, lens, rnd=self.rnd)
        starts = np.cumsum([0] + lens[:-1])
        return y[torch.from_numpy(starts.astype(np.int64))]

    def _zone_sequence(self, member_pos, member_fts, zone_pos):
        """FF:719-727 / 746-753: 4-d feature [pos - zone_pos, |pos|] -> MLP, + instance feature, prepend token."""
        mp = torch.from_numpy(np.ascontiguousarray(member_pos, dtype=F32)).reshape(-1, 3)
        zp = torch.from_numpy(np.ascontiguousarray(zone_pos, dtype=F32)).reshape(1, 3)
        feat4 = torch.cat([mp - zp, torch.sqrt(torch.square(mp).sum(-1, keepdim=True))], dim=-1)
        emb = torch.from_numpy(np.ascontiguousarray(member_fts, dtype=F32)).reshape(-1, 768) + NN.mlp_ln_gelu(
            feat4, self.P, "instance_to_zone_position_embedding", self.rnd)
        return torch.cat([self.P["aggregate_instance_to_zone_embedding"], emb], dim=0)

    def _discriminate(self, prop_fts, view_fts, delta):
        """FF:613-621 -> merge_target [G,K] int (argmax of 2 logits, first max wins) and the logits."""
        x = torch.cat([prop_fts, view_fts, delta], dim=-1)
        logits = NN.mlp_ln_gelu(x, self.P, "instance_merge_discriminator", self.rnd)
        return (logits[..., 1] > logits[..., 0]).to(torch.int64).numpy(), logits.numpy()

    # ---- FF:329-396 ----
    def delete_old_features_from_camera_frustum(self, batch_depth, batch_position, batch_heading, num_of_views=1):
        """batch_depth [B,V,H,W] metres (POL:350)."""
        batch_depth = np.asarray(batch_depth, dtype=F32)
        for b in range(self.batch_size):
            ep = self.eps[b]
            cam = G.habitat_to_internal(batch_position[b])
            for ix in range(num_of_views):
                if len(ep.patch_pos) == 0:
                    continue
                heading = float(batch_heading[b])
                if self.q7_fix:
                    heading = ix * (-math.pi / 6) + heading
                mask = G.frustum_mask_habitat(ep.patch_pos, batch_depth[b, ix], cam, heading, self.hfov, self.vfov, far=self.far)
                self._apply_cull(ep, mask)
            ep.tree = len(ep.inst_pos) > 0

    def _apply_cull(self, ep, mask):
        ep.patch_pos[mask] = -10000.0
        ep.patch_fts[mask] = 0
        ep.patch_dir[mask] = 0
        ep.patch_scale[mask] = 0
        for pid in np.nonzero(mask)[0].tolist():  # Q2: array index used as patch id
            if pid not in ep.p2i:
                continue
            iid = ep.p2i.pop(pid)
            ep.i2p[iid] = ep.i2p[iid][ep.i2p[iid] != pid]
            if len(ep.i2p[iid]) == 0:
                ep.i2p.pop(iid)
                key = tuple(G.zone_keys(ep.inst_pos[iid], self.zone_len).tolist())
                ep.inst_pos[iid] = -10000.0
                ep.inst_fts[iid] = 0
                if key in ep.zone_key_to_id:
                    zid = ep.zone_key_to_id[key]
                    ep.z2i[zid] = ep.z2i[zid][ep.z2i[zid] != iid]
                    if len(ep.z2i[zid]) == 0:
                        ep.zone_key_to_id.pop(key)
                        ep.z2i.pop(zid)
                        ep.zone_pos[zid] = -10000.0
                        ep.zone_fts[zid] = 0

    # ---- FF:493-815 (habitat branch) ----
    def update_feature_fields(self, batch_depth, batch_grid_ft, batch_patch_segm, batch_position, batch_heading, num_of_views=1):
        """batch_depth [B,V,576] metres; batch_grid_ft [B,V,576,768]; batch_patch_segm [B,V,24,24] dense int labels."""
        for b in range(self.batch_size):
            ep = self.eps[b]
            for ix in range(num_of_views):
                self._update_view(ep, np.asarray(batch_depth[b][ix], F32), np.asarray(batch_grid_ft[b][ix]).astype(np.float16),
                                  np.asarray(batch_patch_segm[b][ix]).reshape(-1), batch_position[b], float(batch_heading[b]), ix)

    # ---- posed-dataset branch (FF:343-344, 501-546): same state machine, geometry from intrinsics / poses ----
    def delete_old_features_posed(self, batch_depth, batch_intrinsic, batch_extrinsic):
        """batch_depth[b] [V,H,W] fp32; batch_intrinsic[b][ix] [>=3,>=3]; batch_extrinsic[b][ix] [4,4] world->camera (far = 2.0, FF:64)."""
        for b in range(self.batch_size):
            ep = self.eps[b]
            for ix in range(len(batch_depth[b])):
                if len(ep.patch_pos) == 0:
                    continue
                mask = G.frustum_mask_matrix(ep.patch_pos, np.asarray(batch_depth[b][ix], F32), batch_intrinsic[b][ix], batch_extrinsic[b][ix])
                self._apply_cull(ep, mask)
            ep.tree = len(ep.inst_pos) > 0

    def update_feature_fields_posed(self, batch_depth, batch_grid_ft, batch_patch_segm, batch_intrinsic, batch_rot, batch_trans,
                                    depth_scale=1000.0, depth_trunc=1000.0, ray_distance=3.0):
        fx0 = float(np.asarray(batch_intrinsic[0][0])[0][0])  # get_rays(batch_camera_intrinsic[0][0]) (FF:503)
        for b in range(self.batch_size):
            ep = self.eps[b]
            for ix in range(len(batch_depth[b])):
                xyz, direction, scale = G.unproject_posed_view(batch_depth[b][ix], batch_intrinsic[b][ix], batch_rot[b][ix], batch_trans[b][ix],
                                                               depth_scale, depth_trunc, ray_distance=ray_distance, ray_fx=fx0)
                self._update_view(ep, None, np.asarray(batch_grid_ft[b][ix]).astype(np.float16), np.asarray(batch_patch_segm[b][ix]).reshape(-1),
                                  None, None, ix, geom=(xyz, direction, scale))

    def _update_view(self, ep, depth576, grid_ft16, segm, position, heading, ix, geom=None):
        proposal_num = min(len(ep.i2p), self.num_proposal)
        if geom is not None:
            xyz, direction, scale = geom
        else:
            xyz, direction, scale = G.unproject_view_world(depth576, position, heading, ix, self.hfov, self.vfov)
        ep.patch_pos = np.concatenate([ep.patch_pos, xyz], 0)
        ep.patch_dir = np.concatenate([ep.patch_dir, direction], 0)
        ep.patch_scale = np.concatenate([ep.patch_scale, scale], 0)
        ep.patch_fts = np.concatenate([ep.patch_fts, grid_ft16], 0)

        seg_ids = np.unique(segm).tolist()
        n_seg = len(seg_ids)
        centres = np.stack([mean_f64(xyz[segm == s]) for s in seg_ids], 0)
        seqs = [self._encode_patches(xyz[segm == s], grid_ft16[segm == s], direction[segm == s], scale[segm == s], centres[j])
                for j, s in enumerate(seg_ids)]
        view_fts = self._run_encoder(seqs, "aggregate_patch_to_instance_encoder").numpy()

        if ep.tree:
            d2, idx = G.knn3d(ep.inst_pos, centres, proposal_num)
            ep.last_raw = {"centres": centres.copy(), "d2": d2.copy(), "idx": idx.copy()}
            if float(d2.astype(np.float64).sum()) > 1e6:  # Q9
                col = d2.astype(np.float64).sum(0)
                proposal_num = int((col < 1e6).sum())
                d2, idx = G.knn3d(ep.inst_pos, centres, proposal_num)
            ep.last_knn = (d2.copy(), idx.copy())
            K = proposal_num
            if K > 0:
                prop_pos = ep.inst_pos[idx]  # [G,K,3]
                delta = torch.from_numpy(centres[:, None, :] - prop_pos)
                prop_fts = torch.from_numpy(ep.inst_fts[idx])
                vf = torch.from_numpy(view_fts)[:, None, :].repeat(1, K, 1)
                merge_target, logits = self._discriminate(prop_fts, vf, delta)
            else:
                merge_target = np.zeros((n_seg, 0), np.int64)
                logits = np.zeros((n_seg, 0, 2), F32)
            ep.last_merge = (merge_target.copy(), logits.copy())

            is_new = merge_target.sum(-1) == 0
            new_ids = lowest_free_ids(ep.i2p, int(is_new.sum())) if is_new.any() else None
            patch_ids = lowest_free_ids(ep.p2i, len(segm))
            new_ix = 0
            merged = []  # instance ids that received a merge, in order
            for g in range(n_seg):
                members = patch_ids[segm == g]
                if is_new[g]:
                    iid = int(new_ids[new_ix])
                    new_ix += 1
                    ep.i2p[iid] = members
                    for pid in members.tolist():
                        ep.p2i[pid] = iid
                    if iid < len(ep.inst_pos):
                        ep.inst_pos[iid] = centres[g]
                        ep.inst_fts[iid] = view_fts[g]
                    else:
                        ep.inst_pos = np.concatenate([ep.inst_pos, centres[g:g + 1]], 0)
                        ep.inst_fts = np.concatenate([ep.inst_fts, view_fts[g:g + 1]], 0)
                else:
                    j = int(np.nonzero(merge_target[g])[0][0])  # nearest accepted proposal only
                    iid = int(idx[g, j])
                    ep.i2p[iid] = np.concatenate([ep.i2p[iid], members], 0)
                    for pid in members.tolist():
                        ep.p2i[pid] = iid
                    ids = ep.i2p[iid]  # Q2: ids index the patch arrays directly
                    ep.inst_pos[iid] = mean_f64(ep.patch_pos[ids])
                    seq = self._encode_patches(ep.patch_pos[ids], ep.patch_fts[ids], ep.patch_dir[ids], ep.patch_scale[ids], ep.inst_pos[iid])
                    ep.inst_fts[iid] = self._run_encoder([seq], "aggregate_patch_to_instance_encoder").numpy()[0]
                    merged.append(iid)

            # zones (FF:693-756)
            slot_keys = G.zone_keys(ep.inst_pos, self.zone_len)
            view_keys = G.zone_keys(centres, self.zone_len)
            uniq = np.unique(view_keys, axis=0)
            zone_ids = lowest_free_ids(ep.z2i, len(uniq))
            zi = 0
            for key_arr in uniq:
                key = tuple(key_arr.tolist())
                m = (slot_keys[:, 0] == key_arr[0]) & (slot_keys[:, 1] == key_arr[1]) & (slot_keys[:, 2] == key_arr[2])
                members = np.arange(len(m))[m]
                if key not in ep.zone_key_to_id:
                    zid = int(zone_ids[zi])
                    zi += 1
                    ep.zone_key_to_id[key] = zid
                    ep.z2i[zid] = members
                    zpos = mean_f64(ep.inst_pos[m])
                    seq = self._zone_sequence(ep.inst_pos[m], ep.inst_fts[m], zpos)
                    zft = self._run_encoder([seq], "aggregate_instance_to_zone_encoder").numpy()
                    ep.zone_pos = np.concatenate([ep.zone_pos, zpos[None]], 0)  # Q3: always appended
                    ep.zone_fts = np.concatenate([ep.zone_fts, zft], 0)
                else:
                    zid = ep.zone_key_to_id[key]
                    ep.z2i[zid] = members
                    ep.zone_pos[zid] = mean_f64(slot_keys[m])  # Q5: mean of voxel-centre keys
                    seq = self._zone_sequence(slot_keys[m], ep.inst_fts[m], ep.zone_pos[zid])
                    ep.zone_fts[zid] = self._run_encoder([seq], "aggregate_instance_to_zone_encoder").numpy()[0]
        else:
            # first view of the episode (FF:759-812)
            ep.last_raw = {"centres": centres.copy(), "d2": np.zeros((n_seg, 0), F32), "idx": np.zeros((n_seg, 0), np.int32)}
            ep.last_knn = None
            ep.last_merge = None
            ep.inst_pos = centres.copy()
            ep.inst_fts = view_fts.copy()
            new_ids = lowest_free_ids(ep.i2p, n_seg)
            patch_ids = lowest_free_ids(ep.p2i, len(segm))
            for s in seg_ids:
                members = patch_ids[segm == s]
                iid = int(new_ids[s])
                ep.i2p[iid] = members
                for pid in members.tolist():
                    ep.p2i[pid] = iid
            view_keys = G.zone_keys(centres, self.zone_len)
            uniq = np.unique(view_keys, axis=0)
            zone_ids = lowest_free_ids(ep.z2i, len(uniq))
            zpos_l, zft_l = [], []
            for zi, key_arr in enumerate(uniq):
                key = tuple(key_arr.tolist())
                m = (view_keys[:, 0] == key_arr[0]) & (view_keys[:, 1] == key_arr[1]) & (view_keys[:, 2] == key_arr[2])
                zid = int(zone_ids[zi])
                ep.zone_key_to_id[key] = zid
                ep.z2i[zid] = np.arange(len(m))[m]
                zpos = mean_f64(centres[m])
                zpos_l.append(zpos[None])
                seq = self._zone_sequence(centres[m], view_fts[m], zpos)
                zft_l.append(self._run_encoder([seq], "aggregate_instance_to_zone_encoder").numpy())
            ep.zone_pos = np.concatenate([ep.zone_pos] + zpos_l, 0)
            ep.zone_fts = np.concatenate([ep.zone_fts] + zft_l, 0)
        ep.tree = len(ep.inst_pos) > 0

    # ---- FF:818-862 ----
    def get_environment_features(self, agent_position, agent_heading_angle, instance_distance=5.0, zone_distance=100.0):
        out = {"batch_instance_fts": [], "batch_instance_relative_position": [], "batch_zone_fts": [], "batch_zone_relative_position": []}
        for b in range(self.batch_size):
            ep = self.eps[b]
            ids = np.array(list(ep.i2p.keys()), dtype=np.int64)
            rel, dist = G.to_agent_frame(ep.inst_pos[ids], agent_position[b], agent_heading_angle[b])
            keep = dist <= F32(instance_distance)
            out["batch_instance_relative_position"].append(rel[keep])
            out["batch_instance_fts"].append(ep.inst_fts[ids][keep])
            zids = np.array(list(ep.z2i.keys()), dtype=np.int64)
            rel, dist = G.to_agent_frame(ep.zone_pos[zids], agent_position[b], agent_heading_angle[b])
            keep = dist <= F32(zone_distance)
            out["batch_zone_relative_position"].append(rel[keep])
            out["batch_zone_fts"].append(ep.zone_fts[zids][keep])
        return out

    def get_patch_3d_info(self, batch_depth_map):
        return G.patch_3d_info(batch_depth_map, self.hfov, self.vfov)

    # ---- discrete-state snapshot used by the parity tests ----
    def snapshot(self, b=0):
        ep = self.eps[b]
        return {
            "n_patches": len(ep.patch_pos),
            "p2i": dict(ep.p2i),
            "i2p": {k: v.copy() for k, v in ep.i2p.items()},
            "i2p_order": list(ep.i2p.keys()),
            "n_inst_slots": len(ep.inst_pos),
            "zone_key_to_id": dict(ep.zone_key_to_id),
            "z2i": {k: v.copy() for k, v in ep.z2i.items()},
            "z2i_order": list(ep.z2i.keys()),
            "n_zone_slots": len(ep.zone_pos),
            "patch_tomb": (ep.patch_pos[:, 0] == -10000.0).copy(),
        }
