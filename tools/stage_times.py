"""Per-stage wall/device times of one bench step (sync after each stage) -- a profiling aid, not a bench value."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from dynam3d_b200 import _lib as L, ops, synth  # noqa: E402


def main():
    E = int(os.environ.get("EPISODES", bench.EPISODES_PER_GPU))
    net = bench.build_engine(E)
    steps = bench.make_inputs(0, 5, E)
    instr = [synth.make_instruction(b) for b in range(E)]
    ff = net.feature_fields
    V, P = bench.VIEWS, 576
    rows = []
    for i, s in enumerate(steps):
        obs = {"rgb": torch.from_numpy(s["rgb"]).cuda(), "depth": torch.from_numpy(s["depth"]).cuda(), "patch_segm": s["segm"]}
        torch.cuda.synchronize()
        T = {}

        def lap(name, t0):
            torch.cuda.synchronize()
            T[name] = (time.perf_counter() - t0) * 1000
            return time.perf_counter()
        c0 = L.N_CALLS
        t = time.perf_counter()
        depth = obs["depth"].reshape(E * V, bench.DEPTH, bench.DEPTH).contiguous()
        d576 = ops.depth_patch_grid(depth, E, V, literal_q1=False)
        full = ops.depth_preprocess(depth).view(E, V, bench.DEPTH, bench.DEPTH)
        t = lap("depth", t)
        _, grid = net.rgb_encoder({"rgb": obs["rgb"]})
        t = lap("clip_vit_96img", t)
        ff.delete_old_features_from_camera_frustum(full, s["pos"], s["head"], num_of_views=V)
        t = lap("cull", t)
        ff.update_feature_fields(d576.view(E, V, P), grid.reshape(E, V, P, 768), batch_position=s["pos"], batch_heading=s["head"], num_of_views=V,
                                 batch_patch_segm=s["segm"])
        t = lap("ff_update_12views", t)
        env = ff.get_environment_features(s["pos"], s["head"])
        t = lap("export", t)
        c1 = L.N_CALLS
        ff_calls = c1 - c0
        # the rest through the public path on a fresh copy is not possible (state already updated) -> time the remaining pieces directly
        sel = torch.arange(E, device="cuda") * V
        patch = net.llava.image_features(obs["rgb"][sel].contiguous())
        t = lap("llava_tower_8img", t)
        PW = net._policy_weights()
        inst = [net._project_tokens(env["batch_instance_fts"][b], env["batch_instance_relative_position"][b], PW["inst_pos"], PW["inst_proj"]) for b in range(E)]
        zone = [net._project_tokens(env["batch_zone_fts"][b], env["batch_zone_relative_position"][b], PW["zone_pos"], PW["zone_proj"]) for b in range(E)]
        t = lap("projections", t)
        lens = [2 + 576 + inst[b].shape[0] + zone[b].shape[0] + 150 for b in range(E)]
        X = torch.randn(sum(lens), 3072, device="cuda")
        cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
        pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
        torch.cuda.synchronize()
        t = time.perf_counter()
        net.llava.lm.prefill(X, cu, pos, E, max(lens), (cu[1:] - 1).int().contiguous())
        t = lap("phi3_prefill", t)
        T["tokens"] = sum(lens)
        T["calls_to_export"] = ff_calls
        T["n_inst"] = [ep.n_inst for ep in ff.eps][:3]
        rows.append(T)
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in T.items()}), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/stage_times.json", "w"), indent=1)


if __name__ == "__main__":
    main()
