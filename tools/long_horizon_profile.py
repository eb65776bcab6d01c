"""cProfile of the host side of the long-horizon memory path (steps 200..256 of bench.py --workload long_horizon): where the per-step
host time goes once an episode holds ~4 k instance slots / 147 k patches.  Profiling aid, not a bench value."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dynam3d_b200 import ops, synth  # noqa: E402
from dynam3d_b200.feature_fields import Feature_Fields  # noqa: E402

E, T = 8, 256
pol = synth.policy_state_dict(7, merge_bias=-10.0)
ff = Feature_Fields(batch_size=E, device="cuda", q7_fix=True)
ff.load_state_dict({k[len("feature_fields."):]: v for k, v in pol.items() if k.startswith("feature_fields.")})
ff.reset(E)
steps = bench.make_inputs(0, T, E, views=1, rgb=8, depth=256, n_seg=17, seed0=5000)
grids = [synth.hash_uniform((E, 1, 576, 768), 5000 * 1000 + t % 8, 0.9, device="cuda").half() for t in range(8)]
depth_dev = [torch.from_numpy(s["depth"]).cuda().reshape(E, 256, 256).contiguous() for s in steps]


def step(t):
    s = steps[t]
    d576 = ops.depth_patch_grid(depth_dev[t], E, 1, 24, 24, literal_q1=False)
    full = ops.depth_preprocess(depth_dev[t], 0.0, 10.0).view(E, 1, 256, 256)
    ff.delete_old_features_from_camera_frustum(full, s["pos"], s["head"], num_of_views=1)
    ff.update_feature_fields(d576.view(E, 1, 576), grids[t % 8], batch_position=s["pos"], batch_heading=s["head"], num_of_views=1, batch_patch_segm=s["segm"])
    return ff.get_environment_features(s["pos"], s["head"])


for t in range(200):
    step(t)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for t in range(200, T):
    step(t)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
st.sort_stats("tottime").print_stats(18)
