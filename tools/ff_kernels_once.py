"""A few long-horizon memory-update steps (target of `ncu --set full -k regex:frustum_cull_batched|knn2_batched|knn_`): 8 episodes with
~23 k stored patches / ~650 instance slots each by the time the profiled steps run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dynam3d_b200 import ops, synth  # noqa: E402
from dynam3d_b200.feature_fields import Feature_Fields  # noqa: E402

E, T = 8, 44
pol = synth.policy_state_dict(7, merge_bias=-10.0)
ff = Feature_Fields(batch_size=E, device="cuda", q7_fix=True)
ff.load_state_dict({k[len("feature_fields."):]: v for k, v in pol.items() if k.startswith("feature_fields.")})
ff.reset(E)
steps = bench.make_inputs(0, T, E, views=1, rgb=8, depth=256, n_seg=17, seed0=5000)
grid = synth.hash_uniform((E, 1, 576, 768), 5, 0.9, device="cuda").half()
for t in range(T):
    s = steps[t]
    dd = torch.from_numpy(s["depth"]).cuda().reshape(E, 256, 256).contiguous()
    d576 = ops.depth_patch_grid(dd, E, 1, 24, 24, literal_q1=False)
    full = ops.depth_preprocess(dd, 0.0, 10.0).view(E, 1, 256, 256)
    if t == T - 4:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ff.delete_old_features_from_camera_frustum(full, s["pos"], s["head"], num_of_views=1)
    ff.update_feature_fields(d576.view(E, 1, 576), grid, batch_position=s["pos"], batch_heading=s["head"], num_of_views=1, batch_patch_segm=s["segm"])
    ff.get_environment_features(s["pos"], s["head"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", ff.eps[0].n_patch, ff.eps[0].n_inst)
