"""Host wall time of the two C calls per view of the native view loop (d3d_ff_view_pre / _post) over a few bench steps: is the 12-view
memory update host-issue-bound or device-bound?  Profiling aid."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dynam3d_b200 import feature_fields as FFM, synth  # noqa: E402

E = 8
net = bench.build_engine(E, 8)
instr = [synth.make_instruction(b, 64) for b in range(E)]
steps = bench.make_inputs(0, 8, E)
for i, s in enumerate(steps):
    obs = {"rgb": torch.from_numpy(s["rgb"]).cuda(), "depth": torch.from_numpy(s["depth"]).cuda(), "patch_segm": s["segm"]}
    if i == 4:
        FFM.TRACE = []
    net.forward_logits(obs, instr, s["pos"], s["head"], num_of_views=bench.VIEWS)
torch.cuda.synchronize()
tr = np.array(FFM.TRACE, dtype=np.float64)
print("views traced", len(tr))
print("pre  ms: mean %.3f  median %.3f  (issue knn/disc + device wait + planner)" % (tr[:, 0].mean(), np.median(tr[:, 0])))
print("post ms: mean %.3f  median %.3f  (uploads + slot writes + merge-pass issue; zone pass deferred)" % (tr[:, 1].mean(), np.median(tr[:, 1])))
print("per view: new %.1f merged %.1f (tokens %.0f) zones %.1f (tokens %.0f)" % tuple(tr[:, 2:].mean(0)))
print("per 12-view step: %.2f ms in the two calls" % (tr[:, :2].sum() / (len(tr) / 12)))
