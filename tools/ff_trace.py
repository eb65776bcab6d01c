"""Host-side timeline of Feature_Fields._update_view inside real bench steps (profiling aid)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from dynam3d_b200 import feature_fields as FFM, synth  # noqa: E402

E = 8
net = bench.build_engine(E, 8)
steps = bench.make_inputs(0, 6, E)
instr = [synth.make_instruction(b, 64) for b in range(E)]
for i, s in enumerate(steps):
    obs = {"rgb": torch.from_numpy(s["rgb"]).cuda(), "depth": torch.from_numpy(s["depth"]).cuda(), "patch_segm": s["segm"]}
    if i >= 3:
        FFM.TRACE = []
    torch.cuda.synchronize()
    net.forward_logits(obs, instr, s["pos"], s["head"], num_of_views=12)
    torch.cuda.synchronize()
    if i >= 3:
        t = np.array(FFM.TRACE)
        print("step", i, "per-view mean ms: launch %.3f  sync-wait %.3f  plan %.3f  post-launch %.3f | totals %.2f %.2f %.2f %.2f | new %.0f merged %.0f (%.0f tok) zones %.0f (%.0f tok)" % (
            *t[:, :4].mean(0), *t[:, :4].sum(0), *t[:, 4:].mean(0)))
