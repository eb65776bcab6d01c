"""cProfile of the host side of Feature_Fields.update_feature_fields at the bench shape (profiling aid)."""
import cProfile
import io
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from dynam3d_b200 import ops  # noqa: E402

E, V, P = 8, bench.VIEWS, 576
net = bench.build_engine(E)
ff = net.feature_fields
steps = bench.make_inputs(0, 4, E)
grids = torch.randn(E, V, P, 768, device="cuda").half()
pr = cProfile.Profile()
for i, s in enumerate(steps):
    depth = torch.from_numpy(s["depth"]).cuda().reshape(E * V, bench.DEPTH, bench.DEPTH).contiguous()
    d576 = ops.depth_patch_grid(depth, E, V, literal_q1=False)
    full = ops.depth_preprocess(depth).view(E, V, bench.DEPTH, bench.DEPTH)
    torch.cuda.synchronize()
    if i >= 2:
        pr.enable()
    t = time.perf_counter()
    ff.delete_old_features_from_camera_frustum(full, s["pos"], s["head"], num_of_views=V)
    ff.update_feature_fields(d576.view(E, V, P), grids, batch_position=s["pos"], batch_heading=s["head"], num_of_views=V, batch_patch_segm=s["segm"])
    torch.cuda.synchronize()
    if i >= 2:
        pr.disable()
    print("step", i, round((time.perf_counter() - t) * 1000, 1), "ms")
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(45)
print(st.getvalue()[:9000])
