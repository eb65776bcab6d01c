"""a18 timings: the Pretrain novel-view patch renderer (PFF:494-625) at the reference's ScanNet size -- 16 images = 9 216 stored patches,
144 rays x 501 samples = 72 144 K-NN queries (K = 4), 1 152 important samples through the two tinycudann MLPs -- per stage, CUDA events.

    python tools/render_bench.py [--out gpurun_out/render_bench.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--patches", type=int, default=9216)
    a = ap.parse_args()
    from dynam3d_b200 import ops, synth
    from dynam3d_b200.pretrain_render import NerfRenderer
    rng = np.random.default_rng(0)
    N = a.patches
    xyz = torch.from_numpy((rng.uniform(-4, 4, size=(N, 3)) * [1, 1, 0.35] + [0, 0, 1.2]).astype(np.float32)).cuda()
    dr = torch.from_numpy(rng.uniform(0, 6.28, size=N).astype(np.float32)).cuda()
    sc = torch.from_numpy(rng.uniform(0.01, 0.3, size=N).astype(np.float32)).cuda()
    fts = synth.hash_uniform((N, 768), 5, 0.9, device="cuda").half()
    ren = NerfRenderer(synth.nerf_state_dict(3))
    pos, head = np.array([0.0, 1.25, 0.0], np.float32), 0.4
    for _ in range(3):
        ren.render(xyz, dr, sc, fts, pos, head)
    torch.cuda.synchronize()
    # whole render
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ren.render(xyz, dr, sc, fts, pos, head); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    total = float(np.median(ts))
    # the big K-NN alone: 72 144 queries x N references, K = 4 (brute force: 8 FLOP per pair; compulsory traffic 12 N + 12 Q + 32 Q bytes)
    q = torch.from_numpy(rng.uniform(-4, 4, size=(72144, 3)).astype(np.float32)).cuda()
    ops.knn3d(xyz, q, 4)
    torch.cuda.synchronize()
    tk = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.knn3d(xyz, q, 4); e1.record(); torch.cuda.synchronize()
        tk.append(e0.elapsed_time(e1))
    knn = float(np.median(tk))
    # stage profile of one render (events around every wrapped op)
    ops.STAGE_PROFILE, ops.STAGE_TAG = [], "render"
    ren.render(xyz, dr, sc, fts, pos, head)
    torch.cuda.synchronize()
    prof, ops.STAGE_PROFILE = ops.STAGE_PROFILE, None
    gemm_ms = sum(p[3].elapsed_time(p[4]) for p in prof if p[0].endswith("gemm"))
    gemm_fl = sum(p[2] for p in prof if p[0].endswith("gemm"))
    out = {"patches": N, "queries": 72144, "K": 4, "render_ms": round(total, 3), "knn_72144xN_ms": round(knn, 3),
           "knn_gflops": round(8.0 * 72144 * N / knn / 1e6, 1), "knn_pairs_per_us": round(72144 * N / knn / 1e3, 1),
           "mlp_chain_gemm_ms": round(gemm_ms, 3), "mlp_chain_tflops": round(gemm_fl / gemm_ms / 1e9, 1),
           "note": "render = rays + 2 K-NN + top-8 + gather + Linear/LN x2 + tinycudann encoder/decoder (tcgen05 GEMMs, LeakyReLU epilogue) + volume rendering"}
    print(json.dumps(out))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
