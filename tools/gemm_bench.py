"""Micro-benchmark of the tcgen05 GEMM at the hot path's shapes (CUDA events, L2 flushed between iterations)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from dynam3d_b200 import ops, _lib  # noqa: E402

SHAPES = [  # (M, N, K, dtype, act)
    (6924, 3072, 1024, torch.float16, 0), (6924, 1024, 1024, torch.float16, 0), (6924, 4096, 1024, torch.float16, 1),
    (6924, 1024, 4096, torch.float16, 0), (55392, 3072, 1024, torch.float16, 0), (55392, 1024, 1024, torch.float16, 0), (55392, 4096, 1024, torch.float16, 1), (55392, 1024, 4096, torch.float16, 0),
    (600, 9216, 3072, torch.bfloat16, 0), (600, 16384, 3072, torch.bfloat16, 4), (600, 3072, 8192, torch.bfloat16, 0),
    (6000, 9216, 3072, torch.float16, 0), (6000, 16384, 3072, torch.float16, 4), (6000, 3072, 8192, torch.float16, 0), (6000, 3072, 3072, torch.float16, 0),
    (4800, 9216, 3072, torch.bfloat16, 0), (4800, 16384, 3072, torch.bfloat16, 4), (4800, 3072, 8192, torch.bfloat16, 0),
    (8192, 8192, 8192, torch.bfloat16, 0),
]


def main():
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    res = []
    for M, N, K, dt, act in SHAPES:
        a = torch.randn(M, K, device="cuda").to(dt)
        w = torch.randn(N, K, device="cuda").to(dt)
        out = torch.empty(M, N // 2 if act == 4 else N, device="cuda", dtype=dt)
        for _ in range(3):
            ops.gemm(a, w, out=out, act=act)
        times, tt, t1 = [], [], []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.gemm(a, w, out=out, act=act); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            _lib.lib().d3d_gemm_set_pair_mode(0)
            flush.zero_()
            e0.record(); ops.gemm(a, w, out=out, act=act); e1.record(); torch.cuda.synchronize()
            t1.append(e0.elapsed_time(e1))
            _lib.lib().d3d_gemm_set_pair_mode(1)
            flush.zero_()
            e0.record(); torch.matmul(a, w.t()); e1.record(); torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1))
        t = sorted(times)[len(times) // 2]
        t2 = sorted(tt)[len(tt) // 2]
        t1 = sorted(t1)[len(t1) // 2]
        fl = 2.0 * M * N * K
        res.append({"M": M, "N": N, "K": K, "dtype": str(dt), "act": act, "ms": round(t, 4), "tflops": round(fl / t / 1e9, 1), "single_cta_tflops": round(fl / t1 / 1e9, 1),
                    "cublas_ms": round(t2, 4), "cublas_tflops": round(fl / t2 / 1e9, 1)})
        print(res[-1], flush=True)
    json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    main()
