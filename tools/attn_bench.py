"""Attention kernel micro-benchmark at the hot path's shapes (CUDA events)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import ops  # noqa: E402

CASES = [("vit96", [577] * 96, 16, 64, False, torch.float16), ("tower8", [577] * 8, 16, 64, False, torch.float16),
         ("lm8", [750] * 8, 32, 96, True, torch.float16), ("ffenc", [37] * 128, 12, 64, False, torch.float16)]
res = []
for name, lens, H, Dh, causal, dt in CASES:
    T = sum(lens)
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 0.5).to(dt)
    out = torch.empty(T, H * Dh, device="cuda", dtype=dt)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
    flops = sum(4.0 * n * n * Dh * H * (0.5 if causal else 1.0) for n in lens)
    for impl in ("tc", "mma", "simt"):
        if impl == "tc" and Dh not in (64, 96):
            continue
        if impl == "simt" and T > 20000:
            continue
        for _ in range(3):
            ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=causal, impl=impl)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=causal, impl=impl); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        res.append({"case": name, "impl": impl, "ms": round(t, 4), "tflops": round(flops / t / 1e9, 1)})
        print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/attn_bench.json", "w"), indent=1)
