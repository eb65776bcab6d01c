"""Tuning sweep of the decode (skinny) GEMM: HBM GB/s per variant at the four Phi-3 shapes (CUDA events, weights larger than L2 in rotation)."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import _lib as L  # noqa: E402

SHAPES = [("qkv", 9216, 3072, 0), ("o", 3072, 3072, 0), ("gate_up", 16384, 3072, 4), ("down", 3072, 8192, 0), ("lm_head", 32064, 3072, 0)]
M = 8
res = {}
for name, N, K, act in SHAPES:
    n_w = max(2, int(400e6 // (N * K * 2)) + 1)  # rotate over > 3 x L2 of distinct weights
    ws = [(torch.randn(N, K, device="cuda") * K ** -0.5).half() for _ in range(n_w)]
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    out = torch.empty(M, N // 2 if act == 4 else N, device="cuda", dtype=torch.float16)
    for cfg in range(0, 11):
        L.lib().d3d_gemm_skinny_set_config(cfg)
        ts = []
        for it in range(3 + 3 * n_w):
            w = ws[it % n_w]
            args = L.GemmArgs(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), out.stride(0), M, N, K, 0, 0, None, act, None, 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(L.lib().d3d_gemm_skinny(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
            e1.record(); torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        res[f"{name}/cfg{cfg}"] = {"us": round(t * 1e3, 1), "GBs": round(N * K * 2 / t / 1e6, 0)}
    print(name, {k.split("/")[1]: v["GBs"] for k, v in res.items() if k.startswith(name)}, flush=True)
L.lib().d3d_gemm_skinny_set_config(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/skinny_bench.json", "w"), indent=1)
