"""Debug: clock stamps of CTA 0 of the bulk skinny GEMM on an L2-resident matrix (needs `make EXTRA=-DD3D_SK_STAMPS`)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import _lib as L  # noqa: E402
SH = {"qkv": (9216, 3072, 0), "o": (3072, 3072, 0), "gate_up": (16384, 3072, 4), "down": (3072, 8192, 0)}
for shape in ("qkv", "o", "down"):
    N, K, act = SH[shape]
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    a = (torch.randn(8, K, device="cuda") * 0.5).half()
    out = torch.empty(8, N // 2 if act == 4 else N, device="cuda", dtype=torch.float16)
    args = L.GemmArgs(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), out.stride(0), 8, N, K, 0, 0, None, act, None, 0)
    for _ in range(6):
        L.check(L.lib().d3d_gemm_skinny(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
    buf = (ctypes.c_longlong * 128)()
    fn = L.lib().d3d_debug_skinny_stamps
    fn.argtypes = [ctypes.c_void_p]
    assert fn(buf) == 0
    r = np.array(buf[:])
    t0 = r[0]
    n_it = (K // 512) * (2 if shape == "qkv" else 1)
    print(shape, "consumer: after pdl_wait", r[1] - t0)
    print("  stage (before wait, after wait):", [(int(r[2 + 2 * i] - t0), int(r[3 + 2 * i] - t0)) for i in range(min(n_it, 24))])
    print("  per stage: top -> A issued -> data ready -> MMAs done -> arrived:", [(int(r[100 + i] - t0), int(r[2 + 2 * i] - t0), int(r[3 + 2 * i] - t0), int(r[76 + i] - t0), int(r[88 + i] - t0)) for i in range(min(n_it, 12))])
    print("  tile0 red/epi/bar:", [int(r[52 + i] - t0) for i in range(3)], " tile1:", [int(r[55 + i] - t0) for i in range(3)])
    print("  producer: stage issue times:", [int(r[64 + i] - t0) for i in range(min(n_it, 24))])
