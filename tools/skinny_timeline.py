"""Debug (make EXTRA=-DD3D_SK_STAMPS): globaltimer of every CTA of 8 back-to-back skinny GEMM launches on an L2-resident matrix."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import _lib as L  # noqa: E402
SH = {"qkv": (9216, 3072, 0), "o": (3072, 3072, 0), "gate_up": (16384, 3072, 4), "down": (3072, 8192, 0)}
L.lib().d3d_lm_decode_set_pdl(int(os.environ.get("PDL", 1)))
for shape in ("o", "qkv"):
    N, K, act = SH[shape]
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    a = (torch.randn(8, K, device="cuda") * 0.5).half()
    out = torch.empty(8, N // 2 if act == 4 else N, device="cuda", dtype=torch.float16)
    args = L.GemmArgs(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), out.stride(0), 8, N, K, 0, 0, None, act, None, 0)
    for _ in range(16):
        L.check(L.lib().d3d_gemm_skinny(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
    buf = (ctypes.c_ulonglong * (8 * 320 * 4))()
    fn = L.lib().d3d_debug_skinny_cta_times
    fn.argtypes = [ctypes.c_void_p]
    assert fn(buf) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(8, 320, 4)[:, :296]
    order = np.argsort(t[:, 0, 0])
    t = t[order]
    t0 = t[0, :, 0].min()
    print(shape, "per launch (ns since first CTA start): first/last CTA start | first/last after-wait | first/last consumer end")
    for k in range(8):
        s, wv, e = t[k, :, 0] - t0, t[k, :, 1] - t0, t[k, :, 2] - t0
        print(f"  launch {k}: start {s.min():6d}..{s.max():6d}  waited {wv.min():6d}..{wv.max():6d}  end {e.min():6d}..{e.max():6d}   CTA life median {int(np.median(e - s))} ns, after-wait work median {int(np.median(e - wv))} ns")
