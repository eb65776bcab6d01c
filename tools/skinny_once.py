"""A few launches of one skinny-GEMM variant (target of ncu): SHAPE=gate_up|qkv|o|down|lm_head CFG=0..9."""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import _lib as L  # noqa: E402
SH = {"qkv": (9216, 3072, 0), "o": (3072, 3072, 0), "gate_up": (16384, 3072, 4), "down": (3072, 8192, 0), "lm_head": (32064, 3072, 0)}
N, K, act = SH[os.environ.get("SHAPE", "gate_up")]
L.lib().d3d_gemm_skinny_set_config(int(os.environ.get("CFG", "8")))
ws = [(torch.randn(N, K, device="cuda") * K ** -0.5).half() for _ in range(6)]
a = (torch.randn(8, K, device="cuda") * 0.5).half()
out = torch.empty(8, N // 2 if act == 4 else N, device="cuda", dtype=torch.float16)
for w in ws:
    args = L.GemmArgs(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), out.stride(0), 8, N, K, 0, 0, None, act, None, 0)
    L.check(L.lib().d3d_gemm_skinny(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
torch.cuda.synchronize()
