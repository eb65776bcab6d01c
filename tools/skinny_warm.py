"""ncu target: the bulk skinny GEMM on ONE weight matrix that stays L2-resident (SHAPE=qkv|o|gate_up|down), i.e. the kernel's latency chain
without HBM.  ncu --cache-control none keeps the L2 warm between replays."""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import _lib as L  # noqa: E402
SH = {"qkv": (9216, 3072, 0), "o": (3072, 3072, 0), "gate_up": (16384, 3072, 4), "down": (3072, 8192, 0)}
N, K, act = SH[os.environ.get("SHAPE", "qkv")]
w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
a = (torch.randn(8, K, device="cuda") * 0.5).half()
out = torch.empty(8, N // 2 if act == 4 else N, device="cuda", dtype=torch.float16)
args = L.GemmArgs(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(out), out.stride(0), 8, N, K, 0, 0, None, act, None, 0)
for _ in range(6):
    L.check(L.lib().d3d_gemm_skinny(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    L.check(L.lib().d3d_gemm_skinny(ctypes.cast(ctypes.byref(args), ctypes.c_void_p), L.stream_ptr()))
e1.record(); torch.cuda.synchronize()
print(os.environ.get("SHAPE", "qkv"), "us per launch (L2-warm, back to back):", e0.elapsed_time(e1) * 20)
