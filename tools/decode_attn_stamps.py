"""Debug (make EXTRA=-DD3D_SK_STAMPS): clock stamps of CTA (0, 0) warp 0 and globaltimer of every CTA of the staged decode attention."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import _lib as L  # noqa: E402
lens = [735, 745, 716, 739, 745, 753, 730, 739]
H, Dh, n_seq = 32, 96, 8
T = sum(lens)
rows = T + 20 * n_seq
qkv = (torch.randn(rows, 3 * H * Dh, device="cuda") * 0.5).half()
out = torch.empty(n_seq, H * Dh, device="cuda", dtype=torch.float16)
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
for _ in range(4):
    L.check(L.lib().d3d_decode_attention(L.ptr(qkv), qkv.stride(0), L.ptr(cu), n_seq, T, 3, H, Dh, 0, ctypes.c_float(Dh ** -0.5), L.ptr(out), out.stride(0), L.stream_ptr()))
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
fn = L.lib().d3d_debug_skinny_stamps
fn.argtypes = [ctypes.c_void_p]
assert fn(buf) == 0
r = np.array(buf[:]); t0 = r[0]
print("after wait:", r[1] - t0, " blocks (top, data ready, computed):", [(int(r[2 + 3 * i] - t0), int(r[3 + 3 * i] - t0), int(r[4 + 3 * i] - t0)) for i in range(6)], " loop end:", r[40] - t0, " synced:", r[41] - t0)
big = (ctypes.c_ulonglong * (8 * 320 * 4))()
fn2 = L.lib().d3d_debug_skinny_cta_times
fn2.argtypes = [ctypes.c_void_p]
assert fn2(big) == 0
t = np.array(big[:], dtype=np.int64).reshape(8, 320, 4)[0, :256]
g0 = t[:, 0].min()
print("CTA start ns: %d..%d  end: %d..%d  life median %d" % (t[:, 0].min() - g0, t[:, 0].max() - g0, t[:, 2].min() - g0, t[:, 2].max() - g0, int(np.median(t[:, 2] - t[:, 0]))))
