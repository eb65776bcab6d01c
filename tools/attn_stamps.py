"""Debug: per-phase clock stamps of one CTA of the tcgen05 attention kernel (needs `make EXTRA=-DD3D_ATTN_STAMPS`)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import ops, _lib as L  # noqa: E402

lens = [577] * 96
H, Dh = 16, 64
T = sum(lens)
qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 0.5).to(torch.float16)
out = torch.empty(T, H * Dh, device="cuda", dtype=torch.float16)
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
for _ in range(3):
    ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=False, impl="tc")
buf = (ctypes.c_longlong * 256)()
fn = L.lib().d3d_debug_attn_stamps
fn.argtypes = [ctypes.c_void_p]
assert fn(buf) == 0
raw = np.array(buf[:])
s = raw[:64].reshape(8, 8)[:5]
mm = raw[64:128].reshape(8, 8)[:5]
t0 = s[0, 0]
print("softmax warp, per tile (cycles): stamps 0 top | 1 S ready | 2 S in regs | 3 max exchanged | 4 exps done | 5 P buffer free | 6 P stored+arrived")
for j in range(5):
    r = s[j]
    nxt = s[j + 1, 0] if j < 4 else r[6]
    print(f" tile {j}: top={r[0]-t0} waitS={r[1]-r[0]} ldS={r[2]-r[1]} max+xchg={r[3]-r[2]} resc+exps={r[4]-r[3]} waitPfree={r[5]-r[4]} store={r[6]-r[5]} iter={nxt-r[0]}")
print("MMA warp (cycles since first softmax stamp): S_j read -> S_{j+1} issued -> P_j ready -> PV_j issued")
for j in range(5):
    print(f" tile {j}: ", (mm[j, :4] - t0).tolist())
