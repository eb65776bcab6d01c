"""Time the pooling passes' attention dispatch on the step-level shape (1536 sequences of ~37 tokens + a few long ones)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from dynam3d_b200 import ops
rng = np.random.default_rng(0)
lens = list(rng.integers(8, 64, 1500)) + [100, 130, 77, 90] * 9
H, Dh = 12, 64
T = int(sum(lens))
qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 0.7).half()
out = torch.zeros(T, H * Dh, device="cuda", dtype=torch.float16)
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
for impl in ("mma", "mixed"):
    for _ in range(3):
        ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, impl=impl)
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(20):
        ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, impl=impl)
    b.record(); torch.cuda.synchronize()
    print(impl, "T", T, "n_seq", len(lens), "ms", a.elapsed_time(b) / 20)
