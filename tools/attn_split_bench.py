"""Split-operand (precise mode) attention: tcgen05 vs mma.sync kernels at the tower and Phi-3 prefill shapes."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import precise as PR  # noqa: E402


def run(lens, H, Dh, causal, label):
    T = sum(lens)
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 1.0)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
    outs = {}
    for tc in (False, True):
        PR.SPLIT_TC = tc
        for _ in range(2):
            o = PR.attention(qkv, cu, len(lens), max(lens), H, Dh, causal)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(5):
            o = PR.attention(qkv, cu, len(lens), max(lens), H, Dh, causal)
        b.record(); torch.cuda.synchronize()
        outs[tc] = o
        print(f"{label} tcgen05={tc}: {a.elapsed_time(b) / 5:.3f} ms (incl. the hi/lo split pass)", flush=True)
    print(f"{label} max |tc - mma| = {(outs[True] - outs[False]).abs().max().item():.2e}")


run([577] * 8, 16, 64, False, "tower (8 views)")
run([745] * 8, 32, 96, True, "lm")
