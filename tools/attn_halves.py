"""A/B of the tcgen05 attention tile shapes (key halves 1 / 2) at the ViT and Phi-3 prefill shapes; timing + max error vs the other shape."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import ops, _lib as L  # noqa: E402


def run(lens, H, Dh, causal, label):
    T = sum(lens)
    qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 0.5).to(torch.float16)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
    outs = {}
    for hv in (2, 1):
        L.check(L.lib().d3d_attention_tc_set_halves(hv, hv))
        out = torch.zeros(T, H * Dh, device="cuda", dtype=torch.float16)
        for _ in range(3):
            ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=causal, impl="tc")
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(20):
            ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=causal, impl="tc")
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        fl = 4.0 * sum(n * n for n in lens) * Dh * H * (0.5 if causal else 1.0)
        outs[hv] = out
        print(f"{label} halves={hv}: {ms:.4f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
    print(f"{label} max |halves1 - halves2| = {(outs[1].float() - outs[2].float()).abs().max().item():.2e}")


run([577] * 96, 16, 64, False, "vit")
run([745] * 8, 32, 96, True, "lm")
run([735, 745, 716, 739, 745, 753, 730, 739], 32, 96, True, "lm-ragged")
