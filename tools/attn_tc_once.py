"""One tcgen05 attention launch at the ViT shape (for `ncu --set full -k regex:attn_tc`)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import ops  # noqa: E402

lens = [577] * 96
H, Dh, causal = 16, 64, False
impl = sys.argv[1] if len(sys.argv) > 1 else "tc"
if len(sys.argv) > 2 and sys.argv[2] == "lm":  # Phi-3 prefill shape: 8 sequences x ~745 tokens, 32 heads x 96, causal
    lens, H, Dh, causal = [745] * 8, 32, 96, True
T = sum(lens)
qkv = (torch.randn(T, 3 * H * Dh, device="cuda") * 0.5).to(torch.float16)
out = torch.empty(T, H * Dh, device="cuda", dtype=torch.float16)
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), device="cuda", dtype=torch.int32)
for _ in range(3):
    ops.attention(qkv, out, cu, len(lens), max(lens), H, Dh, causal=causal, impl=impl)
torch.cuda.synchronize()
