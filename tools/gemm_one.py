"""Runs a few hot-path GEMM shapes once each after warm-up (target of `ncu --set full -k regex:gemm_tcgen05`)."""
import sys

import torch

sys.path.insert(0, ".")
from dynam3d_b200 import ops  # noqa: E402

import os

SHAPES = [(55392, 4096, 1024, 1, True, False), (55392, 1024, 4096, 0, True, True), (6000, 16384, 3072, 4, False, False),
          (6000, 3072, 8192, 0, False, True)]  # M, N, K, act, bias, fp32 residual (in place)
if os.environ.get("GEMM_ONE") == "step":  # the eight dominant shapes of a bench step: ViT (96 images) and Phi-3 (8 x ~745 tokens) layers
    SHAPES = [(55392, 3072, 1024, 0, True, False), (55392, 1024, 1024, 0, True, True), (55392, 4096, 1024, 1, True, False),
              (55392, 1024, 4096, 0, True, True), (5960, 9216, 3072, 0, False, False), (5960, 3072, 3072, 0, False, True),
              (5960, 16384, 3072, 4, False, False), (5960, 3072, 8192, 0, False, True)]

bufs = []
for M, N, K, act, bias, res in SHAPES:
    a = (torch.randn(M, K, device="cuda") * 0.5).half()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.randn(N, device="cuda") if bias else None
    out = torch.zeros(M, N // 2 if act == 4 else N, device="cuda", dtype=torch.float32 if res else torch.float16)
    bufs.append((a, w, b, out, act, res))
for rep in range(3):  # 2 warm-up rounds + 1 profiled (ncu -s 8 -c 4)
    for a, w, b, out, act, res in bufs:
        ops.gemm(a, w, out=out, bias=b, act=act, residual=out if res else None)
torch.cuda.synchronize()
print("ok")
