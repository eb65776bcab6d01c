"""Decode-step micro-benchmark: 8 sequences x ~740 prefill tokens, Phi-3-mini (32 layers), 20 greedy tokens (CUDA events)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dynam3d_b200 import synth  # noqa: E402
from dynam3d_b200.phi3 import LMEngine, LMWeights  # noqa: E402

layers = int(os.environ.get("LAYERS", 32))
lens = [735, 745, 716, 739, 745, 753, 730, 739]
sd = synth.lm_state_dict(7, layers=layers, device="cuda", round_to=torch.float16)
eng = LMEngine(LMWeights.from_state_dict(sd, dtype=torch.float16), n_heads=32, max_tokens=sum(lens))
del sd
emb = synth.hash_uniform((sum(lens), 3072), 5, 1.0).cuda()
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
last = (cu[1:] - 1).to(torch.int32).contiguous()
from dynam3d_b200 import _lib as L  # noqa: E402
L.lib().d3d_lm_decode_set_pdl(int(os.environ.get("PDL", 1)))
L.lib().d3d_gemm_skinny_set_config(int(os.environ.get("SKCFG", 0)))
import time  # noqa: E402
_lib = L.lib()
_orig = _lib.d3d_lm_decode_step
host_s = []


def _timed(*a):
    t = time.perf_counter()
    r = _orig(*a)
    host_s.append(time.perf_counter() - t)
    return r


_lib.d3d_lm_decode_step = _timed
res = {}
for n_new in (1, 20):
    ts = []
    for it in range(4):
        x = emb.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if n_new == 1:
            eng.prefill(x, cu, pos, len(lens), max(lens), last)
        else:
            eng.generate(x, cu, pos, len(lens), max(lens), last, max_new_tokens=n_new)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[n_new] = sorted(ts)[1]
step_ms = (res[20] - res[1]) / 19
w = eng.w
wbytes = layers * (3 * 3072 * 3072 + 3072 * 3072 + 2 * 8192 * 3072 + 3072 * 8192) * 2 + 32064 * 3072 * 2
kv_bytes = layers * sum(lens) * 2 * 3072 * 2
out = {"prefill_ms": round(res[1], 2), "generate20_ms": round(res[20], 2), "decode_step_ms": round(step_ms, 3), "weight_GB": round(wbytes / 1e9, 2),
       "kv_GB_per_step": round(kv_bytes / 1e9, 2), "hbm_GBs": round((wbytes + kv_bytes) / step_ms / 1e6, 1),
       "host_ms_first_calls": [round(1e3 * t, 3) for t in host_s[:3] + host_s[19:22]], "pdl": int(os.environ.get("PDL", 1)), "skcfg": int(os.environ.get("SKCFG", 0))}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/decode_bench_pdl%d.json" % out["pdl"], "w"))
