"""In-situ kernel timeline of decode steps (CUPTI through torch.profiler: start / duration of every kernel, PDL overlap preserved)."""
import json, os, sys
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
from dynam3d_b200 import synth, _lib as L  # noqa: E402
from dynam3d_b200.phi3 import LMEngine, LMWeights  # noqa: E402

layers = int(os.environ.get("LAYERS", 32))
L.lib().d3d_lm_decode_set_pdl(int(os.environ.get("PDL", 1)))
lens = [735, 745, 716, 739, 745, 753, 730, 739]
sd = synth.lm_state_dict(7, layers=layers, device="cuda", round_to=torch.float16)
eng = LMEngine(LMWeights.from_state_dict(sd, dtype=torch.float16), n_heads=32, max_tokens=sum(lens))
del sd
emb = synth.hash_uniform((sum(lens), 3072), 5, 1.0).cuda()
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
last = (cu[1:] - 1).to(torch.int32).contiguous()
for _ in range(2):
    eng.generate(emb.clone(), cu, pos, len(lens), max(lens), last, max_new_tokens=4)
torch.cuda.synchronize()
x = emb.clone()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.generate(x, cu, pos, len(lens), max(lens), last, max_new_tokens=6)
    torch.cuda.synchronize()
ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort()
# decode steps start at the first decode_prep kernel
idx = [i for i, e in enumerate(ev) if "decode_prep" in e[2]]
a, b = idx[2], idx[3]  # third decode step
step = ev[a:b]
t0 = step[0][0]
rows = []
prev_end = t0
for s, e, n in step:
    short = n.split("::")[-1].split("(")[0][:28]
    rows.append((round(s - t0, 1), round(e - s, 1), round(s - prev_end, 1), short))
    prev_end = max(prev_end, e)
print("step wall us:", round(step[-1][1] - t0, 1), "kernels:", len(step), "sum of durations:", round(sum(r[1] for r in rows), 1))
agg = {}
for st, d, gap, n in rows:
    k = agg.setdefault(n, [0, 0.0, 0.0])
    k[0] += 1; k[1] += d; k[2] += gap
for n, (c, d, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:30s} x{c:4d}  dur {d:8.1f} us (avg {d / c:6.2f})   gap-before sum {g:8.1f} (avg {g / c:5.2f})")
print("layer 3 kernels (start, dur, gap-before, name):")
per = (len(step) - 4) // layers
for r in rows[2 + 3 * per: 2 + 4 * per]:
    print("   ", r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/decode_trace.json", "w"))
