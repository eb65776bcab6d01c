#!/usr/bin/env python
"""Headline benchmark: navigation steps/sec for the Dynam3D per-step hot path
(12 x 224^2 RGB-D views -> CLIP ViT-L/14@336 -> 3D token memory -> llava-phi-3-mini prefill -> next-action logits).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU/PyTorch path (oracle port) on the host cores
    python bench.py --precise                                # the <= 1e-3 logit-parity mode (split fp16x2 operands, fp32 activations)
    python bench.py --generate                               # the whole POL:463 step: prefill + 20-token greedy decode
    python bench.py --workload long_horizon                  # BASELINE configs[4]: 256 one-view steps, memory -> 4 k instance slots

One "step" = one navigation step for every episode of the rank's shard (EPISODES_PER_GPU episodes, independent -> weak
scaling); `value` = episodes * steps / time summed over ranks.  Prints ONE JSON line (see the task contract).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from dynam3d_b200 import synth  # noqa: E402

EPISODES_PER_GPU = 8   # BASELINE.json configs[3]: 64 episodes over 8 GPUs
VIEWS = 12
RGB = 224
DEPTH = 224
N_SEG = 16
INSTR_CHARS = 64
WEIGHT_SEED = 7
METRIC = "navigation steps/sec (12x224^2 RGB-D->3D tokens->Phi-3) @1/2/4/8 B200"
GEMM_TRAFFIC = os.path.join(ROOT, "profiles", "gemm_traffic_r02.json")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1337.2), d.get("hbm_gbs", 6539.5), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def decode_roofline(net, seq_lens, peak_hbm):
    """The decode token step of the --generate line on its own (POL:463): (prefill + 20 tokens) - (prefill) over 19 steps at the
    step's own prompt lengths, CUDA events; algorithmic bytes per token step = every language-model weight once + the K / V rows of the prompts."""
    lm = net.llava.lm
    w = lm.w
    dev = net.device
    lens = [int(n) for n in seq_lens]
    x = torch.randn((sum(lens), w.hidden), device=dev, dtype=torch.float32) * 0.05
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
    pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(dev)
    last = (cu[1:] - 1).to(torch.int32).contiguous()
    res = {}
    for n_new in (1, 20):
        ts = []
        for _ in range(3):
            xi = x.clone()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if n_new == 1:
                lm.prefill(xi, cu, pos, len(lens), max(lens), last)
            else:
                lm.generate(xi, cu, pos, len(lens), max(lens), last, max_new_tokens=n_new, eos_ids=())
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[n_new] = sorted(ts)[1]
    step_ms = (res[20] - res[1]) / 19
    n_layers = len(w.layers)
    wbytes = n_layers * (4 * w.hidden * w.hidden + 3 * w.ffn * w.hidden) * 2 + w.vocab * w.hidden * 2
    kvbytes = n_layers * sum(lens) * 2 * w.hidden * 2
    gbs = (wbytes + kvbytes) / step_ms / 1e6
    return {"ms_per_token_step": step_ms, "sequences": len(lens), "weight_bytes": wbytes, "kv_bytes": kvbytes, "achieved": gbs, "peak": peak_hbm,
            "unit": "GB/s", "frac": gbs / peak_hbm, "bound": "hbm",
            "note": "round 1: 3.86 ms (39 %); the chain of ~200 dependent kernels per token is latency-, not bandwidth-limited (DESIGN.md, Decode)"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": sorted(reasons)}


def make_inputs(rank, n_steps, episodes, views=VIEWS, rgb=RGB, depth=DEPTH, n_seg=N_SEG, seed0=4000):
    eps = [synth.make_episode(seed0 + rank * 100 + b, n_steps=n_steps, num_views=views, rgb_size=rgb, depth_size=depth, n_seg=n_seg,
                              seg_kind="voronoi") for b in range(episodes)]
    steps = []
    for t in range(n_steps):
        steps.append({
            "rgb": np.concatenate([eps[b][t]["rgb"] for b in range(episodes)], 0),
            "depth": np.concatenate([eps[b][t]["depth"] for b in range(episodes)], 0),
            "segm": np.stack([eps[b][t]["segm"] for b in range(episodes)], 0),
            "pos": [eps[b][t]["position"] for b in range(episodes)], "head": [eps[b][t]["heading"] for b in range(episodes)]})
    return steps


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU PyTorch path (oracle port, fp32) timed on WHOLE episode-steps of the bench workload
# ------------------------------------------------------------------------------------------------
def reference_oracle(threads):
    """PolicyOracle (pure fp32 = the reference's arithmetic on CPU) at full depth with the bench's weights.  The weights are synthesised with
    torch (on the GPU when one is there -- set-up only, bit-identical to the CPU generator) and live on the host."""
    from oracle.policy_oracle import PolicyOracle
    torch.set_num_threads(threads)
    gen = "cuda" if torch.cuda.is_available() else "cpu"
    pol_sd = synth.policy_state_dict(WEIGHT_SEED, merge_bias=0.3)
    clip_sd = {k: v.cpu() for k, v in synth.vit_state_dict(WEIGHT_SEED, layers=24, device=gen).items()}
    llava_sd = {k: v.cpu() for k, v in synth.llava_state_dict(WEIGHT_SEED, clip_layers=24, lm_layers=32, device=gen, lm_round_to=torch.float16).items()}
    if gen == "cuda":
        torch.cuda.empty_cache()
    return PolicyOracle(pol_sd, clip_sd, llava_sd, clip_layers=24, lm_layers=32, batch_size=1, rnd=None, q1_fix=True, q7_fix=True)


def run_reference(args):
    """Times real, whole episode-steps (a1-a16: CLIP ViT-L/14@336 on 12 views, the full 3D-memory update, LLaVA tower, 32-layer Phi-3 prefill
    at the sequence length the memory produces) of ONE episode of the bench workload, consecutive steps of the same rollout.  A step takes
    ~8-11 s on 16 cores, so the number of timed steps is bounded by a wall budget; `steps` in the line is what was actually timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t_setup = time.perf_counter()
    orc = reference_oracle(cores)
    tok = synth.ToyTokenizer()
    from dynam3d_b200.prompt import build_prompt as build  # pure string formatting (POL:436); no engine code runs in this arm
    n_total = max(1, min(args.warmup, 1)) + args.steps
    ep = synth.make_episode(4000, n_steps=n_total, num_views=VIEWS, rgb_size=RGB, depth_size=DEPTH, n_seg=N_SEG, seg_kind="voronoi")
    instr = [synth.make_instruction(0, INSTR_CHARS)]
    setup_s = time.perf_counter() - t_setup
    budget_s = args.reference_budget
    times, lens = [], []
    t_begin = time.perf_counter()
    for t in range(n_total):
        obs = {"rgb": ep[t]["rgb"], "depth": ep[t]["depth"], "patch_segm": ep[t]["segm"][None]}
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.step_logits(obs, [ep[t]["position"]], [ep[t]["heading"]],
                            lambda b, n_img: tok(build(n_img, instr[b], ["none\n"] * 4)), num_of_views=VIEWS)
        dt = time.perf_counter() - t0
        if t >= n_total - args.steps:
            times.append(dt)
            lens.append(orc.last_lens[0])
        if time.perf_counter() - t_begin + dt > budget_s and len(times) >= 1:
            break
    value = len(times) / sum(times)
    sample = (f"{len(times)} whole episode-steps of one bench episode (12 x {RGB}^2 views -> ViT-L/14@336 24L -> 3D memory update -> LLaVA tower 23L -> "
              f"Phi-3 32L prefill at S={lens}), consecutive steps after {n_total - args.steps} warm-up step(s); fp32 PyTorch oracle port, {cores} host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": len(times), "warmup": n_total - args.steps,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"full navigation step of ONE episode: {VIEWS} views of {RGB}^2 RGB + {DEPTH}^2 depth -> CLIP ViT-L/14@336 (24L) -> "
                                   f"patch/instance/zone 3D token memory ({N_SEG} segments/view) -> LLaVA tower (23L, 1 view) + projector -> Phi-3-mini (32L) "
                                   f"prefill -> next-action logits (reference CPU path; the engine arm steps {EPISODES_PER_GPU} such episodes per GPU)",
                       "prefill_tokens": lens, "step_seconds": [round(x, 2) for x in times], "setup_s": round(setup_s, 1),
                       "requested_steps": args.steps, "wall_budget_s": budget_s},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this engine
# ------------------------------------------------------------------------------------------------
def build_engine(episodes, n_steps=16, precise_parts=(), views=VIEWS):
    from dynam3d_b200.policy import Dynam3D_VLN
    net = Dynam3D_VLN(q1_fix=True, q7_fix=True)  # 12-view panorama: per-view depth / heading (the literal Q1/Q7 paths only make sense at V=1)
    net.set_precise_parts(precise_parts)
    net.load_policy_state_dict(synth.policy_state_dict(WEIGHT_SEED, merge_bias=0.3))
    net.rgb_encoder.max_images = episodes * views
    # the checkpoint tensors stay resident as module state (state_dict round trip, TR:75-84): hand them over in 16 bit
    h = lambda sd: {k: (v.half() if v.dim() >= 2 else v) for k, v in sd.items()}
    net.rgb_encoder.load_openai_state_dict(h(synth.vit_state_dict(WEIGHT_SEED, layers=24, device="cuda")))
    net.llava.load_state_dict(h(synth.llava_state_dict(WEIGHT_SEED, clip_layers=24, lm_layers=32, device="cuda", lm_round_to=torch.float16)),
                              max_images=episodes, max_tokens=episodes * 1100)
    net.llava.lm  # build the engine-layout weights now
    net.rgb_encoder.engine
    torch.cuda.empty_cache()
    net.feature_fields.reset(episodes)
    net.feature_fields.reserve(patches=n_steps * views * 576, instances=4096)  # the rollout horizon is known: commit the pools up front
    net.tokenize = synth.ToyTokenizer()
    return net


def stage_table(prof_s, ff_prof):
    peak_tf_, peak_hbm_, _ = peaks()
    agg = {}
    for name, bound, work, a, b in prof_s:
        d = agg.setdefault(name, {"bound": bound, "work": 0.0, "ms": 0.0, "launches": 0})
        d["work"] += work; d["ms"] += a.elapsed_time(b); d["launches"] += 1
    if ff_prof is not None:
        ms, work, n = ff_prof
        for i, (name, bound) in enumerate((("ff.knn", "hbm"), ("ff.disc", "tensor"), ("ff.new_slots", "hbm"), ("ff.merge_pool", "tensor"),
                                           ("ff.zone_pool", "tensor"), ("ff.result_copy", "hbm"))):
            if n[i]:
                agg[name] = {"bound": bound, "work": float(work[i]), "ms": float(ms[i]), "launches": int(n[i])}
    stages = []
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        if d["ms"] <= 0:
            continue
        if d["bound"] == "tensor":
            ach, peak, unit = d["work"] / (d["ms"] * 1e-3) / 1e12, peak_tf_, "TFLOP/s"
        else:
            ach, peak, unit = d["work"] / (d["ms"] * 1e-3) / 1e9, peak_hbm_, "GB/s"
        stages.append({"stage": name, "bound": d["bound"], "launches": d["launches"], "ms": round(d["ms"], 3), "achieved": round(ach, 1),
                       "unit": unit, "frac": round(ach / peak, 4)})
    return stages


def profiled_step(fn):
    """Runs fn() once with CUDA events around every C-ABI call (ops.STAGE_PROFILE) and inside the view runtime (d3d_ff_profile_*)."""
    import ctypes
    from dynam3d_b200 import _lib as L
    from dynam3d_b200 import ops
    ops.STAGE_PROFILE = []
    L.lib().d3d_ff_profile_begin()
    fn()
    ms, work, n = (ctypes.c_float * 6)(), (ctypes.c_double * 6)(), (ctypes.c_int * 6)()
    L.check(L.lib().d3d_ff_profile_end(ctypes.cast(ms, ctypes.c_void_p), ctypes.cast(work, ctypes.c_void_p), ctypes.cast(n, ctypes.c_void_p)))
    torch.cuda.synchronize()
    prof_s, ops.STAGE_PROFILE = ops.STAGE_PROFILE, None
    return stage_table(prof_s, (list(ms), list(work), list(n)))


def gemm_traffic():
    """Per-launch DRAM bytes of the dominant kernel family from the committed ncu --set full capture (launch-weighted over the shapes of a step)."""
    if not os.path.isfile(GEMM_TRAFFIC):
        return None, "no ncu traffic file"
    d = json.load(open(GEMM_TRAFFIC))
    n = sum(s["launches_per_step"] for s in d["shapes"])
    tr = sum(s["launches_per_step"] * s["dram_MB"] for s in d["shapes"]) / n * 1e6
    al = sum(s["launches_per_step"] * s["algorithmic_MB"] for s in d["shapes"]) / n * 1e6
    return tr, (f"launch-weighted mean over the {len(d['shapes'])} dominant GEMM shapes of a step ({d['source']}): {tr / 1e6:.0f} MB DRAM traffic vs "
                f"{al / 1e6:.0f} MB algorithmic per launch = {tr / al:.2f}x")


def run_engine(args):
    import torch.distributed as dist
    from dynam3d_b200 import _lib as L
    from dynam3d_b200 import ops
    from dynam3d_b200.policy import Dynam3D_VLN
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # one process per GPU shares the host: keep PyTorch's intra-op pool from oversubscribing the cores (the reference pins 4, run.py:90),
    # and give every rank its own slice of the cores so the view-loop planners do not migrate onto each other
    ncpu = os.cpu_count() or 4
    torch.set_num_threads(max(1, min(4, ncpu // max(world, 1))))
    if world > 1 and hasattr(os, "sched_setaffinity"):
        per = max(1, ncpu // world)
        try:
            os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))))
        except OSError:
            pass
    torch.cuda.set_device(local)
    L.require_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E = args.episodes
    n_total = max(args.warmup + args.steps + 1, 6)  # region 2 replays the steps of region 1; the precise re-timing uses steps 0..4
    parts = Dynam3D_VLN.PRECISE_DEFAULT if args.precise else tuple(p for p in (args.precise_parts or "").split(",") if p)
    mode = "production" if not parts else ("precise" if set(parts) == set(Dynam3D_VLN.PRECISE_DEFAULT) else
                                           ("precise:all" if set(parts) == set(Dynam3D_VLN.PRECISE_PARTS) else "precise:" + "+".join(parts)))
    net = build_engine(E, n_total, parts)
    if args.chunked_prefill:
        net.chunked_prefill, net.overlap_sms = True, tuple(int(x) for x in args.chunked_prefill.split(","))
    instr = [synth.make_instruction(rank * 100 + b, INSTR_CHARS) for b in range(E)]
    steps = make_inputs(rank, n_total, E)
    dev = torch.device("cuda", local)
    # device-resident copies for the kernel-side number, pinned host copies for the end-to-end number
    dev_in = [{"rgb": torch.from_numpy(s["rgb"]).to(dev), "depth": torch.from_numpy(s["depth"]).to(dev), "patch_segm": s["segm"]} for s in steps]
    host_in = [{"rgb": torch.from_numpy(s["rgb"]).pin_memory(), "depth": torch.from_numpy(s["depth"]).pin_memory(), "patch_segm": s["segm"]} for s in steps]
    from dynam3d_b200.sharding import LogitsGather
    gather_q = LogitsGather(depth=2)   # in-flight all-gathers: episodes are independent (BASE:770), so the gather of step i overlaps step i+1

    def one_step(i, inputs, gather=True):
        if args.generate:
            lg, ids = net.generate_ids(inputs[i], instr, steps[i]["pos"], steps[i]["head"], num_of_views=VIEWS)
        else:
            lg = net.forward_logits(inputs[i], instr, steps[i]["pos"], steps[i]["head"], num_of_views=VIEWS)
        if world > 1 and gather:
            gather_q.post(lg)  # at most two gathers in flight: bounds memory, never stalls the step that just finished
        return lg

    def drain():
        gather_q.drain()

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i, dev_in)
    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    L.lib().d3d_gemm_profile_begin()  # CUDA events around every tcgen05 GEMM launch of the timed region, recorded inside the library
    calls0 = L.lib().d3d_launch_count()
    if args.profile_range:
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only the timed steps are captured
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        one_step(i, dev_in)
    drain()
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = L.lib().d3d_launch_count() - calls0  # kernels launched by libdynam3d_b200.so (counted at every launch site)
    import ctypes
    g_fl, g_ms, g_n = (ctypes.c_double * 3)(), (ctypes.c_float * 3)(), (ctypes.c_int * 3)()
    L.check(L.lib().d3d_gemm_profile_end(ctypes.cast(g_fl, ctypes.c_void_p), ctypes.cast(g_ms, ctypes.c_void_p), ctypes.cast(g_n, ctypes.c_void_p)))
    gemm_flops, gemm_ms, gemm_n = g_fl[0], g_ms[0], g_n[0]                    # the dominant kernel: gemm_tcgen05_pair_kernel<256>
    all_flops, all_ms, all_n = sum(g_fl), sum(g_ms), sum(g_n)                 # every tcgen05 GEMM launch (incl. the small pooled-encoder GEMMs)
    seq_lens = list(net.last_seq_lens)
    # ---- timed region 2: end to end through the public API with HOST buffers (H2D of the step's inputs + D2H of the rank's logits) ----
    # The SAME episode steps as region 1: the rollout is restarted (memory reset, the warm-up steps replayed untimed, also from host buffers), so
    # the two numbers differ by the copies only -- timed on the steps that follow, e2e also paid for a longer 3D memory and longer prompts.
    net.feature_fields.reset(E)
    net.feature_fields.reserve(patches=n_total * VIEWS * 576, instances=4096)
    torch.cuda.empty_cache()  # the caching allocator as cold as it was for region 1 (the prompt lengths, hence the block sizes, grow step by step)
    for i in range(args.warmup):
        one_step(i, host_in)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out_host = torch.empty((E, 32064), dtype=torch.float32).pin_memory()
    e2.record()
    for i in range(args.warmup, args.warmup + args.steps):
        lg = one_step(i, host_in)
        out_host.copy_(lg, non_blocking=True)  # the rank's own rows: E x 32064 fp32
    drain()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    finite = bool(torch.isfinite(out_host).all())
    # ---- per-stage rooflines: ONE extra step (outside both timed regions) with CUDA events around every C-ABI call ----
    stages = []
    if rank == 0:
        stages = profiled_step(lambda: one_step(args.warmup, dev_in, gather=False))  # rank 0 only: no collective in this extra step
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        peak_tf, peak_hbm, which = peaks()
        value = world * E * args.steps / (ms / 1000.0)
        e2e = world * E * args.steps / (ms_e2e / 1000.0)
        achieved = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
        h2d = int(steps[0]["rgb"].nbytes + steps[0]["depth"].nbytes)
        traffic, traffic_note = gemm_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
            "data": "synthetic",
            "config": {"workload": f"full navigation step: {E} episodes/GPU x {VIEWS} views of {RGB}^2 RGB + {DEPTH}^2 depth -> CLIP ViT-L/14@336 (24L) -> "
                                   f"patch/instance/zone 3D token memory ({N_SEG} segments/view) -> LLaVA tower (23L, 1 view) + projector -> "
                                   f"Phi-3-mini (32L) prefill, S~{int(np.mean(seq_lens))} -> next-action logits"
                                   + (" -> greedy decode (<= 20 tokens, KV cache; POL:463)" if args.generate else ""),
                       "mode": mode, "episodes_per_gpu": E, "views": VIEWS, "prefill_tokens": seq_lens, "weights": "random init at true sizes (no checkpoints offline)",
                       "l2": "weights 8.9 GB >> 126 MB L2 are streamed every step (no explicit flush needed)",
                       "precision": ("fp16 GEMM operands (reference: fp16 autocast, TR:385), fp32 accumulate / residual / norm statistics" if not parts else
                                     "split fp16x2 tensor-core operands (A_hi W + A_lo W), fp32 activations / attention in: " + "+".join(parts))},
            "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(E * 32064 * 4), "logits_finite": finite,
                    "episode_steps": [args.warmup, args.warmup + args.steps - 1],
                    "note": "a second, separately timed pass over the same episode steps as `value` (memory reset, allocator cache emptied, warm-up "
                            "replayed untimed) with pinned HOST inputs: every step copies its 12 RGB-D views per episode to the device (on a copy "
                            "stream, under the previous step's queued kernels) and reads the rank's logits back; it lands within run-to-run noise of "
                            "`value` because those copies are hidden"},
            "gpu_launches": int(launches),
            "stages": stages,  # one extra profiled step: per-stage algorithmic FLOPs or bytes / CUDA-event time vs the measured peaks
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_pair_kernel<256> (every launch of the timed region: ViT / tower / Phi-3 layers and the step-level pooled encoder)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": which, "share_of_step": gemm_ms / ms, "launches": gemm_n,
                         "all_tcgen05_gemms": {"achieved": all_flops / (all_ms / 1000.0) / 1e12 if all_ms > 0 else 0.0, "share_of_step": all_ms / ms, "launches": all_n,
                                               "note": "incl. gemm_tcgen05_kernel<128|256> on the small shapes (pooled encoders of merged instances / zones, discriminator, projections)"}},
        }
        if args.generate:
            line["decode"] = decode_roofline(net, seq_lens, peak_hbm)
        if world == 1 and not args.no_parity:
            if mode == "production" and not args.generate:
                # the <= 1e-3 mode on the same engine, same workload, device-timed (split fp16x2 operands / fp32 activations in every stage whose
                # rounding reaches the logits): a fresh rollout of the same episodes, 2 warm-up + 3 timed steps
                net.set_precise_parts(Dynam3D_VLN.PRECISE_DEFAULT)
                net.feature_fields.reset(E)
                net.feature_fields.reserve(patches=n_total * VIEWS * 576, instances=4096)
                for i in range(2):
                    one_step(i, dev_in, gather=False)
                torch.cuda.synchronize()
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for i in range(2, 5):
                    one_step(i, dev_in, gather=False)
                p1.record()
                torch.cuda.synchronize()
                ms_p = p0.elapsed_time(p1) / 3
                line["precise"] = {"parts": list(Dynam3D_VLN.PRECISE_DEFAULT), "value": E / (ms_p / 1e3), "unit": "steps/s", "ms_per_step": ms_p, "steps": 3,
                                   "warmup": 2, "cost_vs_production": ms_p / (ms / args.steps),
                                   "note": "same engine and workload in the <= 1e-3 logit-parity mode (see parity.precise); bench.py --precise times it alone"}
            # ONE episode of this exact workload at full depth vs the CPU oracle (checker only, outside the timed regions): production is compared
            # with the oracle that rounds at the same points AND with the pure-fp32 oracle; the precise mode with the pure-fp32 oracle
            del net, dev_in, host_in
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from full_depth import full_depth_parity
            modes = ("production", "precise") if mode == "production" else (mode,)
            r = full_depth_parity(steps=1, modes=modes, log=lambda *a: None)
            m = r[mode]
            line["parity"] = {"mode": mode, "max_abs_vs_matched": m["max_abs_vs_matched"], "max_abs_vs_fp32": m["max_abs_vs_fp32"],
                              "argmax_equal": m["argmax_equal"], "discrete_state_equal": m["discrete_state_equal"], "config": r["config"],
                              "seq_lens": r["seq_lens"], "logit_absmax": r["logit_absmax"], "tolerance_north_star": 1e-3}
            if mode == "production":
                q = r["precise"]
                line["parity"]["precise"] = {"parts": list(Dynam3D_VLN.PRECISE_DEFAULT), "max_abs_vs_fp32": q["max_abs_vs_fp32"],
                                             "argmax_equal": q["argmax_equal"], "discrete_state_equal": q["discrete_state_equal"]}
            if not args.no_cpu_baseline:
                sec = r["oracle_s_per_step"]["fp32"][0]
                line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": f"1 whole episode-step (step 0 of a bench episode: 12 views, 24L ViT, 23L tower, 32L Phi-3 at S={r['seq_lens'][0]}) "
                                                  f"through the fp32 PyTorch oracle port on all host threads: {sec} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]: long horizon -- 256 one-view steps, the token memory growing to ~4 k instance slots / 147 k patches
# ------------------------------------------------------------------------------------------------
def run_long_horizon(args):
    """The memory path only (frustum cull over all stored patches, K-NN over all instance slots, pooling / merge re-encode, zones, export) on
    hash-generated CLIP grid features: per-step device time as the memory grows, and the growing stages as roofline fractions."""
    import ctypes
    from dynam3d_b200 import _lib as L
    from dynam3d_b200 import ops
    from dynam3d_b200.feature_fields import Feature_Fields
    torch.cuda.set_device(0)
    L.require_device(0)
    E, T = args.episodes, args.horizon
    pol = synth.policy_state_dict(WEIGHT_SEED, merge_bias=args.merge_bias)
    ff = Feature_Fields(batch_size=E, device="cuda", q7_fix=True)
    ff.load_state_dict({k[len("feature_fields."):]: v for k, v in pol.items() if k.startswith("feature_fields.")})
    ff.reset(E)  # no reserve(): the pools grow inside the rollout (VMM chunks, no copy / no stall)
    steps = make_inputs(0, T, E, views=1, rgb=8, depth=256, n_seg=17, seed0=5000)
    grids = [synth.hash_uniform((E, 1, 576, 768), 5000 * 1000 + t % 8, 0.9, device="cuda").half() for t in range(8)]
    depth_dev = [torch.from_numpy(s["depth"]).cuda().reshape(E, 256, 256).contiguous() for s in steps]
    checkpoints = sorted({1, 2, 4, 8, 16, 32, 64, 128, 192, T})
    curve, ev = [], []
    sampler = ClockSampler(0)
    sampler.start()
    torch.cuda.synchronize()
    calls0 = L.lib().d3d_launch_count()
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    stage_rows = {}
    for t in range(T):
        s = steps[t]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prof = (t + 1) in (128, T)
        if prof:
            torch.cuda.synchronize()
            ops.STAGE_PROFILE = []
            L.lib().d3d_ff_profile_begin()
        a.record()
        d576 = ops.depth_patch_grid(depth_dev[t], E, 1, 24, 24, literal_q1=False)
        full = ops.depth_preprocess(depth_dev[t], 0.0, 10.0).view(E, 1, 256, 256)
        ff.delete_old_features_from_camera_frustum(full, s["pos"], s["head"], num_of_views=1)
        ff.update_feature_fields(d576.view(E, 1, 576), grids[t % 8], batch_position=s["pos"], batch_heading=s["head"], num_of_views=1,
                                 batch_patch_segm=s["segm"])
        env = ff.get_environment_features(s["pos"], s["head"])
        b.record()
        ev.append((a, b))
        if prof:
            ms_, work_, n_ = (ctypes.c_float * 6)(), (ctypes.c_double * 6)(), (ctypes.c_int * 6)()
            L.check(L.lib().d3d_ff_profile_end(ctypes.cast(ms_, ctypes.c_void_p), ctypes.cast(work_, ctypes.c_void_p), ctypes.cast(n_, ctypes.c_void_p)))
            ps, ops.STAGE_PROFILE = ops.STAGE_PROFILE, None
            stage_rows[t + 1] = stage_table(ps, (list(ms_), list(work_), list(n_)))
        if (t + 1) in checkpoints:
            curve.append({"step": t + 1, "n_patches": int(ff.eps[0].n_patch), "n_inst_slots": int(ff.eps[0].n_inst), "n_zone_slots": int(ff.eps[0].n_zone),
                          "env_instance_tokens": int(env["batch_instance_fts"][0].shape[0]), "i": t})
    t_all1.record()
    torch.cuda.synchronize()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = L.lib().d3d_launch_count() - calls0
    per_step = [a.elapsed_time(b) for a, b in ev]
    for c in curve:
        i = c.pop("i")
        lo = max(0, i - 3)
        c["ms_per_step"] = round(float(np.mean(per_step[lo:i + 1])), 3)
    total_ms = t_all0.elapsed_time(t_all1)
    peak_tf, peak_hbm, which = peaks()
    cull = next((r for r in stage_rows.get(T, []) if r["stage"] == "ff.frustum_cull"), None)
    line = {"metric": "memory-update steps/sec, long horizon (256 one-view steps, token memory -> 4k instance slots) @1 B200", "value": E * T / (total_ms / 1e3),
            "unit": "steps/s", "n_gpus": 1, "steps": T, "warmup": 0, "ms_per_step": total_ms / T, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: {E} episodes x {T} one-view steps, 17 segments/view, merge bias {args.merge_bias}: frustum cull over all stored "
                                   "patches + K-NN over all instance slots + patch->instance->zone pooling + export (memory path only, CLIP grid features synthetic)",
                       "episodes_per_gpu": E, "pools": "VMM-backed, no reserve(): grown inside the rollout"},
            "curve": curve, "stages_at_step": stage_rows, "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": ({"bound": "hbm", "kernel": "frustum_cull_kernel (all stored patches of an episode, step %d)" % T, "achieved": cull["achieved"], "peak": peak_hbm,
                          "unit": "GB/s", "frac": cull["frac"], "traffic": None, "peak_source": which} if cull else None)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--episodes", type=int, default=EPISODES_PER_GPU)
    ap.add_argument("--workload", default="step", choices=["step", "long_horizon"])
    ap.add_argument("--precise", action="store_true", help="the <= 1e-3 mode: split-operand fp32-activation arithmetic in tower, 3D memory, projections and LM")
    ap.add_argument("--precise-parts", default=None, help="comma list out of vit,tower,ff,proj,lm (error-vs-cost curve)")
    ap.add_argument("--generate", action="store_true", help="time the whole POL:463 step: prefill + greedy decode of 20 tokens")
    ap.add_argument("--chunked-prefill", default=None, help="SIDE,MAIN SM caps: prefill the prompt prefix on the side stream during the memory update (experiment)")
    ap.add_argument("--horizon", type=int, default=256)
    ap.add_argument("--merge-bias", type=float, default=-10.0, help="long_horizon: discriminator bias (-10 = merges off -> 4k instance slots)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-depth parity check against the CPU oracle after the timed regions")
    ap.add_argument("--reference-budget", type=float, default=170.0, help="--impl reference: wall budget in seconds for the timed whole steps")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed region (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "long_horizon":
        run_long_horizon(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
