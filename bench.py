#!/usr/bin/env python
"""Headline benchmark: navigation steps/sec for the Dynam3D per-step hot path
(12 x 224^2 RGB-D views -> CLIP ViT-L/14@336 -> 3D token memory -> llava-phi-3-mini prefill -> next-action logits).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU/PyTorch path (oracle port) on the host cores

One "step" = one navigation step for every episode of the rank's shard (EPISODES_PER_GPU episodes, independent -> weak
scaling); `value` = episodes * steps / time summed over ranks.  Prints ONE JSON line (see the task contract).
"""
import argparse
import json
import math
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1337.2), d.get("hbm_gbs", 6539.5), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


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


def make_inputs(rank, n_steps, episodes):
    eps = [synth.make_episode(1000 * 4 + rank * 100 + b, n_steps=n_steps, num_views=VIEWS, rgb_size=RGB, depth_size=DEPTH, n_seg=N_SEG,
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
# reference arm / cpu_baseline: the reference's own CPU PyTorch path (oracle port), bounded sample, extrapolated
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(threads, budget_s=20.0):
    """Times the reference CPU path (fp32 PyTorch, oracle port) on a bounded sample of ONE episode-step of the bench workload and
    extrapolates to the full step: ViT on 1 of 12 views, 2 of 32 LM layers, 3 of 23 LLaVA-tower layers, the 3D memory on 2 views.
    Returns (steps_per_sec, detail dict)."""
    from oracle import nn_ops as NN
    from oracle.ff_oracle import FeatureFieldsOracle
    from oracle import geometry as G
    torch.set_num_threads(threads)
    ep = synth.make_episode(4242, n_steps=2, num_views=2, rgb_size=RGB, depth_size=DEPTH, n_seg=N_SEG, seg_kind="voronoi")
    t0 = time.perf_counter()
    clip_sd = synth.vit_state_dict(WEIGHT_SEED, layers=24)
    x = NN.clip_preprocess(ep[0]["rgb"][:1], 336)
    t = time.perf_counter()
    with torch.no_grad():
        _, grid = NN.vit_forward(x, clip_sd, 24, 16)
    t_vit_view = time.perf_counter() - t
    with torch.no_grad():
        t = time.perf_counter()
        NN.vit_forward(x, clip_sd, 24, 16, n_layers_run=3, return_hidden=True)
        t_tower = (time.perf_counter() - t) / 3 * 23
    del clip_sd
    pol = synth.policy_state_dict(WEIGHT_SEED)
    ff_sd = {k[len("feature_fields."):]: v for k, v in pol.items() if k.startswith("feature_fields.")}
    ff = FeatureFieldsOracle(ff_sd, batch_size=1)
    rng = np.random.default_rng(0)
    t_ff = 0.0
    with torch.no_grad():
        for s in range(2):
            d576 = G.depth_patch_grid(ep[s]["depth"], 1, 2, q1_fix=True)
            full = G.preprocess_depth(ep[s]["depth"], (0.0, 10.0)).reshape(1, 2, DEPTH, DEPTH)
            g = (rng.standard_normal((1, 2, 576, 768)) * 0.5).astype(np.float16)
            t = time.perf_counter()
            ff.delete_old_features_from_camera_frustum(full, [ep[s]["position"]], [ep[s]["heading"]], num_of_views=2)
            ff.update_feature_fields(d576, g, ep[s]["segm"][None], [ep[s]["position"]], [ep[s]["heading"]], num_of_views=2)
            ff.get_environment_features([ep[s]["position"]], [ep[s]["heading"]])
            if s == 1:
                t_ff = (time.perf_counter() - t) / 2 * VIEWS
    S = 900
    lm_sd = synth.lm_state_dict(WEIGHT_SEED, layers=2)
    emb = synth.hash_uniform((S, 3072), 5, 1.0)
    with torch.no_grad():
        t = time.perf_counter()
        NN.lm_prefill(emb, [S], lm_sd, 2, 32)
        t_lm = (time.perf_counter() - t)
    # the 2-layer call also includes the lm_head (1 row) -- negligible; per-layer time ~ t_lm / 2
    t_lm_full = t_lm / 2 * 32
    step = t_vit_view * VIEWS + t_tower + t_ff + t_lm_full
    detail = {"vit_s_per_view": round(t_vit_view, 3), "llava_tower_s": round(t_tower, 3), "ff_s_per_step": round(t_ff, 3),
              "lm_prefill_s_S900": round(t_lm_full, 3), "sample_wall_s": round(time.perf_counter() - t0, 1)}
    return 1.0 / step, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, detail = [], None
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_sample(cores)
        if i >= args.warmup:
            vals.append(v)
        if i == 0 and detail["sample_wall_s"] * (args.warmup + args.steps) > 240:  # keep the whole arm within a few minutes
            vals = vals or [v]
            break
    value = float(np.mean(vals))
    sample = ("one episode-step extrapolated from: CLIP ViT-L/14@336 on 1 of 12 views, 3 of 23 LLaVA-tower layers, 3D memory on 2 of 12 views, "
              "2 of 32 Phi-3 layers at S=900; fp32 PyTorch, all host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{VIEWS}x{RGB}^2 RGB-D views/episode-step, ViT-L/14@336 + 3D token memory + Phi-3-mini prefill (reference CPU path)",
                       "detail": detail},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this engine
# ------------------------------------------------------------------------------------------------
def build_engine(episodes, n_steps=16):
    from dynam3d_b200.policy import Dynam3D_VLN
    net = Dynam3D_VLN(q1_fix=True, q7_fix=True)  # 12-view panorama: per-view depth / heading (the literal Q1/Q7 paths only make sense at V=1)
    net.load_policy_state_dict(synth.policy_state_dict(WEIGHT_SEED, merge_bias=0.3))
    net.rgb_encoder.max_images = episodes * VIEWS
    net.rgb_encoder.load_openai_state_dict(synth.vit_state_dict(WEIGHT_SEED, layers=24, device="cuda"))
    net.llava.load_state_dict(synth.llava_state_dict(WEIGHT_SEED, clip_layers=24, lm_layers=32, device="cuda", lm_round_to=torch.float16),
                              max_images=episodes, max_tokens=episodes * 1100)
    net.feature_fields.reset(episodes)
    net.feature_fields.reserve(patches=n_steps * VIEWS * 576, instances=4096)  # the rollout horizon is known: no pool growth inside the timed region
    net.tokenize = synth.ToyTokenizer()
    return net


def run_engine(args):
    import torch.distributed as dist
    from dynam3d_b200 import _lib as L
    from dynam3d_b200 import ops
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # one process per GPU shares the host: keep PyTorch's intra-op pool from oversubscribing the cores (the reference pins 4, run.py:90)
    torch.set_num_threads(max(1, min(4, (os.cpu_count() or 4) // max(world, 1))))
    torch.cuda.set_device(local)
    L.require_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E = args.episodes
    n_total = args.warmup + 2 * args.steps
    net = build_engine(E, n_total)
    instr = [synth.make_instruction(rank * 100 + b, INSTR_CHARS) for b in range(E)]
    steps = make_inputs(rank, n_total, E)
    dev = torch.device("cuda", local)
    # device-resident copies for the kernel-side number, pinned host copies for the end-to-end number
    dev_in = [{"rgb": torch.from_numpy(s["rgb"]).to(dev), "depth": torch.from_numpy(s["depth"]).to(dev), "patch_segm": s["segm"]} for s in steps]
    host_in = [{"rgb": torch.from_numpy(s["rgb"]).pin_memory(), "depth": torch.from_numpy(s["depth"]).pin_memory(), "patch_segm": s["segm"]} for s in steps]
    from dynam3d_b200.sharding import allgather_last_logits

    def one_step(i, inputs, gather=True):
        lg = net.forward_logits(inputs[i], instr, steps[i]["pos"], steps[i]["head"], num_of_views=VIEWS)
        if world > 1 and gather:
            lg = allgather_last_logits(lg)
        return lg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i, dev_in)
    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    ops.GEMM_PROFILE = []
    calls0 = L.lib().d3d_launch_count()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only the timed steps are captured
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        one_step(i, dev_in)
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = L.lib().d3d_launch_count() - calls0  # kernels launched by libdynam3d_b200.so (counted at every launch site)
    prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    gemm_flops = sum(p[0] for p in prof)
    gemm_ms = sum(p[1].elapsed_time(p[2]) for p in prof)
    seq_lens = list(net.last_seq_lens)
    # ---- timed region 2: end to end through the public API with HOST buffers (H2D of the step's inputs + D2H of the logits) ----
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out_host = torch.empty((E * world if world > 1 else E, 32064), dtype=torch.float32).pin_memory()
    e2.record()
    for i in range(args.warmup + args.steps, n_total):
        lg = one_step(i, host_in)
        out_host.copy_(lg, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    # ---- per-stage rooflines: ONE extra step (outside both timed regions) with CUDA events around every C-ABI call ----
    stages = []
    if rank == 0:
        ops.STAGE_PROFILE = []
        one_step(args.warmup, dev_in, gather=False)  # rank 0 only: no collective in this extra step
        torch.cuda.synchronize()
        prof_s, ops.STAGE_PROFILE = ops.STAGE_PROFILE, None
        peak_tf_, peak_hbm_, _ = peaks()
        agg = {}
        for name, bound, work, a, b in prof_s:
            d = agg.setdefault(name, {"bound": bound, "work": 0.0, "ms": 0.0, "launches": 0})
            d["work"] += work; d["ms"] += a.elapsed_time(b); d["launches"] += 1
        for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            if d["ms"] <= 0:
                continue
            if d["bound"] == "tensor":
                ach, peak, unit = d["work"] / (d["ms"] * 1e-3) / 1e12, peak_tf_, "TFLOP/s"
            else:
                ach, peak, unit = d["work"] / (d["ms"] * 1e-3) / 1e9, peak_hbm_, "GB/s"
            stages.append({"stage": name, "bound": d["bound"], "launches": d["launches"], "ms": round(d["ms"], 3), "achieved": round(ach, 1),
                           "unit": unit, "frac": round(ach / peak, 3)})
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
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
            "data": "synthetic",
            "config": {"workload": f"full navigation step: {E} episodes/GPU x {VIEWS} views of {RGB}^2 RGB + {DEPTH}^2 depth -> CLIP ViT-L/14@336 (24L) -> "
                                   f"patch/instance/zone 3D token memory ({N_SEG} segments/view) -> LLaVA tower (23L, 1 view) + projector -> "
                                   f"Phi-3-mini (32L) prefill, S~{int(np.mean(seq_lens))} -> next-action logits",
                       "episodes_per_gpu": E, "views": VIEWS, "prefill_tokens": seq_lens, "weights": "random init at true sizes (no checkpoints offline)",
                       "l2": "weights 8.9 GB >> 126 MB L2 are streamed every step (no explicit flush needed)",
                       "precision": "fp16 GEMM operands (reference: fp16 autocast, TR:385), fp32 accumulate / residual / norm statistics"},
            "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(E * 32064 * 4)},
            "gpu_launches": int(launches),
            "stages": stages,  # one extra profiled step: per-stage algorithmic FLOPs or bytes / CUDA-event time vs the measured peaks
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_pair_kernel<256> / gemm_tcgen05_kernel<128|256> (all tcgen05 GEMM launches of the timed region)", "achieved": achieved,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": None, "traffic_note": "per-launch DRAM bytes of the four dominant shapes (ncu --set full) = 1.0-1.4x their algorithmic bytes: profiles/ncu_full_gemm_r01_final_summary.json", "peak_source": which,
                         "gemm_share_of_step": gemm_ms / ms, "gemm_launches": len(prof)},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, detail = cpu_reference_sample(os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "one episode-step extrapolated from: ViT on 1 of 12 views, 3 of 23 tower layers, 3D memory on 2 of 12 views, "
                                              "2 of 32 Phi-3 layers at S=900 (fp32 PyTorch oracle port, all host threads)", "detail": detail}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--episodes", type=int, default=EPISODES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed region (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
