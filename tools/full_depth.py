"""Full-depth parity of ONE episode of the exact bench workload (12 views of 224^2 RGB-D, 24-layer CLIP ViT-L/14@336, 23-layer LLaVA
tower, 32-layer Phi-3-mini) against the CPU oracle -- test infrastructure (used by tests/test_full_depth_gpu.py and by bench.py's
`parity` post-step, outside every timed region).

Two engine modes and two oracles:
  production (fp16 GEMM operands = the reference's fp16 autocast, TR:385)  vs  oracle with the same rounding points (rnd=round_fp16)
  precise    (split fp16x2 operands, fp32 activations, precise.py)        vs  the oracle's pure fp32 path (rnd=None = the reference on CPU)
and production vs the fp32 oracle as the distance the north star's 1e-3 is about.

    python tools/full_depth.py [--steps 2] [--lm-layers 32] [--clip-layers 24] [--modes production,precise] [--out gpurun_out/full_depth.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VIEWS, RGB, DEPTH, N_SEG, INSTR_CHARS, WEIGHT_SEED = 12, 224, 224, 16, 64, 7  # bench.py's workload constants


def full_depth_parity(steps=1, clip_layers=24, lm_layers=32, modes=("production", "precise"), views=VIEWS, rgb=RGB, episode_seed=4000, log=print):
    """Runs `steps` navigation steps of one bench episode through the oracles (matched rounding + pure fp32) and then through every engine
    mode.  `modes`: "production", "precise" (Dynam3D_VLN.PRECISE_DEFAULT: every stage but the CLIP ViT), "precise:all", or "precise:<parts>" with
    parts a '+'-joined subset of Dynam3D_VLN.PRECISE_PARTS (e.g. "precise:lm+tower").  Returns a dict with, per mode, max_abs_vs_matched (production only) /
    max_abs_vs_fp32 (worst over the steps), argmax_equal, discrete_state_equal, engine ms per step, plus the oracle wall times."""
    from dynam3d_b200 import synth
    from dynam3d_b200.policy import Dynam3D_VLN
    from oracle import nn_ops as NN
    from oracle.policy_oracle import PolicyOracle
    from oracle.ref_compare import snapshots_equal
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    dev = "cuda"
    pol_sd = synth.policy_state_dict(WEIGHT_SEED, merge_bias=0.3)
    clip_dev = synth.vit_state_dict(WEIGHT_SEED, layers=clip_layers, device=dev)
    llava_dev = synth.llava_state_dict(WEIGHT_SEED, clip_layers=clip_layers, lm_layers=lm_layers, device=dev, lm_round_to=torch.float16)
    net = Dynam3D_VLN(q1_fix=True, q7_fix=True)
    net.load_policy_state_dict(pol_sd)
    net.rgb_encoder.max_images = views
    net.rgb_encoder.load_openai_state_dict(clip_dev)
    net.llava.load_state_dict(llava_dev, max_images=1, max_tokens=1100)
    net.tokenize = synth.ToyTokenizer()
    net.llava.lm  # build the engine-layout weights now
    net.rgb_encoder.engine
    clip_cpu = {k: v.cpu() for k, v in clip_dev.items()}
    llava_cpu = {k: v.cpu() for k, v in llava_dev.items()}
    del clip_dev, llava_dev
    log(f"[full_depth] weights ready ({time.perf_counter() - t0:.1f} s): clip {clip_layers}L, tower {clip_layers - 1}L of {clip_layers}, LM {lm_layers}L")
    orcs = {"fp32": PolicyOracle(pol_sd, clip_cpu, llava_cpu, clip_layers=clip_layers, lm_layers=lm_layers, batch_size=1, rnd=None, q1_fix=True, q7_fix=True)}
    if any(m == "production" for m in modes):
        orcs["matched"] = PolicyOracle(pol_sd, clip_cpu, llava_cpu, clip_layers=clip_layers, lm_layers=lm_layers, batch_size=1, rnd=NN.round_fp16,
                                       q1_fix=True, q7_fix=True)
    ep = synth.make_episode(episode_seed, n_steps=steps, num_views=views, rgb_size=rgb, depth_size=DEPTH, n_seg=N_SEG, seg_kind="voronoi")
    instr = [synth.make_instruction(0, INSTR_CHARS)]
    out = {"config": {"episodes": 1, "views": views, "rgb": rgb, "clip_layers": clip_layers, "tower_layers": clip_layers - 1, "lm_layers": lm_layers,
                      "steps": steps, "segments_per_view": N_SEG}, "oracle_s_per_step": {}, "seq_lens": [], "logit_absmax": 0.0}
    tok, build = net.tokenize, net.build_prompt
    want = {name: [] for name in orcs}   # per step: (logits, snapshot, lens)
    for t in range(steps):
        obs = {"rgb": ep[t]["rgb"], "depth": ep[t]["depth"], "patch_segm": ep[t]["segm"][None]}
        pos, head = [ep[t]["position"]], [ep[t]["heading"]]
        for name, orc in orcs.items():
            tt = time.perf_counter()
            with torch.no_grad():
                lg = orc.step_logits(obs, pos, head, lambda b, n_img: tok(build(n_img, instr[b], ["none\n"] * 4)), num_of_views=views)
            want[name].append((lg, orc.ff.snapshot(0), list(orc.last_lens)))
            out["oracle_s_per_step"].setdefault(name, []).append(round(time.perf_counter() - tt, 2))
            log(f"[full_depth] step {t} oracle[{name}] S={orc.last_lens} {out['oracle_s_per_step'][name][-1]} s")
        out["seq_lens"].append(want["fp32"][-1][2][0])
        out["logit_absmax"] = max(out["logit_absmax"], float(want["fp32"][-1][0].abs().max()))
    for mode in modes:
        parts = () if mode == "production" else (Dynam3D_VLN.PRECISE_DEFAULT if mode == "precise" else
                                                 (Dynam3D_VLN.PRECISE_PARTS if mode == "precise:all" else tuple(mode.split(":", 1)[1].split("+"))))
        net.set_precise_parts(parts)
        net.feature_fields.reset(1)
        ref_name = "matched" if mode == "production" else "fp32"
        r = {"max_abs_vs_matched": 0.0 if mode == "production" else None, "max_abs_vs_fp32": 0.0, "argmax_equal": True, "discrete_state_equal": True,
             "state_equal_fp32": True, "engine_ms": []}
        for t in range(steps):
            t_obs = {"rgb": torch.from_numpy(ep[t]["rgb"]), "depth": torch.from_numpy(ep[t]["depth"]), "patch_segm": ep[t]["segm"][None]}
            pos, head = [ep[t]["position"]], [ep[t]["heading"]]
            torch.cuda.synchronize()
            tt = time.perf_counter()
            got = net.forward_logits(t_obs, instr, pos, head, num_of_views=views).float().cpu()
            r["engine_ms"].append(round(1e3 * (time.perf_counter() - tt), 1))
            snap = net.feature_fields.snapshot(0)
            lg_ref, snap_ref, lens_ref = want[ref_name][t]
            lg32, snap32, lens32 = want["fp32"][t]
            same = snapshots_equal(snap_ref, snap) == [] and net.last_seq_lens == lens_ref
            same32 = snapshots_equal(snap32, snap) == [] and net.last_seq_lens == lens32
            r["discrete_state_equal"] &= bool(same)
            r["state_equal_fp32"] &= bool(same32)
            if mode == "production" and same:
                r["max_abs_vs_matched"] = max(r["max_abs_vs_matched"], float((got - lg_ref).abs().max()))
            if same32:
                r["max_abs_vs_fp32"] = max(r["max_abs_vs_fp32"], float((got - lg32).abs().max()))
            if same:
                r["argmax_equal"] &= bool(torch.equal(got.argmax(-1), lg_ref.argmax(-1)))
            log(f"[full_depth] step {t} {mode}: S={net.last_seq_lens} {r['engine_ms'][-1]} ms  vs matched {r['max_abs_vs_matched']}, vs fp32 "
                f"{r['max_abs_vs_fp32']:.3e}, argmax_equal {r['argmax_equal']}, state equal {r['discrete_state_equal']} (fp32 oracle: {r['state_equal_fp32']})")
        out[mode] = r
    out["wall_s"] = round(time.perf_counter() - t0, 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--clip-layers", type=int, default=24)
    ap.add_argument("--lm-layers", type=int, default=32)
    ap.add_argument("--modes", default="production,precise", help='comma list of production | precise | precise:<part+part> (parts: vit,tower,ff,proj,lm)')
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    r = full_depth_parity(a.steps, a.clip_layers, a.lm_layers, tuple(a.modes.split(",")))
    print(json.dumps(r))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(r, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
