"""Full-depth parity on the bench configuration: ONE episode of the exact bench workload -- 12 views of 224^2 RGB-D,
24-layer CLIP ViT-L/14@336, 23-layer LLaVA tower, 32-layer Phi-3-mini, S ~ 750 -- through the production path and the precise path, against
the CPU oracle run on the host cores (tools/full_depth.py).

  production (fp16 GEMM operands = the reference's fp16 autocast, TR:385) vs the oracle that rounds at the same points:  <= 1.2e-2, same arg-max,
      identical discrete 3D-memory state (16-bit rounding-flip noise floor at 24+32 layers: measured 6.7e-3..7.8e-3 on logits of magnitude 4.7)
  precise (split fp16x2 operands, fp32 activations in the LLaVA tower, the 3D memory, the projections and the LM -- Dynam3D_VLN.PRECISE_DEFAULT;
      the CLIP ViT stays fp16 like the reference's own fp16 feature store, FF:500) vs the pure-fp32 oracle (= the reference on CPU):  <= 1e-3
      -- the north star's tolerance (measured 3.3e-4..3.5e-4)
"""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_bench_config_full_depth_production_and_precise():
    from full_depth import full_depth_parity
    r = full_depth_parity(steps=1, clip_layers=24, lm_layers=32, modes=("production", "precise"))
    assert r["config"]["clip_layers"] == 24 and r["config"]["tower_layers"] == 23 and r["config"]["lm_layers"] == 32 and r["config"]["views"] == 12
    p, q = r["production"], r["precise"]
    print(f"full depth: production vs matched {p['max_abs_vs_matched']:.2e}, vs fp32 {p['max_abs_vs_fp32']:.2e}; precise vs fp32 {q['max_abs_vs_fp32']:.2e}; "
          f"S={r['seq_lens']}, |logit| max {r['logit_absmax']:.2f}, oracle {r['oracle_s_per_step']}")
    assert p["discrete_state_equal"] and p["argmax_equal"] and p["max_abs_vs_matched"] <= 1.2e-2
    assert p["state_equal_fp32"] and p["max_abs_vs_fp32"] <= 1.2e-2
    assert q["discrete_state_equal"] and q["argmax_equal"]
    assert q["max_abs_vs_fp32"] <= 1e-3  # north star: fp action logits within 1e-3 of the reference path
