"""profiles/ncu_full_gemm_r02_summary.json (tools/ncu_summary.py of the `GEMM_ONE=step python tools/gemm_one.py` capture) ->
profiles/gemm_traffic_r02.json: per dominant GEMM shape of a bench step the measured DRAM bytes per launch, the algorithmic bytes
(A + W + C [+ residual read]) and how often the shape launches per step.  bench.py reads it for `roofline.traffic`."""
import json
import sys

SHAPES = [  # label, M, N, K, out bytes/elt, residual, launches per step
    ("ViT qkv 55392x3072x1024 +bias fp16 out", 55392, 3072, 1024, 2, False, 24),
    ("ViT out_proj 55392x1024x1024 +bias +fp32 residual", 55392, 1024, 1024, 4, True, 24),
    ("ViT c_fc 55392x4096x1024 +bias QuickGELU fp16 out", 55392, 4096, 1024, 2, False, 24),
    ("ViT c_proj 55392x1024x4096 +bias +fp32 residual", 55392, 1024, 4096, 4, True, 24),
    ("Phi-3 qkv 5960x9216x3072 fp16 out", 5960, 9216, 3072, 2, False, 32),
    ("Phi-3 o_proj 5960x3072x3072 +fp32 residual", 5960, 3072, 3072, 4, True, 32),
    ("Phi-3 gate_up 5960x16384x3072 SwiGLU fp16 out", 5960, 16384, 3072, 1, False, 32),   # output is [M, N/2] fp16
    ("Phi-3 down 5960x3072x8192 +fp32 residual", 5960, 3072, 8192, 4, True, 32),
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    d = json.load(open(src))
    out = {"source": src, "shapes": []}
    for (label, M, N, K, ob, res, n), l in zip(SHAPES, d["launches"]):
        alg = (M * K * 2 + N * K * 2 + M * N * ob + (M * N * 4 if res else 0)) / 1e6
        mb = float(l["dram__bytes_read.sum"]) + float(l["dram__bytes_write.sum"])
        assert d["units"]["dram__bytes_read.sum"] == "Mbyte", d["units"]["dram__bytes_read.sum"]
        out["shapes"].append({"label": label, "M": M, "N": N, "K": K, "launches_per_step": n, "dram_MB": round(mb, 1), "algorithmic_MB": round(alg, 1),
                              "ratio": round(mb / alg, 3), "us_under_ncu": float(l["gpu__time_duration.sum"]),
                              "tensor_pipe_pct_active": float(l["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])})
    json.dump(out, open(dst, "w"), indent=1)
    for s in out["shapes"]:
        print(s)


if __name__ == "__main__":
    main()
