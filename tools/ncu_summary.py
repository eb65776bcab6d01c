"""Summarises an .ncu-rep (read here with `ncu -i`, no GPU needed) into a small JSON under profiles/:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json ["label 0" "label 1" ...]"""
import csv
import json
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__cluster_size", "launch__shared_mem_per_block_dynamic"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    labels = sys.argv[3:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(k) for k in KEYS if k in hdr}
    res = {"source": rep, "units": {k: units[i] for k, i in idx.items()}, "launches": []}
    for n, r in enumerate(rows[2:]):
        d = {k: r[i] for k, i in idx.items()}
        if n < len(labels):
            d["label"] = labels[n]
        res["launches"].append(d)
    json.dump(res, open(out, "w"), indent=1)
    print("wrote", out, len(res["launches"]), "launches")


if __name__ == "__main__":
    main()
