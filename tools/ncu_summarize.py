#!/usr/bin/env python
"""Turn the `ncu --set full` captures of tools/ncu_evidence.sh (gpurun_out/r2_ncu/*.ncu-rep) into the committed evidence:
profiles/r2_ncu_full.txt (per kernel: duration, tensor-pipe %, DRAM bytes and GB/s, L2 hit rate, issue %, registers, smem, grid)
and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py's `roofline.traffic`).  Runs without a GPU."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r2_ncu")
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tensor_op_hmma.sum"]
TO_CLASS = {"chain_block": "gemm_chain", "attention": "attention", "gemm_qkv": "gemm_qkv", "gemm_fc1": "gemm_fc1_gelu", "gemm_fc2_proj": "gemm_fc2",
            "deconv": "gemm_deconv", "final_conv": "gemm_final_conv", "decode": "decode", "layernorm": "layernorm", "patch_im2col": "patch_im2col"}


def to_bytes(v: float, unit: str) -> float:
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main() -> None:
    out_lines = ["round 2: ncu --set full --clock-control none --import-source on, one launch of each kernel class taken from a WARM step of the real path",
                 "(python bench.py --config b17x64, ViT-B K=17, 64 crops; tools/ncu_evidence.sh).  ncu replays the launch with caches flushed:",
                 "DRAM bytes are the cold-cache figure.  Captures: gpurun_out/r2_ncu/*.ncu-rep (scratch, not committed).", ""]
    traffic = {}
    for name in sorted(TO_CLASS):
        rep = os.path.join(SRC, name + ".ncu-rep")
        if not os.path.exists(rep):
            out_lines.append(f"== {name}: no capture")
            continue
        res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(res.stdout)))
        if len(rows) < 3:
            out_lines.append(f"== {name}: empty report")
            continue
        head, units, vals = rows[0], rows[1], rows[2]
        col = {h: i for i, h in enumerate(head)}
        out_lines.append(f"== {name}   [{vals[col['Kernel Name']][:110]}]")
        rec = {}
        for m in WANT:
            if m in col:
                out_lines.append(f"   {m:78s} {units[col[m]]:14s} {vals[col[m]]}")
                try:
                    rec[m] = (float(vals[col[m]].replace(",", "")), units[col[m]])
                except ValueError:
                    pass
        if "dram__bytes_read.sum" in rec and "dram__bytes_write.sum" in rec and "gpu__time_duration.sum" in rec:
            b = to_bytes(*rec["dram__bytes_read.sum"]) + to_bytes(*rec["dram__bytes_write.sum"])
            t_us = rec["gpu__time_duration.sum"][0] * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(rec["gpu__time_duration.sum"][1], 1)
            traffic[TO_CLASS[name]] = b
            out_lines.append(f"   -> DRAM traffic {b / 1e6:.2f} MB per launch, {b / t_us / 1e3:.1f} GB/s achieved under ncu (duration {t_us:.1f} us)")
        out_lines.append("")
    with open(os.path.join(ROOT, "profiles", "r2_ncu_full.txt"), "w") as fh:
        fh.write("\n".join(out_lines) + "\n")
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
        json.dump({"source": "profiles/r2_ncu_full.txt (ncu --set full, cold cache, per launch, round-2 final build)",
                   "dram_bytes_per_launch": traffic}, fh, indent=1)
    print("\n".join(out_lines[:60]))


if __name__ == "__main__":
    main()
