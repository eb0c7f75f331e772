#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel.
Usage: tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n]   (per-launch times are cold-cache
and serialised: compare SHARES with bench.py's live event timing, not absolutes)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as fh:
    lines = [ln for ln in fh if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
rows = rows[skip:]
if "--step" in sys.argv:
    # one warm CHAINED step: the last window [patch_im2col .. decode_heatmaps] that holds exactly 30 launches (1 gather, 13 chains,
    # 12 attention, 2 deconv, 1x1 conv, decode)
    starts = [i for i, r in enumerate(rows) if "patch_im2col" in r[0] or "frame_to_patch_rows" in r[0]]
    pick = None
    for i in starts:
        j = next((k for k in range(i, min(i + 120, len(rows))) if "decode_heatmaps" in rows[k][0]), None)
        if j is not None and j - i + 1 == 30 and any("gemm_chain" in r[0] for r in rows[i:j + 1]):
            pick = (i, j)
    if pick is None:
        raise SystemExit("no 30-launch chained step found in the capture")
    print(f"one warm chained step = launches {skip + pick[0]}..{skip + pick[1]} of the capture")
    rows = rows[pick[0]:pick[1] + 1]
agg = defaultdict(lambda: [0, 0.0])
for name, ns, grid, block in rows:
    m = re.match(r"(?:void )?(?:vpb::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    key = (m.group(1) + (m.group(2) or "")) if m else name[:60]
    agg[key][0] += 1
    agg[key][1] += ns
total = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {total / 1e6:.3f} ms total (serialised, cold cache)")
print(f"{'kernel':58s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {n:8d} {ns / 1e3:10.1f} {ns / 1e3 / n:8.2f} {ns / total:7.1%}")
