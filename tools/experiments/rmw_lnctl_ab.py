#!/usr/bin/env python
"""A/B of the two round-2 restructurings of the chained launches on one box: residual epilogues as load + add + TMA store
(option resid_rmw) and the LayerNorm control warp (option ln_ctl).  For every combination: bit identity of the heatmaps at the
metric's batch (ViT-B, 64 crops) against the (0, 0) build, then device-event time of `steps` forwards + decodes, the
combinations interleaved and repeated so that clock drift shows up as spread, not as a difference."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from easy_vitpose_b200 import ViTPose, model_cfg
from easy_vitpose_b200.synthetic import random_state_dict

size = sys.argv[1] if len(sys.argv) > 1 else "b"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 17
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3

m = ViTPose(model_cfg(size, K), max_batch=B)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in random_state_dict(size, K, seed=1).items()}).to("cuda:0")
xs = [torch.randn(B, 3, 256, 192, device="cuda") for _ in range(4)]          # 4 x 37.7 MB (B = 64): rotated, like bench.py
org = torch.tensor([[192, 256]] * B, dtype=torch.int32)
# (resid_rmw, ln_ctl, ln_job_rows); the first one is the reference for the bit-identity check
combos = [tuple(int(v) for v in c.split(",")) for c in sys.argv[6].split(";")] if len(sys.argv) > 6 else \
    [(0, 0, 16), (1, 0, 16), (0, 1, 16), (1, 1, 16), (0, 1, 8), (1, 1, 8)]


def setopt(rmw, ctl, rows):
    m.set_option("resid_rmw", rmw)
    m.set_option("ln_ctl", ctl)
    m.set_option("ln_job_rows", rows)


base = None
for rmw, ctl, rows in combos:
    setopt(rmw, ctl, rows)
    hm = m(xs[0]).cpu().numpy()
    kp, idx = m.infer_crops(xs[0], org)
    kp, idx = kp.cpu().numpy(), idx.cpu().numpy()
    if base is None:
        base = (hm, kp, idx)
    same = np.array_equal(hm, base[0]) and np.array_equal(kp, base[1]) and np.array_equal(idx, base[2])
    print(f"resid_rmw={rmw} ln_ctl={ctl} ln_job_rows={rows}: heatmaps / keypoints / argmax bit-identical to {combos[0]}: {same}", flush=True)
    assert same

res = {c: [] for c in combos}
for rep in range(reps):
    for rmw, ctl, rows in combos:
        setopt(rmw, ctl, rows)
        for i in range(5):
            m.infer_crops(xs[i % 4], org)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            m.infer_crops(xs[i % 4], org)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        res[(rmw, ctl, rows)].append(ms)
        time.sleep(0.3)
for c in combos:
    v = res[c]
    print(f"ViT-{size} K={K} B={B} resid_rmw={c[0]} ln_ctl={c[1]} ln_job_rows={c[2]}: ms/step {' '.join(f'{t:.4f}' for t in v)}  -> best {min(v):.4f} ms = {B / min(v) * 1e3:.0f} crops/s, "
          f"median {sorted(v)[len(v) // 2]:.4f} ms = {B / sorted(v)[len(v) // 2] * 1e3:.0f} crops/s", flush=True)
