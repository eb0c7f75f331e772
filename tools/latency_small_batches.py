#!/usr/bin/env python
"""Latency of small ragged batches (BASELINE configs[4]: a video stream's per-frame crop batch): chained launches (chain.cuh)
against one kernel per GEMM / LayerNorm, CUDA-graph replay on.  Synchronous host-visible latency per call, and back-to-back
throughput (no sync between calls) for the batch sizes around the chain's break-even point."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from easy_vitpose_b200 import ViTPose, dyn_model_import
from easy_vitpose_b200.synthetic import random_state_dict
m = ViTPose(dyn_model_import("ap10k", "b"), max_batch=64)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in random_state_dict("b", 17, seed=1).items()}).to("cuda:0")
m.set_option("chain_min_batch", 1)
side = torch.cuda.Stream()
for n in (1, 2, 6, 12, 16, 24, 32, 48, 64):
    x = torch.randn(n, 3, 256, 192, device="cuda"); org = torch.tensor([[192, 256]] * n, dtype=torch.int32, device="cuda")
    row = []
    for chain in (0, 1):
        m.set_option("chain", chain)
        with torch.cuda.stream(side):
            for _ in range(5):
                kp, _ = m.infer_crops(x, org)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(30):
                kp, _ = m.infer_crops(x, org); torch.cuda.synchronize()
            lat = (time.perf_counter() - t0) / 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                kp, _ = m.infer_crops(x, org)
            e1.record(); torch.cuda.synchronize()
            thr = e0.elapsed_time(e1) / 30
        row.append((lat * 1e3, thr))
    print(f"crops/call={n:2d}: unchained latency {row[0][0]:.3f} ms, back-to-back {row[0][1]:.3f} ms | chained latency {row[1][0]:.3f} ms, back-to-back {row[1][1]:.3f} ms"
          f" | chained/unchained back-to-back {row[1][1] / row[0][1]:.3f}")
