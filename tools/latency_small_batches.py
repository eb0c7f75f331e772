#!/usr/bin/env python
"""Latency of small ragged batches (BASELINE configs[4]: a video stream's per-frame crop batch), with and without CUDA-graph
replay of the kernel chain.  Synchronous host-visible latency per call."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from easy_vitpose_b200 import ViTPose, dyn_model_import
from easy_vitpose_b200.synthetic import random_state_dict
m = ViTPose(dyn_model_import("ap10k", "b"), max_batch=32)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in random_state_dict("b", 17, seed=1).items()}).to("cuda:0")
for graph in (0, 1):
    m.set_option("graph", graph)
    for n in (1, 2, 6, 16, 32):
        x = torch.randn(n, 3, 256, 192, device="cuda"); org = torch.tensor([[192, 256]] * n, dtype=torch.int32, device="cuda")
        for _ in range(5):
            kp, _ = m.infer_crops(x, org); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            kp, _ = m.infer_crops(x, org); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        print(f"graph={graph} crops/frame={n:2d}: {dt*1e3:.3f} ms/frame  ({n/dt:.0f} crops/s)")
