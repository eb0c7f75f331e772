#!/usr/bin/env python
"""Latency of small ragged batches (BASELINE configs[4]: a video stream's per-frame crop batch), CUDA-graph replay on:
  wide      one kernel per GEMM / LayerNorm, 256-wide tiles always (round 1)
  narrow    the same with 128-wide tiles while they fit one wave (pick_tile, the default for small batches)
  narrow+ln-in-gemm  narrow + LayerNorm and its consumer GEMM (qkv / fc1) as one two-stage chained launch (option ln_in_gemm; slower, off)
  (LayerNorm in the TAIL of the residual GEMMs, option ln_fused, lost at every batch size: 1.03 vs 0.71 ms at 1 crop, 3.16 vs 2.73 at 64)
  chained   chained launches (chain.cuh; the default from 48 crops on)
Synchronous host-visible latency per call and back-to-back time per call (no sync between calls)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from easy_vitpose_b200 import ViTPose, dyn_model_import, _lib
from easy_vitpose_b200.synthetic import random_state_dict
m = ViTPose(dyn_model_import("ap10k", "b"), max_batch=64)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in random_state_dict("b", 17, seed=1).items()}).to("cuda:0")
m.set_option("chain_min_batch", 1)
L = _lib.lib()
side = torch.cuda.Stream()
MODES = [("wide", 16 << 8, 0, 0, 0), ("narrow", 0, 0, 0, 0), ("narrow+ln-in-gemm", 0, 0, 0, 1), ("chained", 0, 0, 1, 0)]
for n in (1, 2, 4, 6, 9, 12, 16, 24, 32, 48, 64):
    x = torch.randn(n, 3, 256, 192, device="cuda"); org = torch.tensor([[192, 256]] * n, dtype=torch.int32, device="cuda")
    out = []
    ref = None
    for name, flags, lnf, chain, lig in MODES:
        L.vpb_debug_gemm(flags, None)
        m.set_option("ln_fused", lnf)
        m.set_option("ln_in_gemm", lig)
        m.set_option("chain", chain)            # also drops the captured graphs: they embed the choices above
        with torch.cuda.stream(side):
            for _ in range(5):
                kp, _ = m.infer_crops(x, org)
            torch.cuda.synchronize()
            if ref is None:
                ref = kp.clone()
            same = bool(torch.equal(ref, kp))
            t0 = time.perf_counter()
            for _ in range(30):
                kp, _ = m.infer_crops(x, org); torch.cuda.synchronize()
            lat = (time.perf_counter() - t0) / 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                kp, _ = m.infer_crops(x, org)
            e1.record(); torch.cuda.synchronize()
            out.append(f"{name} {lat * 1e3:.3f} / {e0.elapsed_time(e1) / 30:.3f}{'' if same else ' (DIFFERS!)'}")
    print(f"crops/call={n:2d}  latency / back-to-back ms:  " + "  |  ".join(out))
L.vpb_debug_gemm(0, None); m.set_option("ln_fused", 0); m.set_option("ln_in_gemm", 0); m.set_option("chain", 1)
