#!/usr/bin/env python
"""Probe: does running two independent half batches on two streams (kernels of one filling the SMs while the other's
kernel drains) beat one full batch?  Two engines (own workspaces), B/2 crops each, vs one engine with B crops."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from easy_vitpose_b200 import ViTPose, model_cfg
from oracle import vitpose_oracle as O
sd = {k: torch.from_numpy(np.asarray(v)) for k, v in O.make_state_dict(768, 12, 17, 1, peaky=0.1, bumps=True).items()}
def make(B):
    m = ViTPose(model_cfg("b", 17), max_batch=B); m.load_state_dict(sd).to("cuda:0"); return m
B = 64
def run(models, streams, xs, orgs, steps):
    for i in range(steps):
        for m, s, x, o in zip(models, streams, xs, orgs):
            with torch.cuda.stream(s):
                m.infer_crops(x, o)
for nsplit in (1, 2, 4):
    b = B // nsplit
    models = [make(b) for _ in range(nsplit)]
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    xs = [torch.randn(b, 3, 256, 192, device="cuda") for _ in range(nsplit)]
    orgs = [torch.tensor([[192, 256]] * b, dtype=torch.int32, device="cuda") for _ in range(nsplit)]
    run(models, streams, xs, orgs, 10); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(models, streams, xs, orgs, 100); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{nsplit} stream(s) x {b} crops: {B * 100 / dt:.0f} crops/s  ({dt * 10:.3f} ms per {B} crops)")
    del models
