#!/usr/bin/env python
"""One 1080p frame + 64 boxes through ViTPose.infer_frame a few times (CUDA graph off so that every kernel is a visible launch);
used under ncu to list the frame path's kernels (profiles/r1_launches_frame_path.txt)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from easy_vitpose_b200 import ViTPose, model_cfg
from easy_vitpose_b200.synthetic import random_state_dict

B, K = 64, 17
m = ViTPose(model_cfg("b", K), max_batch=B)
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in random_state_dict("b", K, seed=11).items()}).to("cuda:0")
m.set_option("graph", 0)
rs = np.random.RandomState(5)
frame = torch.from_numpy(rs.randint(0, 256, size=(1080, 1920, 3), dtype=np.uint8)).cuda()
w = rs.randint(90, 420, size=B); h = (w * rs.uniform(1.6, 2.6, size=B)).astype(np.int64)
x0 = rs.randint(0, 1820, size=B); y0 = rs.randint(0, 880, size=B)
boxes = torch.from_numpy(np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.int32)).cuda()
with torch.cuda.stream(torch.cuda.Stream()):
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        kp, idx = m.infer_frame(frame, boxes)
torch.cuda.synchronize()
print("ok", float(kp[..., 2].mean()))
