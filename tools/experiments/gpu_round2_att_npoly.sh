#!/bin/bash
# fraction of the packed kernel's exponentials on the FMA pipe: 4 / 8 / 12 / 16 of every 32 (standalone, then in-step)
mkdir -p gpurun_out/r2s
for n in 4 8 12 16; do
  echo "== NPOLY $n"
  VPB_ATT_NPOLY=$n timeout 200 python - <<'PY'
import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import ctypes as C, torch
from easy_vitpose_b200 import _lib
from gpu_util import attention
L = _lib.lib(); dev = torch.device("cuda", 0)
for heads, hd, B in ((12, 64, 64), (16, 64, 64)):
    D = heads * hd
    qkv = (torch.randn(B * 192, 3 * D, device=dev) * 0.5).bfloat16()
    out = torch.empty((B * 192, D), dtype=torch.bfloat16, device=dev)
    for _ in range(5): attention(qkv, B, heads, hd)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): L.vpb_attention(C.c_void_p(qkv.data_ptr()), B, heads, hd, C.c_void_p(out.data_ptr()), None)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 50)
    q, k, v = (qkv.float().reshape(B, 192, 3, heads, hd)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(B * 192, D)
    rel = float((out.float() - ref).norm() / ref.norm())
    print(f"  heads={heads} hd={hd} B={B}: {best:.2f} us/launch, rel err {rel:.5f}")
PY
done
for n in 8 12 8 12; do
  VPB_ATT_NPOLY=$n timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2s/bench_npoly$n.json 2> gpurun_out/r2s/bench_npoly$n.err
  python -c "
import json
d=json.load(open('gpurun_out/r2s/bench_npoly$n.json')); print('in-step NPOLY $n', round(d['value']), 'crops/s', round(d['ms_per_step'],4), 'attention', round(d['kernels']['attention']['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
