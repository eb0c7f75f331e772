#!/bin/bash
# GPU call C (round 2): chained launches + attention v6 after the p_ready phase fix
mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention" > gpurun_out/r2c/pytest_attn.log 2>&1; echo "attn pytest exit $?"
timeout 300 python tools/attn_diag.py > gpurun_out/r2c/attn_diag.log 2>&1; echo "diag exit $?"; cat gpurun_out/r2c/attn_diag.log
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_batch_parity.py tests/test_gpu_coco_ap.py -m gpu -q -s > gpurun_out/r2c/pytest_engine.log 2>&1; echo "engine pytest (chain on) exit $?"
tail -4 gpurun_out/r2c/pytest_engine.log
for ch in 1 0; do
  VPB_CHAIN=$ch timeout 600 python bench.py --config b17x64 --steps 100 --warmup 10 --no-cpu-baseline --no-frame-path > gpurun_out/r2c/bench_b17x64_chain$ch.json 2> gpurun_out/r2c/bench_b17x64_chain$ch.err; echo "bench chain=$ch exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c/bench_b17x64_chain$ch.json"))
    print("chain=$ch", round(d["value"]), d["ms_per_step"], d["clocks"], "e2e", round(d["e2e"]["value"]), "parity", d["parity_check"]["batch_equals_single_crop_calls"])
    for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]): print(f"  {k:18s} {v['ms_per_step']*1000:8.1f} us n={v['launches_per_step']:.0f} tflops {v.get('tflops',0):.0f}")
except Exception as e: print("no json", e)
PY
done
VPB_ATT_POLY=1 timeout 600 python bench.py --config b17x64 --steps 100 --warmup 10 --no-cpu-baseline --no-frame-path > gpurun_out/r2c/bench_b17x64_poly.json 2> gpurun_out/r2c/bench_b17x64_poly.err; echo "bench poly exit $?"
python -c "
import json
d=json.load(open('gpurun_out/r2c/bench_b17x64_poly.json')); print('poly', round(d['value']), d['ms_per_step'], 'attention us', d['kernels']['attention']['ms_per_step']*1000)"
timeout 600 python bench.py --config ap10k-streams --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c/bench_streams.json 2> gpurun_out/r2c/bench_streams.err; echo "bench streams exit $?"
python -c "
import json
d=json.load(open('gpurun_out/r2c/bench_streams.json')); print('streams', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"
