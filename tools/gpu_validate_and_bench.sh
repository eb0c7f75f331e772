#!/bin/bash
# GPU call M: full GPU test suite with the final code, smoke(), then the four BASELINE configs + the driver-length run
mkdir -p gpurun_out/r2m
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2m/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2m/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2m/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2m/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m/bench_driver_len.json 2> gpurun_out/r2m/bench_driver_len.err; echo "bench driver-length exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m/bench_reference_arm.json 2> gpurun_out/r2m/bench_reference_arm.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/r2m/bench_reference_arm.json
for c in b17x64 h133x32 l25x64 ap10k-streams; do
  steps=200; [ "$c" = "ap10k-streams" ] && steps=30
  timeout 900 python bench.py --config $c --steps $steps --warmup 10 > gpurun_out/r2m/bench_$c.json 2> gpurun_out/r2m/bench_$c.err; echo "$c exit $?"
done
python - <<'PY'
import json
for c in ["driver_len","b17x64","h133x32","l25x64","ap10k-streams"]:
    try:
        d=json.load(open(f"gpurun_out/r2m/bench_{c}.json"))
        r=d["roofline"]
        print(c, round(d["value"]), "crops/s", round(d["ms_per_step"],3), "ms | e2e", round(d["e2e"]["value"]), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "| roofline", r["kernel"], round(r["achieved"]), round(r["frac"],3), "whole", round(r["whole_step_tflops"]), "| launches", d["gpu_launches"]/d["steps"], "| parity", (d.get("parity_check") or {}).get("batch_equals_single_crop_calls"))
    except Exception as e: print(c, "failed", e)
PY
