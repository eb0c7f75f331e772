#!/bin/bash
# Final validation of the round-2 build: the whole `-m gpu` suite, smoke(), the driver-length bench (N = 1)
mkdir -p gpurun_out/r2z
timeout 330 python -m pytest tests -m gpu -q -x > gpurun_out/r2z/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2z/pytest.log
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2z/bench_driver_len.json 2> gpurun_out/r2z/bench_driver_len.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2z/bench_driver_len.json"))
    r = d["roofline"]
    print(round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms | e2e", round(d["e2e"]["value"]), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
          "| roofline", r["kernel"], round(r["achieved"]), round(r["frac"], 3), "whole", round(r["whole_step_tflops"]), "| launches/step", d["gpu_launches"] / d["steps"],
          "| parity", (d.get("parity_check") or {}).get("batch_equals_single_crop_calls"))
    print({k: round(v["ms_per_step"], 4) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 60 python __graft_entry__.py smoke > gpurun_out/r2z/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2z/smoke.log
