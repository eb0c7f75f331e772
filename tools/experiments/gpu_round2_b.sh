#!/bin/bash
# GPU call B (round 2): attention v6 (ping-pong softmax groups) + first run of the chained GEMM launches
mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "attention or gelu" > gpurun_out/r2b/pytest_attn.log 2>&1; echo "attn pytest exit $?"
tail -3 gpurun_out/r2b/pytest_attn.log
timeout 300 python tools/attn_diag.py > gpurun_out/r2b/attn_diag.log 2>&1; echo "diag exit $?"; cat gpurun_out/r2b/attn_diag.log
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -x -k "chain" > gpurun_out/r2b/pytest_chain.log 2>&1; rc=$?; echo "chain pytest exit $rc"
tail -15 gpurun_out/r2b/pytest_chain.log
if [ $rc -ne 0 ]; then export VPB_CHAIN=0; echo "chain failed: rest of the run with VPB_CHAIN=0"; fi
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_engine.py::test_chain_is_bit_identical > gpurun_out/r2b/pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r2b/pytest.log
for ch in 1 0; do
  [ $rc -ne 0 ] && [ $ch -eq 1 ] && continue
  VPB_CHAIN=$ch timeout 600 python bench.py --config b17x64 --steps 100 --warmup 10 --no-cpu-baseline --no-frame-path > gpurun_out/r2b/bench_b17x64_chain$ch.json 2> gpurun_out/r2b/bench_b17x64_chain$ch.err; echo "bench chain=$ch exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2b/bench_b17x64_chain$ch.json"))
print("chain=$ch", round(d["value"]), d["ms_per_step"], d["clocks"], "e2e", round(d["e2e"]["value"]))
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]): print(f"  {k:18s} {v['ms_per_step']*1000:8.1f} us n={v['launches_per_step']:.0f} tflops {v.get('tflops',0):.0f}")
PY
done
