#!/bin/bash
mkdir -p gpurun_out/r2g
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -k "chain or gelu" > gpurun_out/r2g/pytest_chain.log 2>&1; echo "chain+gelu tests exit $?"; tail -3 gpurun_out/r2g/pytest_chain.log
timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2g/chain_diag_64.log 2>&1; cat gpurun_out/r2g/chain_diag_64.log
for ch in 1 0; do
  VPB_CHAIN=$ch timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2g/bench_burst_chain$ch.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2g/bench_burst_chain$ch.json')); print('burst chain=$ch', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e', round(d['e2e']['value']))"
done
VPB_CHAIN=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame-path > gpurun_out/r2g/bench_sust_chain1.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2g/bench_sust_chain1.json')); print('sustained chain=1', round(d['value']), d['ms_per_step'], d['clocks'])"
timeout 600 python tools/latency_small_batches.py > gpurun_out/r2g/latency.log 2>&1; cat gpurun_out/r2g/latency.log
