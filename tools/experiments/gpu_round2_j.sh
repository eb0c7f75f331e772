#!/bin/bash
mkdir -p gpurun_out/r2j
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -k "chain" > gpurun_out/r2j/pytest_chain.log 2>&1; echo "chain tests exit $?"; tail -3 gpurun_out/r2j/pytest_chain.log
timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2j/chain_diag_64.log 2>&1; cat gpurun_out/r2j/chain_diag_64.log
for i in 1 2; do
for ch in 1 0; do
  VPB_CHAIN=$ch timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2j/bench_burst_chain${ch}_$i.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2j/bench_burst_chain${ch}_$i.json')); print('burst chain=$ch run $i', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
VPB_CHAIN=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame-path > gpurun_out/r2j/bench_sust_chain1.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2j/bench_sust_chain1.json')); print('sustained chain=1', round(d['value']), d['ms_per_step'], d['clocks'])"
