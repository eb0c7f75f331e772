#!/bin/bash
mkdir -p gpurun_out/r2l
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -x -k "chain" > gpurun_out/r2l/pytest_chain.log 2>&1; rc=$?; echo "chain tests exit $rc"; tail -3 gpurun_out/r2l/pytest_chain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2l/chain_diag_64.log 2>&1; cat gpurun_out/r2l/chain_diag_64.log
for i in 1 2; do
for ch in 1 0; do
  VPB_CHAIN=$ch timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2l/bench_burst_chain${ch}_$i.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2l/bench_burst_chain${ch}_$i.json')); print('burst chain=$ch run $i', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
for ch in 1 0; do
VPB_CHAIN=$ch timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame-path > gpurun_out/r2l/bench_sust_chain$ch.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2l/bench_sust_chain$ch.json')); print('sustained chain=$ch', round(d['value']), d['ms_per_step'], d['clocks'])"
done
