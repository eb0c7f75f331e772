#!/bin/bash
mkdir -p gpurun_out/r2k
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -x -k "chain" > gpurun_out/r2k/pytest_chain.log 2>&1; rc=$?; echo "chain tests exit $rc"; tail -3 gpurun_out/r2k/pytest_chain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2k/chain_diag_64.log 2>&1; cat gpurun_out/r2k/chain_diag_64.log
for stg in 0 1; do
for i in 1 2; do
  VPB_CHAIN_STG2=$stg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2k/bench_burst_stg${stg}_$i.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2k/bench_burst_stg${stg}_$i.json')); print('burst chain stg2=$stg run $i', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
VPB_CHAIN_STG2=1 timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2k/chain_diag_64_stg2.log 2>&1; cat gpurun_out/r2k/chain_diag_64_stg2.log
VPB_CHAIN_STG2=1 timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -x -k "chain" > gpurun_out/r2k/pytest_chain_stg2.log 2>&1; echo "chain tests stg2 exit $?"
for stg in 0 1; do
VPB_CHAIN_STG2=$stg timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame-path > gpurun_out/r2k/bench_sust_stg$stg.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2k/bench_sust_stg$stg.json')); print('sustained stg2=$stg', round(d['value']), d['ms_per_step'], d['clocks'])"
done
