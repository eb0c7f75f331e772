#!/bin/bash
mkdir -p gpurun_out/r2q
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -x -k "chain" > gpurun_out/r2q/pytest_chain.log 2>&1; rc=$?; echo "chain tests (wavefront order) exit $rc"; tail -3 gpurun_out/r2q/pytest_chain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2q/chain_diag_64.log 2>&1; cat gpurun_out/r2q/chain_diag_64.log
for lags in "100 100" "16 22" "8 12" "24 32" "16 100" "100 22"; do
  set -- $lags
  VPB_CHAIN_LAG0=$1 VPB_CHAIN_LAG1=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2q/bench_lag_$1_$2.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2q/bench_lag_$1_$2.json')); print('burst lag $1 $2:', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'], 'chain us/launch', round(d['kernels']['gemm_chain']['ms_per_step']*1000/13,1))"
done
for lags in "100 100" "16 22"; do
  set -- $lags
  VPB_CHAIN_LAG0=$1 VPB_CHAIN_LAG1=$2 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame-path > gpurun_out/r2q/bench_sust_lag_$1_$2.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2q/bench_sust_lag_$1_$2.json')); print('sustained lag $1 $2:', round(d['value']), d['ms_per_step'], d['clocks'])"
done
