#!/bin/bash
mkdir -p gpurun_out/r2h
VPB_F32_RED=1 timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -k "chain" > gpurun_out/r2h/pytest_chain_red.log 2>&1; echo "chain tests (red epilogue) exit $?"; tail -3 gpurun_out/r2h/pytest_chain_red.log
VPB_F32_RED=1 timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2h/chain_diag_64_red.log 2>&1; cat gpurun_out/r2h/chain_diag_64_red.log
for red in 1 0; do
  VPB_F32_RED=$red timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2h/bench_burst_red$red.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2h/bench_burst_red$red.json')); print('burst red=$red', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
VPB_F32_RED=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-frame-path > gpurun_out/r2h/bench_sust_red1.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2h/bench_sust_red1.json')); print('sustained red=1', round(d['value']), d['ms_per_step'], d['clocks'])"
