#!/bin/bash
mkdir -p gpurun_out/r2t
for erf in 0 1 0 1; do
  VPB_GELU_ERF=$erf timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2t/bench_gelu_erf$erf.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2t/bench_gelu_erf$erf.json')); print('gelu_erf=$erf', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], 'chain us', round(d['kernels']['gemm_chain']['ms_per_step']*1000), 'unchained fc1 us', round(d['kernels_unchained']['gemm_fc1_gelu']['ms_per_step']*1000))"
done
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "shard_pipeline" 2>&1 | tail -2
bash tools/gpu_sanitizer.sh
