#!/bin/bash
mkdir -p gpurun_out/r2u
timeout 600 python tools/latency_small_batches.py > gpurun_out/r2u/latency.log 2>&1; cat gpurun_out/r2u/latency.log
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -k "batch_invariance or ragged or host_api or fused_layernorm or chain_is_bit" 2>&1 | tail -2
timeout 600 python bench.py --config ap10k-streams --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2u/bench_streams.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2u/bench_streams.json')); print('streams', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step']): print(f\"  {k:18s} {v['ms_per_step']*1000:8.1f} us n={v['launches_per_step']:.0f} per-launch {v['ms_per_step']*1000/v['launches_per_step']:.1f}\")"
