#!/bin/bash
# GPU call E (2 GPUs): N=2 bench (ShardPipeline e2e, side-stream gather); on GPU 0 alone: chain timing experiment + small-batch latency chain on/off
mkdir -p gpurun_out/r2e
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2e/bench_n2.json 2> gpurun_out/r2e/bench_n2.err; echo "n2 exit $?"
python -c "
import json
d=json.load(open('gpurun_out/r2e/bench_n2.json')); print('N=2', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['api'][:60], d['clocks'])"
tail -3 gpurun_out/r2e/bench_n2.err
for nw in 0 1; do
  VPB_CHAIN_NOWAIT=$nw timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2e/bench_nowait$nw.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2e/bench_nowait$nw.json')); print('nowait=$nw', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], 'chain us', d['kernels']['gemm_chain']['ms_per_step']*1000/13)"
done
timeout 600 python tools/latency_small_batches.py > gpurun_out/r2e/latency.log 2>&1; cat gpurun_out/r2e/latency.log
