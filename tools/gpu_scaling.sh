#!/bin/bash
# GPU call P (8 GPUs of one box): scaling N = 1, 2, 4, 8 back to back (configs[1], weak scaling), then configs[3] and configs[4] at N = 8
mkdir -p gpurun_out/r2p
run() { n=$1; tag=$2; shift 2
  if [ $n -eq 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/r2p/$tag.json 2> gpurun_out/r2p/$tag.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > gpurun_out/r2p/$tag.json 2> gpurun_out/r2p/$tag.err; fi
  echo "$tag exit $?"
  python -c "
import json
d=json.load(open('gpurun_out/r2p/$tag.json')); print('$tag', 'N', d['n_gpus'], round(d['value']), 'crops/s', round(d['ms_per_step'],3), 'ms | e2e', round(d['e2e']['value']), '| clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 gpurun_out/r2p/$tag.err
}
run 1 scale_n1 --steps 100 --warmup 10 --no-cpu-baseline --no-frame-path
run 2 scale_n2 --steps 100 --warmup 10
run 4 scale_n4 --steps 100 --warmup 10
run 8 scale_n8 --steps 100 --warmup 10
run 8 scale_n8_l25x64 --config l25x64 --steps 100 --warmup 10
run 8 scale_n8_streams --config ap10k-streams --steps 20 --warmup 5
