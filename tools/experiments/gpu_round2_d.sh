#!/bin/bash
# GPU call D: burst-length benches (the driver's own run length) chain on/off, then ncu --set full of the chain and attention kernels
mkdir -p gpurun_out/r2d
for ch in 1 0; do
  VPB_CHAIN=$ch timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2d/bench_burst_chain$ch.json 2> gpurun_out/r2d/bench_burst_chain$ch.err; echo "bench chain=$ch exit $?"
  python -c "
import json
d=json.load(open('gpurun_out/r2d/bench_burst_chain$ch.json')); print('burst chain=$ch', round(d['value']), d['ms_per_step'], d['clocks'], 'e2e', round(d['e2e']['value']), 'roofline', d['roofline']['kernel'], round(d['roofline']['achieved']), round(d['roofline']['frac'],3))"
done
out=gpurun_out/r2d
BENCH="python bench.py --config b17x64 --steps 2 --warmup 3 --no-cpu-baseline --no-frame-path"
VPB_CHAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_chain_tcgen05 -s 20 -c 1 -f -o $out/chain_block $BENCH > $out/chain_block.log 2>&1; echo "ncu chain rc=$?"
VPB_CHAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 20 -c 1 -f -o $out/attention $BENCH > $out/attention.log 2>&1; echo "ncu attention rc=$?"
ls -la $out
