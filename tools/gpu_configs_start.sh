#!/bin/bash
# GPU call A (round 2): full GPU test-suite + the four BASELINE configs at N=1
mkdir -p gpurun_out/r2a
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
for c in b17x64 h133x32 l25x64 ap10k-streams; do
  steps=100; [ "$c" = "ap10k-streams" ] && steps=20
  timeout 600 python bench.py --config $c --steps $steps --warmup 10 > gpurun_out/r2a/bench_$c.json 2> gpurun_out/r2a/bench_$c.err
  echo "$c exit $?"; head -c 600 gpurun_out/r2a/bench_$c.json; echo
done
