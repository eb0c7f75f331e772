#!/bin/bash
# Runs every -m gpu test in its own process (a trapped kernel poisons the CUDA context) with a timeout,
# collecting per-test logs under gpurun_out/bringup/.  Usage: tools/gpu_bringup.sh [pytest -k expr]
mkdir -p gpurun_out/bringup
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/bringup/gpu.txt 2>&1
python -m pytest tests -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::" > gpurun_out/bringup/tests.txt
: > gpurun_out/bringup/summary.txt
while read -r t; do
  name=$(echo "$t" | tr '/:[] ' '_____')
  timeout 300 python -m pytest "$t" -x -q -s -m gpu > "gpurun_out/bringup/$name.log" 2>&1
  rc=$?
  echo "$rc $t" >> gpurun_out/bringup/summary.txt
done < gpurun_out/bringup/tests.txt
cat gpurun_out/bringup/summary.txt
