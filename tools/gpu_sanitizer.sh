#!/bin/bash
# compute-sanitizer memcheck over the kernel tests and the chained-launch / attention engine tests (small shapes: the tool slows
# kernels down by two orders of magnitude; the in-kernel hang guards count spins, not time, so they do not fire)
mkdir -p gpurun_out/r2san
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -x \
  -k "attention or gelu or layernorm or decode or chain_is_bit_identical or shard_pipeline or host_api or two_engines" > gpurun_out/r2san/memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/r2san/memcheck.log | tail -8
