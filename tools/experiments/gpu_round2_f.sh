#!/bin/bash
mkdir -p gpurun_out/r2f
timeout 300 python tools/chain_diag.py 64 > gpurun_out/r2f/chain_diag_64.log 2>&1; cat gpurun_out/r2f/chain_diag_64.log
timeout 300 python tools/chain_diag.py 8 > gpurun_out/r2f/chain_diag_8.log 2>&1; cat gpurun_out/r2f/chain_diag_8.log
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -s -k "gelu" > gpurun_out/r2f/pytest_gelu.log 2>&1; echo "gelu tests exit $?"; grep -E "fit vs erf|heatmap Linf vs fp32|passed|failed" gpurun_out/r2f/pytest_gelu.log
