#!/bin/bash
# GPU call Y: residual epilogue as load + add + store entirely in the generic proxy (no TMA store): bit identity, A/B, counters
mkdir -p gpurun_out/r2y
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -x -k "residual_rmw or f32_add" > gpurun_out/r2y/pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2y/pytest.log
timeout 200 python tools/experiments/rmw_lnctl_ab.py b 17 64 20 4 "0,1,16;1,1,16" > gpurun_out/r2y/ab_b.log 2>&1; echo "ab b exit $?"; tail -4 gpurun_out/r2y/ab_b.log
VPB_RESID_RMW=1 timeout 120 python tools/chain_diag.py 64 > gpurun_out/r2y/chain_diag_r1.log 2>&1; echo "chain_diag rmw=1 exit $?"; cat gpurun_out/r2y/chain_diag_r1.log
timeout 200 python tools/experiments/rmw_lnctl_ab.py b 17 8 30 2 "0,1,16;1,1,16" > gpurun_out/r2y/ab_b8.log 2>&1; echo "ab b B=8 exit $?"; tail -3 gpurun_out/r2y/ab_b8.log
