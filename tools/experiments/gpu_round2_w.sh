#!/bin/bash
# GPU call W: coalesced form of the load + add + store residual epilogue, 8-row LayerNorm jobs with the control warp
mkdir -p gpurun_out/r2w
timeout 420 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -x -k "residual_rmw or f32_add or chain_is_bit_identical" > gpurun_out/r2w/pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2w/pytest.log
timeout 240 python tools/experiments/rmw_lnctl_ab.py b 17 64 20 3 "0,1,16;1,1,16;0,1,8;1,1,8;0,0,16" > gpurun_out/r2w/ab_b.log 2>&1; echo "ab b exit $?"; tail -7 gpurun_out/r2w/ab_b.log
for v in "0 1 16" "1 1 16" "1 1 8"; do
  set -- $v
  VPB_RESID_RMW=$1 VPB_LN_CTL=$2 VPB_LN_JOB_ROWS=$3 timeout 120 python tools/chain_diag.py 64 > gpurun_out/r2w/chain_diag_r$1_c$2_j$3.log 2>&1; echo "chain_diag rmw=$1 ctl=$2 rows=$3 exit $?"; cat gpurun_out/r2w/chain_diag_r$1_c$2_j$3.log
done
timeout 200 python tools/experiments/rmw_lnctl_ab.py l 25 64 10 2 "0,1,16;1,1,16;1,1,8;0,0,16" > gpurun_out/r2w/ab_l.log 2>&1; echo "ab l exit $?"; tail -5 gpurun_out/r2w/ab_l.log
timeout 200 python tools/experiments/rmw_lnctl_ab.py b 17 8 30 2 "0,1,16;1,1,16;0,0,16" > gpurun_out/r2w/ab_b8.log 2>&1; echo "ab b B=8 exit $?"; tail -4 gpurun_out/r2w/ab_b8.log
