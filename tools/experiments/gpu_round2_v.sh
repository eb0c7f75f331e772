#!/bin/bash
# GPU call V: residual epilogue as load + add + store (resid_rmw) and the LayerNorm control warp (ln_ctl): bit identity, A/B, cycle counters
mkdir -p gpurun_out/r2v
timeout 420 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -m gpu -q -x -k "residual_rmw or f32_add or chain_is_bit_identical" > gpurun_out/r2v/pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2v/pytest.log
timeout 200 python tools/experiments/rmw_lnctl_ab.py b 17 64 20 3 > gpurun_out/r2v/ab_b.log 2>&1; echo "ab b exit $?"; tail -8 gpurun_out/r2v/ab_b.log
for v in "0 0" "1 1"; do
  set -- $v
  VPB_RESID_RMW=$1 VPB_LN_CTL=$2 timeout 120 python tools/chain_diag.py 64 > gpurun_out/r2v/chain_diag_r$1_c$2.log 2>&1; echo "chain_diag rmw=$1 ctl=$2 exit $?"; cat gpurun_out/r2v/chain_diag_r$1_c$2.log
done
timeout 200 python tools/experiments/rmw_lnctl_ab.py l 25 64 10 2 > gpurun_out/r2v/ab_l.log 2>&1; echo "ab l exit $?"; tail -5 gpurun_out/r2v/ab_l.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
