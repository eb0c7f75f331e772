#!/bin/bash
# GPU call X: LayerNorm control warp with the "ready" hand-off on named barriers (LayerNorm warps asleep, not polling)
mkdir -p gpurun_out/r2x
timeout 200 python tools/experiments/rmw_lnctl_ab.py b 17 64 20 5 "0,0,16;0,1,16" > gpurun_out/r2x/ab_b.log 2>&1; echo "ab b exit $?"; tail -4 gpurun_out/r2x/ab_b.log
VPB_LN_CTL=1 timeout 120 python tools/chain_diag.py 64 > gpurun_out/r2x/chain_diag_c1.log 2>&1; echo "chain_diag ctl=1 exit $?"; cat gpurun_out/r2x/chain_diag_c1.log
VPB_LN_CTL=0 timeout 120 python tools/chain_diag.py 64 > gpurun_out/r2x/chain_diag_c0.log 2>&1; echo "chain_diag ctl=0 exit $?"; cat gpurun_out/r2x/chain_diag_c0.log
