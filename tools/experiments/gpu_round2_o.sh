#!/bin/bash
# GPU call O (1 GPU): the ncu captures that failed on quoting + the new decode kernel
set -u
out=gpurun_out/r2_ncu; mkdir -p $out
BENCH="python bench.py --config b17x64 --steps 2 --warmup 3 --no-cpu-baseline --no-frame-path"
cap() { name=$1; pat=$2; skip=$3; shift 3
  env "$@" timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$pat" -s $skip -c 1 -f -o $out/$name $BENCH > $out/$name.log 2>&1
  echo "$name rc=$? $(ls -la $out/$name.ncu-rep 2>/dev/null | awk '{print $5}') bytes"; }
cap deconv        'gemm_bf16_tcgen05ILi256ELi2E'   5  VPB_CHAIN=1
cap final_conv    'gemm_bf16_tcgen05ILi32ELi4E'    2  VPB_CHAIN=1
cap gemm_qkv      'gemm_bf16_tcgen05ILi256ELi0E'   14 VPB_CHAIN=0
cap gemm_fc1      'gemm_bf16_tcgen05ILi256ELi1E'   14 VPB_CHAIN=0
cap gemm_fc2_proj 'gemm_bf16_tcgen05ILi256ELi5E'   29 VPB_CHAIN=0
