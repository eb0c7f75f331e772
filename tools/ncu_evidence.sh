#!/bin/bash
# Run under gpurun (ONE GPU): `ncu --set full` of one launch of every kernel class of the current build, each taken from a
# warm step of the real path (bench.py, ViT-B K=17 batch 64), plus the launch list of one warm step.
# Outputs gpurun_out/r2_ncu/<name>.ncu-rep and launches.csv; tools/ncu_summarize.py turns them into profiles/ here.
set -u
out=gpurun_out/r2_ncu; mkdir -p $out
BENCH="python bench.py --config b17x64 --steps 2 --warmup 3 --no-cpu-baseline --no-frame-path"
# ONLY="decode launches" restricts the run to the named captures (default: all, plus the ViT-H threshold study)
want() { [ -z "${ONLY:-}" ] || [[ " $ONLY " == *" $1 "* ]]; }
cap() {   # name, kernel regex, launches to skip, extra env
  name=$1; pat=$2; skip=$3; shift 3
  want $name || return 0
  env "$@" timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$pat" -s $skip -c 1 -f -o $out/$name $BENCH > $out/$name.log 2>&1
  echo "$name rc=$? $(ls -la $out/$name.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
}
# chained build: one warm forward = 1 gather + 13 chains + 12 attention + 2 deconv + 1 final + 1 decode
cap chain_block   'gemm_chain_tcgen05'          20 VPB_CHAIN=1     # a full block chain: proj -> LN -> fc1 -> fc2 -> LN -> qkv
cap attention     'attention_pack_tcgen05'      20 VPB_CHAIN=1
cap deconv        'gemm_bf16_tcgen05ILi256ELi2E'   4  VPB_CHAIN=1
cap final_conv    'gemm_bf16_tcgen05ILi32ELi4E'    2  VPB_CHAIN=1
cap decode        'decode_heatmaps'             2  VPB_CHAIN=1
cap patch_im2col  'patch_im2col'                2  VPB_CHAIN=1
# one kernel per GEMM / LayerNorm (the path the per-class roofline numbers come from)
cap gemm_qkv      'gemm_bf16_tcgen05ILi256ELi0E'   14 VPB_CHAIN=0
cap gemm_fc1      'gemm_bf16_tcgen05ILi256ELi1E'   14 VPB_CHAIN=0
cap gemm_fc2_proj 'gemm_bf16_tcgen05ILi256ELi5E'   29 VPB_CHAIN=0     # skip 29 -> an fc2 launch (patch, then proj/fc2 alternate)
cap layernorm     'layernorm_f32_to_bf16'       30 VPB_CHAIN=0
if want launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv $BENCH > $out/launches.log 2>&1
  echo "launch list rc=$?"
fi
ls -la $out
[ -n "${ONLY:-}" ] && exit 0
# ViT-H wholebody B=32: chained vs one kernel per GEMM (threshold study), burst-length runs
for mb in 1 999; do
  VPB_CHAIN_MIN_BATCH=$mb timeout 600 python bench.py --config h133x32 --steps 30 --warmup 5 --no-cpu-baseline --no-frame-path > $out/bench_h_minbatch$mb.json 2>/dev/null
  python -c "
import json
d=json.load(open('$out/bench_h_minbatch$mb.json')); print('ViT-H B=32 chain_min_batch=$mb', round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d['gpu_launches']/d['steps'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/bench_driver_len2.json 2>/dev/null
python -c "
import json
d=json.load(open('$out/bench_driver_len2.json')); print('driver-length', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'single', round(d['e2e']['single_call_value']), 'frame_path', round(d['frame_path']['value']))"
