mkdir -p gpurun_out/r2s
for v in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do
  set -- $v
  VPB_ATT_POLY=$1 VPB_ATT_PACK=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2s/bench_p$1_k$2.json 2> gpurun_out/r2s/bench_p$1_k$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s/bench_p$1_k$2.json"))
    print("poly=$1 pack=$2", round(d["value"]), "crops/s", round(d["ms_per_step"],4), "ms  attention", round(d["kernels"]["attention"]["ms_per_step"],4), "chain", round(d["kernels"]["gemm_chain"]["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "parity", d["parity_check"]["batch_equals_single_crop_calls"])
except Exception as e: print("poly=$1 pack=$2 failed", e)
PY
done
for v in "0 0" "1 1"; do
  set -- $v
  VPB_ATT_POLY=$1 VPB_ATT_PACK=$2 timeout 300 python bench.py --config l25x64 --steps 20 --warmup 5 --no-cpu-baseline --no-frame-path > gpurun_out/r2s/bench_l_p$1_k$2.json 2> gpurun_out/r2s/bench_l_p$1_k$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s/bench_l_p$1_k$2.json"))
    print("ViT-L poly=$1 pack=$2", round(d["value"]), "crops/s", round(d["ms_per_step"],4), "ms  attention", round(d["kernels"]["attention"]["ms_per_step"],4), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("L poly=$1 pack=$2 failed", e)
PY
done
