# split bins: parity subset (incl. the self-checking build), then config 4 for several TRB_WARP_MAX / TRB_SPLIT_S, config 3 check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_bins or indexing_invariants or both_raster or c4_sphere or bin_overflow" > gpurun_out/r02b_split_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_split_test.log
tail -5 gpurun_out/r02b_split_test.log
printf 'base TRB_SPLIT=0\nw1024s256 TRB_WARP_MAX=1024\nw512s256 TRB_WARP_MAX=512\nw256s256 TRB_WARP_MAX=256\nw256s128 TRB_WARP_MAX=256 TRB_SPLIT_S=128\nw128s128 TRB_WARP_MAX=128 TRB_SPLIT_S=128\nw512s512 TRB_WARP_MAX=512 TRB_SPLIT_S=512\n' | while read name envs; do
  env $envs timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02b_${name}_c4.json 2> gpurun_out/r02b_${name}_c4.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02b_${name}_c4.json"))
    print("$name c4", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.02}, d["parity_check"].get("depth"))
except Exception as e:
    print("$name c4 failed", e)
PY
done
printf 'c3base TRB_SPLIT=0\nc3w256 TRB_WARP_MAX=256\nc3w1024 TRB_WARP_MAX=1024\n' | while read name envs; do
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02b_${name}.json 2> gpurun_out/r02b_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02b_${name}.json"))
    print("$name", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.015}, d["parity_check"].get("depth"))
except Exception as e:
    print("$name failed", e)
PY
done
