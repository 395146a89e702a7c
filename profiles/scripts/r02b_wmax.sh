# one GPU: TRB_WARP_MAX sweep on config 4 (and config 3 for the two candidates)
mkdir -p gpurun_out
for wm in 1024 512 256 128 64; do
TRB_WARP_MAX=$wm timeout 300 python bench.py --workload c4 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02b_c4_wm$wm.json 2> gpurun_out/r02b_c4_wm$wm.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02b_c4_wm$wm.json"))
    print("c4 warp_max $wm", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.02}, d["parity_check"].get("depth"))
except Exception as e:
    print("failed", e)
PY
done
for wm in 512 256; do
TRB_WARP_MAX=$wm timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02b_c3_wm$wm.json 2> gpurun_out/r02b_c3_wm$wm.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02b_c3_wm$wm.json"))
    print("c3 warp_max $wm", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.05}, d["parity_check"].get("depth"))
except Exception as e:
    print("failed", e)
PY
done
