# configs 4 / 5 quick A/B: "name ENV=.." lines on stdin; prints step and the big kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "soup or rejects or k7b_small or sub_range or big_tri or c4_sphere" > gpurun_out/r02_c45_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_c45_test.log
tail -3 gpurun_out/r02_c45_test.log
while read name envs; do
  [ -z "$name" ] && continue
  for wl in c4 c5; do
    env $envs timeout 400 python bench.py --workload $wl --c5-tris 20000000 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02_${name}_$wl.json 2> gpurun_out/r02_${name}_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_${name}_$wl.json"))
    print("$name $wl", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.03}, d["parity_check"].get("depth"))
except Exception as e:
    print("$name $wl failed", e)
PY
  done
done
