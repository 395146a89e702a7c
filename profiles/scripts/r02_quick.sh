# quick loop: span-sensitive parity subset + c3 bench variants given as "name ENV=.. ENV=.." lines on stdin
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not c5 and not c4 and not full_size" > gpurun_out/r02_quick_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_quick_test.log
tail -4 gpurun_out/r02_quick_test.log
while read name envs; do
  [ -z "$name" ] && continue
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_$name.json"))
    print("$name", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.05}, d["parity_check"]["depth"])
except Exception as e:
    print("$name failed", e)
PY
done
