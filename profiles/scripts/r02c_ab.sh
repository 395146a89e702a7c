# A/B of environment knobs / variant libraries on the config-3 step: lines "name ENV=.. ENV=.." on stdin (no test run)
mkdir -p gpurun_out
while read name envs; do
  [ -z "$name" ] && continue
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02c_$name.json 2> gpurun_out/r02c_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), d["gpu_launches"], {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.05}, d["parity_check"]["depth"])
except Exception as e:
    print("$name failed", e)
PY
done
