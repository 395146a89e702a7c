# launch diet check: full parity suite, then config 3 / 1 / 2 / 4 / 5 step times (device-timed, no e2e)
mkdir -p gpurun_out
TAG=${1:-diet}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_${TAG}_gputest.log
tail -4 gpurun_out/r02_${TAG}_gputest.log
for wl in c3 c1 c2 c4 c5; do
  steps=10; [ $wl = c5 ] && steps=3
  timeout 600 python bench.py --workload $wl --steps $steps --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02_${TAG}_$wl.json 2> gpurun_out/r02_${TAG}_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_${TAG}_$wl.json").read().strip().splitlines()[-1])
    print("$wl", round(d["ms_per_step"],4), round(d.get("ms_per_step_unprofiled") or 0,4), d.get("gpu_launches"), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.03}, (d.get("parity_check") or {}).get("depth"))
except Exception as e:
    print("$wl failed", e)
PY
done
