# config 3 / 4 / 5 lines (no tests)
mkdir -p gpurun_out
bash profiles/scripts/r02c_ab.sh < profiles/scripts/ab_in.txt
for wlk in c4 c5; do
timeout 600 python bench.py --workload $wlk --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02c_now_$wlk.json 2> gpurun_out/r02c_now_$wlk.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_now_$wlk.json").read().strip().splitlines()[-1])
    print("$wlk", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.02}, d["parity_check"].get("depth"), d["parity_check"].get("colour"), round(d["roofline"]["frac"],3))
except Exception as e:
    print("$wlk failed", e)
PY
done
