# shade-pass trims: lit parity subset first, then the config-3 step (and config 2 / 4 as a regression check)
mkdir -p gpurun_out
TAG=${1:-r02c_shade}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lighting or exact_shading or fullsize or full_size or golden or head or orbit or shadow or gouraud" > gpurun_out/${TAG}_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_test.log
tail -5 gpurun_out/${TAG}_test.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_c3.json").read().strip().splitlines()[-1])
    print("c3", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), d["gpu_launches"], {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.015}, d["parity_check"])
except Exception as e:
    print("c3 failed", e)
PY
