# span-clipped warp raster: parity first (normal + self-checking build on the span-sensitive cases), then A/B benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not c5" > gpurun_out/r02_span_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_span_test.log
tail -6 gpurun_out/r02_span_test.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02_$name.json 2> gpurun_out/r02_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_$name.json"))
    print("$name", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.05}, d["parity_check"]["depth"])
except Exception as e:
    print("$name failed", e)
PY
}
run w1_mb6 TRB_RW_BLOCKS=6
run w1_mb7 TRB_RW_BLOCKS=7
run w1_mb8 TRB_RW_BLOCKS=8
run w4_mb6 TRB_RW_BLOCKS=6 TRB_CUDA_LIB=build/variants/libtrb_w4.so
run w4_mb8 TRB_RW_BLOCKS=8 TRB_CUDA_LIB=build/variants/libtrb_w4.so
