# round 2, first GPU visit: parity suite with the rewritten warp raster kernel, then quick c3 benches for the three
# register variants of k_raster_warp, then the default bench line (with "also" c4 / c5 and parity_check)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest.log
tail -5 gpurun_out/r02_gputest.log
for mb in 8 7 6; do
  TRB_RW_BLOCKS=$mb timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02_c3_mb$mb.json 2> gpurun_out/r02_c3_mb$mb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_c3_mb$mb.json"))
print("mb$mb", d["ms_per_step"], d["ms_per_step_unprofiled"], {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.05}, d["parity_check"])
PY
done
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_default.err
