# ordered soups: parity subset (processing order forced onto every mesh, split bins, full-size config 5), then the config-5
# and config-4 lines and a config-3 check
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "processing_order or split_bins or soup or golden or indexing_invariants or both_raster" > gpurun_out/r02c_soup_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_soup_test.log
tail -6 gpurun_out/r02c_soup_test.log
for wlk in c5 c4; do
timeout 600 python bench.py --workload $wlk --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02c_soup_$wlk.json 2> gpurun_out/r02c_soup_$wlk.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_soup_$wlk.json").read().strip().splitlines()[-1])
    print("$wlk", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.02}, d["parity_check"], d["roofline"]["frac"])
except Exception as e:
    print("$wlk failed", e)
PY
done
bash profiles/scripts/r02c_ab.sh < profiles/scripts/ab_in.txt
