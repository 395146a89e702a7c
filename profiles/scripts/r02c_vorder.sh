# vertex renumbering: parity subset with the order forced onto every mesh, in-process composite groups, config 4 A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q -k "processing_order or split_bins or golden or shard or composite or indexing_invariants" > gpurun_out/r02c_vorder_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_vorder_test.log
tail -5 gpurun_out/r02c_vorder_test.log
printf 'vord X=1\nnovord TRB_MESH_ORDER_VERTICES=0\n' | while read name envs; do
env $envs timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02c_${name}_c4.json 2> gpurun_out/r02c_${name}_c4.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_${name}_c4.json").read().strip().splitlines()[-1])
    print("$name c4", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.02}, d["parity_check"].get("depth"), d["parity_check"].get("colour"), round(d["roofline"]["frac"],3))
except Exception as e:
    print("$name c4 failed", e)
PY
done
