mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_example.py -m gpu -x -q 2>&1 | tail -5
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print(sys.argv[1], round(d["ms_per_step_unprofiled"],2), "e2e", round(e["ms_per_step"],2), e["host_ms_per_step"], "depth", round(e["with_depth_readback"]["ms_per_step"],2))'
for v in "" noupload noreadback; do
TRB_E2E_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c3.err | python -c "$show" "c3 variant=[$v]"
done
