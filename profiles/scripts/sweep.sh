mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get("e2e"); print(sys.argv[1], round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), "Mtri/s", round(d["value"]/1e6,2), "e2e", e and round(e["ms_per_step"],3), e["host_ms_per_step"], "depth", round(e["with_depth_readback"]["ms_per_step"],2), "resident", round(e["scene_resident"]["ms_per_step"],2))'
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_v8.json | python -c "$show" "c3 default run $i"
done
