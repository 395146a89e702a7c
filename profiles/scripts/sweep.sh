python profiles/trace_e2e.py 6 e2e 2>&1 | grep -v "^   " | tail -12
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print(sys.argv[1], round(d["ms_per_step_unprofiled"],2), "e2e", round(e["ms_per_step"],2), e["host_ms_per_step"], "depth", round(e["with_depth_readback"]["ms_per_step"],2), "resident", round(e["scene_resident"]["ms_per_step"],2))'
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c3.err | python -c "$show" "c3"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
