mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get("e2e"); print(sys.argv[1], round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), "Mtri/s", round(d["value"]/1e6,2), "e2e", e and round(e["ms_per_step"],3), "roof", d["roofline"]["frac"], d["step_roofline"]["frac"])'
timeout 600 python bench.py --tga 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_v6.json | python -c "$show" "c3 default"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_c4_full_v4.json | python -c "$show" "c4 full"
timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tee gpurun_out/bench_c5_full_v4.json | python -c "$show" "c5 full"
