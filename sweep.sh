mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), round(d["ms_per_step_unprofiled"],2), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.02})'
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_c3.err | python -c "$show" "c3"
timeout 100 python bench.py --workload c5 --c5-tris 20000000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c5-20M"
timeout 100 python bench.py --workload c4 --c4-level 9 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c4-l9"
