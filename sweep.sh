mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), round(d["ms_per_step_unprofiled"],2), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.02})'
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_c3.err | python -c "$show" "c3"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_c4.err | tee gpurun_out/bench_c4_full_v3.json | python -c "$show" "c4 full"
timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_c5.err | tee gpurun_out/bench_c5_full_v3.json | python -c "$show" "c5 full"
