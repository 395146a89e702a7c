mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get("e2e"); print(sys.argv[1], round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), "Mtri/s", round(d["value"]/1e6,2), "e2e", e and round(e["ms_per_step"],3), "cpu", d.get("cpu_baseline") and d["cpu_baseline"]["value"], {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.02})'
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 2>gpurun_out/bench_c2.err | tee gpurun_out/bench_c2.json | python -c "$show" "c2"
timeout 600 python bench.py --workload c1 --steps 20 --warmup 3 2>gpurun_out/bench_c1.err | tee gpurun_out/bench_c1_v3.json | python -c "$show" "c1"
timeout 600 python bench.py --tga 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_v5.json | python -c "$show" "c3 default"
tail -2 gpurun_out/bench_c2.err
