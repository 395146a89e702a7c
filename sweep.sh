mkdir -p gpurun_out
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), round(d["ms_per_step_unprofiled"],2), "Mtri/s", round(d["value"]/1e6,1), d.get("e2e"), d.get("tga_encode"))'
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tga 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_async4.json | python -c "$show" "c3 async 10 steps"
tail -3 gpurun_out/bench_c3.err
