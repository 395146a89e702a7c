mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_raster","k_raster_warp","k_shade","k_shade_dense","k_shade_collect","k_setup_count","k_direct_resolve","k_fill","k_vertex_mesh")}, d.get("e2e"))'
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_fast.json | python -c "$show" "c3 fast"
TRB_SHADE_EXACT=1 timeout 100 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c3 exact"
timeout 100 python bench.py --workload c1 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c1 fast"
