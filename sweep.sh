mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), round(d["ms_per_step_unprofiled"],2), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_raster","k_raster_warp","k_shade","k_shade_dense","k_shade_collect","k_setup_count","k_direct_resolve","k_fill","k_vertex_mesh","k_unbinned_depth","k_unbinned_ids")}, d.get("e2e"))'
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_async2.json | python -c "$show" "c3 async"
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c3.err | python -c "$show" "c3 async 20 steps"
