mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_raster","k_raster_warp","k_shade","k_shade_dense","k_shade_collect","k_setup_count","k_direct_resolve","k_fill","k_vertex_mesh")})'
for wm in 0 1024 1000000000; do
TRB_WARP_MAX=$wm timeout 100 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c3 wm=$wm"
done
for wm in 0 256 1024 8192; do
TRB_WARP_MAX=$wm timeout 100 python bench.py --workload c5 --c5-tris 20000000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c5-20M wm=$wm"
TRB_WARP_MAX=$wm timeout 100 python bench.py --workload c4 --c4-level 9 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c4-l9 wm=$wm"
done
TRB_WARP_MAX=1024 TRB_DIRECT_AREA=0 timeout 100 python bench.py --workload c5 --c5-tris 20000000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c5-20M wm=1024 nodirect"
