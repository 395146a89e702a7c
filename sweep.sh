show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_raster","k_shade","k_setup_count","k_fill","k_vertex_mesh","k_clear")})'
for mb in 4 3; do cp gpurun_out_libtrb_mb$mb.so tinyrenderder_b200/libtrb.so
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c3 mb=$mb"
timeout 300 python bench.py --workload c5 --c5-tris 20000000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c5-20M mb=$mb"
done
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
