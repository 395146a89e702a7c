mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get("e2e"); print(sys.argv[1], round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), "Mtri/s", round(d["value"]/1e6,2), "e2e", e and round(e["ms_per_step"],3), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_raster_warp","k_shade_dense")})'
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench_c3.err | tee gpurun_out/bench_c3_v10.json | python -c "$show" "c3 default"
