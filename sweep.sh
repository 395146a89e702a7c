mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],2), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_raster","k_shade","k_setup_count","k_fill","k_vertex_mesh","k_clear")})'
for cfg in "16 24" "16 1" "8 24" "12 24" "16 64" "32 24" "9999 1"; do set -- $cfg; TRB_BIG_NS=$1 TRB_SMALL_MIN=$2 timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c3 big_ns=$1 small_min=$2"; done
for cfg in "16 24" "8 24" "16 1" "10 24" ; do set -- $cfg; TRB_BIG_NS=$1 TRB_SMALL_MIN=$2 timeout 300 python bench.py --workload c5 --c5-tris 20000000 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c5-20M big_ns=$1 small_min=$2"; done
for cfg in "16 24" "8 24"; do set -- $cfg; TRB_BIG_NS=$1 TRB_SMALL_MIN=$2 timeout 300 python bench.py --workload c4 --c4-level 9 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$show" "c4-l9 big_ns=$1 small_min=$2"; done
