mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), "Mtri/s", round(d["value"]/1e6,1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_setup_count","k_direct_resolve","k_fill")})'
for v in default c1 c2 c8 c4b3 c8b3; do
if [ $v = default ]; then L=tinyrenderder_b200/libtrb.so; else L=build/libtrb_$v.so; fi
TRB_CUDA_LIB=$L timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$show" "c4 full $v"
done
for v in default c1 c8b3; do
if [ $v = default ]; then L=tinyrenderder_b200/libtrb.so; else L=build/libtrb_$v.so; fi
TRB_CUDA_LIB=$L timeout 300 python bench.py --workload c5 --c5-tris 20000000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$show" "c5-20M $v"
TRB_CUDA_LIB=$L timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$show" "c3 $v"
done
