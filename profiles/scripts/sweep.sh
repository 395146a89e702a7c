show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if k in ("k_raster_warp","k_raster")})'
for v in default rw9 default rw9; do
if [ $v = default ]; then L=tinyrenderder_b200/libtrb.so; else L=build/libtrb_$v.so; fi
TRB_CUDA_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$show" "c3 $v"
done
for v in default rw9; do
if [ $v = default ]; then L=tinyrenderder_b200/libtrb.so; else L=build/libtrb_$v.so; fi
TRB_CUDA_LIB=$L timeout 300 python bench.py --workload c4 --c4-level 9 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "$show" "c4-l9 $v"
done
