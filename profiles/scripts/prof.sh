mkdir -p gpurun_out
CMD="python bench.py --workload c4 --c4-level 9 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_setup_count' -s 2 -c 1 -o gpurun_out/r01_c4_setup_v2 $CMD > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
