mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --frames-per-step 8 --no-e2e --no-cpu-baseline"
timeout 200 $CMD > gpurun_out/plain2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c3_v3.csv $CMD > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_raster_warp|k_shade_dense|k_setup_count' -s 10 -c 4 -o gpurun_out/r01_c3_v3 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
