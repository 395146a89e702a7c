mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --frames-per-step 8 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_raster' -s 9 -c 1 -o gpurun_out/r01_c3_raster_v3 $CMD > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
