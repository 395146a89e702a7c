mkdir -p gpurun_out
CMD="python bench.py --workload c5 --c5-tris 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_setup_count|k_direct_resolve|k_vertex_mesh' -s 6 -c 3 -o gpurun_out/r01_c5_setup $CMD > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
CMD="python bench.py --workload c4 --c4-level 9 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:'k_setup_count|k_raster' -s 6 -c 3 -o gpurun_out/r01_c4_setup $CMD > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
