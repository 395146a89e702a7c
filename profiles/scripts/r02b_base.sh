# re-entry baseline: full GPU parity suite, default bench line, ncu --set full of the config-4 / config-5 set-up kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_gputest.log
tail -4 gpurun_out/r02b_gputest.log
timeout 900 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
CMD4="python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-exact-shade"
timeout 300 $CMD4 > gpurun_out/r02b_plain_c4.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_setup_count|k_vertex_mesh|k_fill' -s 3 -c 3 -o gpurun_out/r02b_c4_setup $CMD4 > gpurun_out/ncu_r02b_c4.log 2>&1
CMD5="python bench.py --workload c5 --c5-tris 20000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-exact-shade"
timeout 300 $CMD5 > gpurun_out/r02b_plain_c5.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_setup_count|k_direct_resolve|k_vertex' -s 3 -c 3 -o gpurun_out/r02b_c5_setup $CMD5 > gpurun_out/ncu_r02b_c5.log 2>&1
tail -2 gpurun_out/ncu_r02b_c4.log gpurun_out/ncu_r02b_c5.log
