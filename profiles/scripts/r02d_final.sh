# round 2 (last part) evidence on one GPU after the tile-granular snapshot: full parity suite, smoke, the default bench line,
# the reference arm, and the ncu launch list of the same command (per-launch durations, cold cache / serialised: shares)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_final_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_final_gputest.log
tail -4 gpurun_out/r02d_final_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_reference_arm.json 2> gpurun_out/r02d_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02d_plain_small.json 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02d_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02d_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
