# round 2 final evidence on one GPU: full parity suite, smoke, the default bench line, the ncu launch list of the same
# command (per-launch durations, cold cache / serialised: shares, not absolutes) and one --set full capture of the hot kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_gputest.log
tail -4 gpurun_out/r02_final_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02_plain_small.json 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02_ncu_launches.log 2>&1
bash profiles/scripts/r02_prof.sh r02_c3_final
