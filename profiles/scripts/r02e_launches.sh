# ncu launch list of the final build (per-launch durations, cold cache / serialised: shares, not absolutes)
mkdir -p gpurun_out
timeout 25 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02e_plain_small.json 2>&1 || exit 1
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02e_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-also --no-exact-shade > gpurun_out/r02e_ncu_launches.log 2>&1
echo "ncu rc=$?"
