# last evidence of round 2 on one GPU (the budget left is ~5 minutes): smoke, the default bench line, then the full parity suite
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python bench.py --steps 20 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo "bench rc=$?"
timeout 215 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_final_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_final_gputest.log
tail -4 gpurun_out/r02e_final_gputest.log
