mkdir -p gpurun_out
( time python bench.py --tga > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
tail -3 gpurun_out/bench_default.time
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2> gpurun_out/bench_reference.time
tail -3 gpurun_out/bench_reference.time
cat gpurun_out/bench_reference.json
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -2
