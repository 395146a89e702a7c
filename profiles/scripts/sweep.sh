mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "shadow or gouraud or smoke or abi" 2>&1 | tail -3
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d.get("e2e"); print(sys.argv[1], round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), "Mtri/s", round(d["value"]/1e6,2), "e2e", e and round(e["ms_per_step"],3), e and e["host_ms_per_step"], "resident", e and round(e["scene_resident"]["ms_per_step"],3))'
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_c2.err | tee gpurun_out/bench_c2.json | python -c "$show" "c2"
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -1
