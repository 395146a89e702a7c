N=$1
mkdir -p gpurun_out
nproc; free -g | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err
tail -1 gpurun_out/scale_c3_n$N.json | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d["e2e"]; print("c3 N=%d"%d["n_gpus"], round(d["ms_per_step"],2), round(d["ms_per_step_unprofiled"],2), "Gtri/s", round(d["value"]/1e9,2), "e2e", round(e["ms_per_step"],2), e["host_ms_per_step"], "resident", round(e["scene_resident"]["ms_per_step"],2), "depth", round(e["with_depth_readback"]["ms_per_step"],2))'
