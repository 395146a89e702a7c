# usage: bash scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err
tail -1 gpurun_out/scale_c3_n$N.json | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("c3 N=%d"%d["n_gpus"], round(d["ms_per_step"],2), round(d["ms_per_step_unprofiled"],2), "Gtri/s", round(d["value"]/1e9,2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), d["clocks"])'
timeout 400 $TR bench.py --gpus $N --workload c4 --composite p2p --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/scale_c4_n${N}_p2p.json 2> gpurun_out/scale_c4_n${N}_p2p.err
tail -1 gpurun_out/scale_c4_n${N}_p2p.json | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("c4 p2p N=%d"%d["n_gpus"], round(d["ms_per_step"],2), "Gtri/s", round(d["value"]/1e9,2))'
if [ "$N" = "2" ]; then
timeout 400 $TR bench.py --gpus $N --workload c4 --composite nccl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/scale_c4_n${N}_nccl.json 2> gpurun_out/scale_c4_n${N}_nccl.err
tail -1 gpurun_out/scale_c4_n${N}_nccl.json | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("c4 nccl N=%d"%d["n_gpus"], round(d["ms_per_step"],2), "Gtri/s", round(d["value"]/1e9,2))'
fi
tail -3 gpurun_out/scale_c3_n$N.err gpurun_out/scale_c4_n${N}_p2p.err | cut -c1-300
