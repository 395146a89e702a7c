# N GPUs (arg 1): config 4 sharded with the fused NVLink composite only
mkdir -p gpurun_out
N=${1:-8}
for envs in "X=1" "TRB_SHARE_VERTEX=0"; do
timeout 300 env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c4 --composite p2p --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_mg${N}_c4_p2p.json 2> gpurun_out/r02c_mg${N}_c4_p2p.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_mg${N}_c4_p2p.json").read().strip().splitlines()[-1])
    print("c4 p2p $envs", d["value"], d["ms_per_step"], d["ms_per_step_unprofiled"], d["parity_check"]["depth"], {kk:round(vv["ms"]/d["steps"],3) for kk,vv in d["kernels"].items() if vv["ms"]/d["steps"]>0.01})
except Exception as e:
    print("failed", e)
PY
done
