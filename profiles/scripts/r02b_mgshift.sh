# N GPUs (arg 1): config 4 sharded (fused NVLink composite) for several block sizes of trb_draw_shard
mkdir -p gpurun_out
N=${1:-2}
for sh in 12 14 16 18; do
TRB_SHARD_SHIFT=$sh timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c4 --composite p2p --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_mg${N}_c4_sh$sh.json 2> gpurun_out/r02b_mg${N}_c4_sh$sh.err; echo "bench $sh rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02b_mg${N}_c4_sh$sh.json").read().strip().splitlines()[-1])
    print("c4 shift $sh", d["ms_per_step"], d["ms_per_step_unprofiled"], d["parity_check"]["depth"], {kk:round(vv["ms"]/d["steps"],3) for kk,vv in d["kernels"].items() if vv["ms"]/d["steps"]>0.01})
except Exception as e:
    print("failed", e)
PY
done
