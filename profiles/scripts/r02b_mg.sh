# N GPUs (arg 1): the IPC composite test (one process per GPU), then config 4 sharded with the fused NVLink composite and with NCCL
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "across_processes" > gpurun_out/r02b_mg${N}_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_mg${N}_test.log
tail -3 gpurun_out/r02b_mg${N}_test.log
for comp in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload c4 --composite $comp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_mg${N}_c4_$comp.json 2> gpurun_out/r02b_mg${N}_c4_$comp.err; echo "bench $comp rc=$?"
tail -c 300 gpurun_out/r02b_mg${N}_c4_$comp.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02b_mg${N}_c4_$comp.json").read().strip().splitlines()[-1])
    print("c4 $comp", d["value"], d["ms_per_step"], d["ms_per_step_unprofiled"], d["parity_check"], {kk:round(vv["ms"]/d["steps"],3) for kk,vv in d["kernels"].items() if vv["ms"]/d["steps"]>0.01})
except Exception as e:
    print("failed", e)
PY
done
