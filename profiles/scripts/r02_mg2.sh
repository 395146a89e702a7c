# 2 GPUs: the IPC composite test (one process per GPU), then the default bench under torchrun (c3 frames sharded +
# "also" config 4 sharded with the fused P2P composite and with NCCL)
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "across_processes" > gpurun_out/r02_mg${N}_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_mg${N}_test.log
tail -5 gpurun_out/r02_mg${N}_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 ${2:-} > gpurun_out/r02_mg${N}_bench.json 2> gpurun_out/r02_mg${N}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_mg${N}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_mg${N}_bench.json").read().strip().splitlines()[-1])
print("c3", d["value"], d["ms_per_step"], d["parity_check"], "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("ms_per_step"))
for k,v in d.get("also",{}).items():
    if isinstance(v,dict): print(k, v["value"], v["ms_per_step"], v["ms_per_step_unprofiled"], v["parity_check"], {kk:round(vv["ms"]/v["steps"],3) for kk,vv in v["kernels"].items() if vv["ms"]/v["steps"]>0.02})
    else: print(k, v)
PY
