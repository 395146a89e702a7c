# 8 GPUs: the default bench under torchrun exactly as the driver launches it (c3 frames sharded incl. e2e, "also" c4 sharded)
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02c_mg${N}_bench.json 2> gpurun_out/r02c_mg${N}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02c_mg${N}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02c_mg${N}_bench.json").read().strip().splitlines()[-1])
e=d.get("e2e") or {}
print("c3", d["value"], d["ms_per_step"], d["parity_check"]["depth"], "e2e", e.get("value"), e.get("ms_per_step"), "tga", (e.get("tga_files") or {}).get("ms_per_step"), (e.get("tga_files") or {}).get("d2h_bytes_per_step"), "resident", (e.get("scene_resident") or {}).get("ms_per_step"))
for k,v in d.get("also",{}).items():
    if isinstance(v,dict): print(k, v["value"], v["ms_per_step"], v["ms_per_step_unprofiled"], v["parity_check"]["depth"], {kk:round(vv["ms"]/v["steps"],3) for kk,vv in v["kernels"].items() if vv["ms"]/v["steps"]>0.02})
    else: print(k, v)
PY
