# mesh processing order: config 4 with and without it, then ncu of the ordered set-up (gpurun forwards no stdin: the
# variants are listed here)
mkdir -p gpurun_out
printf 'order\nnoorder TRB_MESH_ORDER_MIN_TRIS=0\n' | while read name envs; do
  [ -z "$name" ] && continue
  env $envs timeout 400 python bench.py --workload c4 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02b_${name}_c4.json 2> gpurun_out/r02b_${name}_c4.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02b_${name}_c4.json"))
    print("$name c4", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.02}, d["parity_check"].get("depth"), d["parity_check"].get("colour"))
except Exception as e:
    print("$name c4 failed", e)
PY
done
CMD4="python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-exact-shade"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_setup_count|k_fill|k_raster' -s 4 -c 4 -o gpurun_out/r02b_c4_order $CMD4 > gpurun_out/ncu_r02b_c4_order.log 2>&1
tail -2 gpurun_out/ncu_r02b_c4_order.log
