# ordered soups, direct-path vote by pixel: config 5 / config 4 A/B over the two switches, config 3 check
mkdir -p gpurun_out
printf 'ord_px X=1\nord_tile TRB_DIRECT_BY_PIXEL=0\nnoord_px TRB_MESH_ORDER_MIN_TRIS=0\nnoord_tile TRB_MESH_ORDER_MIN_TRIS=0 TRB_DIRECT_BY_PIXEL=0\n' | while read name envs; do
for wlk in c5 c4; do
[ "$wlk" = c4 ] && [ "$name" != ord_px ] && [ "$name" != ord_tile ] && continue
env $envs timeout 600 python bench.py --workload $wlk --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-exact-shade > gpurun_out/r02c_${name}_$wlk.json 2> gpurun_out/r02c_${name}_$wlk.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_${name}_$wlk.json").read().strip().splitlines()[-1])
    print("$name $wlk", round(d["ms_per_step"],3), round(d["ms_per_step_unprofiled"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]/d["steps"]>0.02}, d["parity_check"].get("depth"), d["parity_check"].get("colour"), round(d["roofline"]["frac"],3))
except Exception as e:
    print("$name $wlk failed", e)
PY
done
done
bash profiles/scripts/r02c_ab.sh < profiles/scripts/ab_in.txt
