# ncu capture of the c3 hot kernels (config 3, 8 frames per step).  Matching launches per step:
# setup(room) raster(room) setup(head) raster(head) shade(full) setup(eyes) raster(eyes) shade(eyes);
# 8 in the diagnostics step + 7 in the visible-triangle pass are skipped, the warm-up step is captured.
mkdir -p gpurun_out
TAG=${1:-r02_c3}
CMD="python bench.py --steps 2 --warmup 1 --frames-per-step 8 --no-e2e --no-cpu-baseline --no-also --no-exact-shade"
timeout 200 $CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_raster_warp|k_shade_dense|k_shade_rec|k_setup_count' -s 15 -c 5 -o gpurun_out/$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
