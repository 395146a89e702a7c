# tile-granular snapshot, second version of the copy kernels (flagged tiles dealt out to the CTA's warps, interior tiles with
# all loads in flight): snapshot / orbit / replay parity cases, then the config-3 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu --maxfail=8 -q -k "snapshot or orbit or record or replay or c3_bench or host_example" > gpurun_out/r02d_snap2_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_snap2_test.log
tail -6 gpurun_out/r02d_snap2_test.log
printf 'snap2_lazy\n' | bash profiles/scripts/r02c_ab.sh
