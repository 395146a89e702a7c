# tile-granular depth snapshot: the parity cases that use snapshot / restore (every variant), then A/B on the config-3 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu --maxfail=8 -q -k "snapshot or orbit or record or replay or c3_bench or host_example" > gpurun_out/r02d_snap_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_snap_test.log
tail -6 gpurun_out/r02d_snap_test.log
printf 'snap_lazy\nsnap_eager TRB_LAZY_SNAPSHOT=0\n' | bash profiles/scripts/r02c_ab.sh
