# flushes inside the snapshot window collect their pixel list from the saved tiles: parity cases, then the config-3 step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu --maxfail=8 -q -k "snapshot or orbit or record or replay or c3_bench or host_example" > gpurun_out/r02d_snap3_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_snap3_test.log
tail -6 gpurun_out/r02d_snap3_test.log
printf 'snap3_lazy\n' | bash profiles/scripts/r02c_ab.sh
