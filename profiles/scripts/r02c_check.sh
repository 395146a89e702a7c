# quick check after a shade / host change: lit + composite parity subset, the record/replay tests (Python and C++), config 3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_record_replay.py tests/test_host_example.py tests/test_multigpu.py -m gpu -x -q -k "lighting or exact_shading or golden or head or orbit or shadow or gouraud or record or replay or views or composite or contexts or clip" > gpurun_out/r02c_check_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_check_test.log
tail -12 gpurun_out/r02c_check_test.log
bash profiles/scripts/r02c_ab.sh < profiles/scripts/ab_in.txt
