# one GPU: the multi-context tests (composite group with touched maps, trb_draw_shard), the C++ host example tests, config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_host_example.py -m gpu -x -q > gpurun_out/r02b_shard1_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_shard1_test.log
tail -15 gpurun_out/r02b_shard1_test.log
