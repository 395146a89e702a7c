# frame recordings: the new tests, the shade parity subset, then plain-vs-replay timing of configs 1 and 2, then config 3
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_record_replay.py -m gpu -x -q > gpurun_out/r02c_replay_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_replay_test.log
tail -25 gpurun_out/r02c_replay_test.log
timeout 300 python profiles/experiments/replay_c1c2.py 200 > gpurun_out/r02c_replay_c1c2.json 2> gpurun_out/r02c_replay_c1c2.err; echo "replay bench rc=$?"; tail -3 gpurun_out/r02c_replay_c1c2.err; cat gpurun_out/r02c_replay_c1c2.json
bash profiles/scripts/r02c_ab.sh < profiles/scripts/ab_in.txt
