timeout 600 python -m pytest tests/test_host_example.py -m gpu -x -q 2>&1 | tail -12
