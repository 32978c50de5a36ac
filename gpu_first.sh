timeout 900 python -m pytest tests/test_gpu_host_cli.py tests/test_gpu_persist.py -m gpu -x -q 2>&1 | tail -5
