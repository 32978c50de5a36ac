timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_persist.py tests/test_gpu_linear.py -x -q -m gpu 2>&1 | tail -8
