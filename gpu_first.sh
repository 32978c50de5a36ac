python -m pytest tests/test_gpu_mih.py -m gpu -x -q 2>&1 | tail -40
