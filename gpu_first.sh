set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests/test_gpu_linear.py -m gpu -x -q 2>&1 | tail -30
