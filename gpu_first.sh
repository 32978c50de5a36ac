timeout 1200 python -m pytest tests/test_gpu_bmih.py tests/test_gpu_sharded.py tests/test_gpu_mih.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
python tools/scan_probe.py mih 1000000000 4096
python tools/scan_probe.py mih 125000000 4096
python tools/latency_probe.py 1000000000 1,4096,16384
