timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/scan_probe.py mih 1000000000 4096
python tools/scan_probe.py mih 1000000000 4096 mih.table_steps=0
python tools/scan_probe.py mih 125000000 4096
python tools/latency_probe.py 1000000000 1,4096
