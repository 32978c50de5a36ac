python tools/scan_probe.py mih 1000000000 4096
python tools/scan_probe.py mih 125000000 4096
