python -m pytest tests/test_gpu_mih.py -m gpu -x -q 2>&1 | tail -3
python tools/scan_probe.py mih 100000000 1024
python tools/scan_probe.py mih 1000000000 1024
python tools/scan_probe.py mih 1000000000 256 && ncu --set full --clock-control none --import-source on -k regex:mih_search -s 2 -c 1 -o gpurun_out/mih_b256 python tools/scan_probe.py mih 1000000000 256 > gpurun_out/ncu_mih.log 2>&1
