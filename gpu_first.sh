python -m pytest tests/test_gpu_linear.py -m gpu -x -q 2>&1 | tail -3
for a in "" "scan.stages=3" "scan.stages=8" "scan.ctas_per_sm=1 scan.stages=8" "scan.ctas_per_sm=3" "scan.ctas_per_sm=4" "scan.interleave=0" "scan.ctas_per_sm=3 scan.interleave=0"; do
python tools/scan_probe.py linear 1000000000 1 $a
done
python tools/scan_probe.py linear 1000000000 2
python tools/scan_probe.py linear 1000000000 4
python tools/scan_probe.py linear 1000000000 8
python tools/scan_probe.py linear 100000000 256
python tools/scan_probe.py linear 100000000 256 scan.prefilter=0
python tools/scan_probe.py linear 1000000000 1024
python tools/scan_probe.py linear 1000000000 1024 scan.waves=8
