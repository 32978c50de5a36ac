for args in "" "mih.prefilter=1"; do
  timeout 300 python tools/scan_probe.py mih 1000000000 4096 $args 2>&1 | tail -1
done
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
