# A/B of builds of the C-ABI library / knob sweep on the bench workload
for args in "" "mih.boot_sample=32768" "mih.boot_sample=65536" ""; do
  timeout 300 python tools/scan_probe.py mih 1000000000 4096 $args 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_bmih.py -x -q -m gpu 2>&1 | tail -3
