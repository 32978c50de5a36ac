timeout 120 python tools/scan_probe.py scan 1000000000 64 scan.tc=1 tc.trace=1 > gpurun_out/tc_trace_full.log 2>&1
