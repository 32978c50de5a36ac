# A/B of builds of the C-ABI library: batched MIH and batched scan
for lib in tools/bin/ab_c3.so verticut_b200/lib/libverticut_gpu.so; do
  echo $lib
  VC_GPU_LIB=$PWD/$lib timeout 300 python tools/scan_probe.py mih 1000000000 4096 2>&1 | tail -1
  VC_GPU_LIB=$PWD/$lib timeout 300 python tools/scan_probe.py mih 100000000 4096 2>&1 | tail -1
  VC_GPU_LIB=$PWD/$lib timeout 300 python tools/scan_probe.py linear 1000000000 64 2>&1 | tail -1
  VC_GPU_LIB=$PWD/$lib timeout 300 python tools/scan_probe.py linear 1000000000 1024 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_bmih.py tests/test_gpu_scan_batched.py tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
