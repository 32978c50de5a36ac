#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
T=r02b
for lib in tools/bin/ab_base.so verticut_b200/lib/libverticut_gpu.so; do
  echo "== $lib"
  VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py mih 1000000000 16384 check=4 2>&1 | tail -1
  VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py mih 125000000 16384 shards=8 2>&1 | tail -1
  VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py mih 125000000 4096 shards=8 2>&1 | tail -1
  VC_GPU_LIB=$PWD/$lib timeout 100 python tools/merge_probe.py 2>&1 | tail -1
  VC_GPU_LIB=$PWD/$lib timeout 100 python tools/merge_probe.py lists=2 nq=4096 2>&1 | tail -1
done > gpurun_out/${T}_ab1.log 2>&1
cat gpurun_out/${T}_ab1.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest1.log 2>&1; tail -n 5 gpurun_out/${T}_pytest1.log
