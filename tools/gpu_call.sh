#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
T=r02b
{
  for lib in tools/bin/ab_base.so tools/bin/ab_fence.so tools/bin/ab_late.so tools/bin/ab_latefence.so; do
    echo "== $lib"
    VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py linear 1000000000 1 reps=10 2>&1 | tail -1
    VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py linear 1000000000 1 k=1000 reps=10 2>&1 | tail -1
    VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py linear 250000000 1 bits=256 reps=10 2>&1 | tail -1
    VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py linear 100000000 4 reps=10 2>&1 | tail -1
  done
  for lib in tools/bin/ab_base.so tools/bin/ab_fence.so; do
    echo "== tc $lib"
    VC_GPU_LIB=$PWD/$lib timeout 200 python tools/probe.py linear 1000000000 1024 reps=2 2>&1 | tail -1
  done
  for lib in tools/bin/ab_fence.so tools/bin/ab_latefence.so; do
    echo "== flake $lib"; VC_GPU_LIB=$PWD/$lib timeout 300 python tools/flake_probe.py reps=100 2>&1 | tail -2
    VC_GPU_LIB=$PWD/$lib timeout 300 python tools/flake_probe.py reps=50 scan.qt=2 2>&1 | tail -1
  done
} > gpurun_out/${T}_fence1.log 2>&1
cat gpurun_out/${T}_fence1.log
