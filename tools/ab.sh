#!/bin/bash
# A/B of builds of the C-ABI library on a GPU box: every tools/bin/ab_*.so and the in-tree library, each in its own process,
# on the bench workload (or the scan_probe.py arguments given):   gpurun -- 'bash tools/ab.sh [mih 1000000000 4096 [knob=value ...]]'
# About 5 s per line on a B200 (the 1 B-code index builds in under a second).  Log of the round-1 runs: profiles/ab_r01.md.
args=${@:-mih 1000000000 4096}
for lib in verticut_b200/lib/libverticut_gpu.so tools/bin/ab_*.so verticut_b200/lib/libverticut_gpu.so; do
  [ -f "$lib" ] || continue
  echo "$lib"; VC_GPU_LIB=$PWD/$lib timeout 300 python tools/scan_probe.py $args 2>&1 | tail -1
done
