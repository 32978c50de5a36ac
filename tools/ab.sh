#!/bin/bash
# A/B of builds of the C-ABI library on a GPU box: the in-tree library and every tools/bin/ab_*.so, each in its own process,
# on the tools/probe.py workload given (default: the bench workload):   gpurun -- 'bash tools/ab.sh mih 1000000000 4096 [name=value ...]'
# About 5 s per line on a B200 (the 1 B-code index builds in under a second).  Logs: profiles/ab_r01.md, profiles/ab_r02.md.
args=${@:-mih 1000000000 4096}
for lib in verticut_b200/lib/libverticut_gpu.so tools/bin/ab_*.so; do
  [ -f "$lib" ] || continue
  echo "$lib"; VC_GPU_LIB=$PWD/$lib timeout 300 python tools/probe.py $args 2>&1 | tail -1
done
