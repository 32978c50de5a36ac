#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
T=r02b
{
  for sp in 0 1; do
    timeout 200 python tools/probe.py mih 1000000000 16384 check=16 reps=4 mih.speculate=$sp 2>&1 | tail -1
    timeout 200 python tools/probe.py mih 1000000000 4096 reps=4 mih.speculate=$sp 2>&1 | tail -1
    timeout 200 python tools/probe.py mih 125000000 16384 shards=8 reps=4 mih.speculate=$sp 2>&1 | tail -1
    timeout 200 python tools/probe.py mih 100000000 4096 reps=4 mih.speculate=$sp 2>&1 | tail -1
    timeout 200 python tools/probe.py mih 125000000 4096 bits=128 m=8 reps=3 mih.speculate=$sp 2>&1 | tail -1
  done
} > gpurun_out/${T}_spec1.log 2>&1
cut -c1-700 gpurun_out/${T}_spec1.log
timeout 600 python -m pytest tests/test_gpu_bmih.py tests/test_gpu_sharded.py tests/test_gpu_mih.py -x -q > gpurun_out/${T}_pytest4.log 2>&1; tail -n 15 gpurun_out/${T}_pytest4.log
