#!/bin/bash
# round 2, GPU call 9 (one GPU): two-phase first step (own buckets on the exact distance, then the rest on the lower bound)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bmih.py tests/test_gpu_mih.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest9.log 2>&1; tail -3 gpurun_out/r02_pytest9.log
{
echo "== headline batch 16384: auto / off"; python tools/probe.py mih 1000000000 16384 reps=2 check=4 | tail -1; python tools/probe.py mih 1000000000 16384 reps=2 mih.r0_first=0 | tail -1
echo "== headline batch 32768: auto / off"; python tools/probe.py mih 1000000000 32768 reps=2 | tail -1; python tools/probe.py mih 1000000000 32768 reps=2 mih.r0_first=0 | tail -1
echo "== headline batch 8192: auto / off"; python tools/probe.py mih 1000000000 8192 reps=2 | tail -1; python tools/probe.py mih 1000000000 8192 reps=2 mih.r0_first=0 | tail -1
echo "== headline batch 4096: auto / forced on"; python tools/probe.py mih 1000000000 4096 reps=3 | tail -1; python tools/probe.py mih 1000000000 4096 reps=3 mih.r0_first=1 | tail -1
echo "== C3 shard batch 4096: auto / off"; python tools/probe.py mih 125000000 4096 bits=128 m=8 check=4 | tail -1; python tools/probe.py mih 125000000 4096 bits=128 m=8 mih.r0_first=0 | tail -1
echo "== C2: auto / off"; python tools/probe.py mih 100000000 16384 reps=3 | tail -1; python tools/probe.py mih 100000000 16384 reps=3 mih.r0_first=0 | tail -1
} > gpurun_out/r02_ab9.log 2>&1
grep -c kernel_ms gpurun_out/r02_ab9.log
