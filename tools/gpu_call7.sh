#!/bin/bash
# round 2, GPU call 7 (one GPU): per-record filter switch + 4 CTAs for 128/256-bit: parity, then the workloads
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bmih.py tests/test_gpu_mih.py tests/test_gpu_sharded.py tests/test_gpu_scan_batched.py -m gpu -x -q > gpurun_out/r02_pytest7b.log 2>&1; tail -3 gpurun_out/r02_pytest7b.log
{
echo "== headline batch 16384"; python tools/probe.py mih 1000000000 16384 reps=2 check=4 | tail -1


echo "== headline batch 4096"; python tools/probe.py mih 1000000000 4096 reps=3 | tail -1
echo "== C2"; python tools/probe.py mih 100000000 4096 reps=3 | tail -1
echo "== C3 shard batch 4096 / 1024"; python tools/probe.py mih 125000000 4096 bits=128 m=8 check=4 | tail -1; python tools/probe.py mih 125000000 1024 bits=128 m=8 | tail -1
echo "== C5 r=3 batch 1024 / 256"; python tools/probe.py mih 60000000 1024 bits=256 m=16 k=1000 r=3 | tail -1; python tools/probe.py mih 60000000 256 bits=256 m=16 k=1000 r=3 | tail -1
echo "== scan 1 B x 64-bit"; for b in 2 4 8 16 64; do python tools/probe.py linear 1000000000 $b | tail -1; done
} > gpurun_out/r02_ab7b.log 2>&1
grep -c kernel_ms gpurun_out/r02_ab7b.log
