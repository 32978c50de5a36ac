#!/bin/bash
# round 2, GPU call 6 (one GPU): first search step at batch 16384 - exact vs lower-bound filter, radius-0 buckets first, radius 0 as its own step
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bmih.py tests/test_gpu_mih.py tests/test_gpu_sharded.py tests/test_gpu_persist.py -m gpu -x -q > gpurun_out/r02_pytest6.log 2>&1; tail -3 gpurun_out/r02_pytest6.log
{
P="python tools/probe.py mih 1000000000 16384 reps=2 check=4"
echo "== headline batch 16384: first-step variants"
echo "-- auto (exact filter in the first step)"; $P | tail -1
echo "-- lower-bound filter everywhere (mih.prefilter=1)"; $P mih.prefilter=1 | tail -1
echo "-- radius-0 buckets first, exact"; $P mih.r0_first=1 | tail -1
echo "-- radius-0 buckets first, lower bound"; $P mih.r0_first=1 mih.prefilter=1 | tail -1
echo "-- radius 0 as its own step"; $P mih.split_r0=1 | tail -1
echo "-- radius 0 as its own step, lower bound everywhere"; $P mih.split_r0=1 mih.prefilter=1 | tail -1
echo "-- bootstrap sample 65536"; $P mih.boot_sample=65536 | tail -1
echo "-- bootstrap sample 65536, radius-0 first"; $P mih.boot_sample=65536 mih.r0_first=1 | tail -1
echo "== batch 4096"
P4="python tools/probe.py mih 1000000000 4096 reps=3"
$P4 | tail -1; $P4 mih.r0_first=1 | tail -1; $P4 mih.split_r0=1 | tail -1
echo "== C3 shard batch 4096: in-tree vs 4 CTAs per SM"; bash tools/ab.sh mih 125000000 4096 bits=128 m=8
echo "== C3 shard first-step variants"; python tools/probe.py mih 125000000 4096 bits=128 m=8 mih.r0_first=1 | tail -1; python tools/probe.py mih 125000000 4096 bits=128 m=8 mih.split_r0=1 | tail -1
echo "== C5 r=3 batch 1024: in-tree vs 4 CTAs per SM"; bash tools/ab.sh mih 60000000 1024 bits=256 m=16 k=1000 r=3
} > gpurun_out/r02_ab6.log 2>&1
grep -c kernel_ms gpurun_out/r02_ab6.log
