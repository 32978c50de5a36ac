#!/bin/bash
# round 2, GPU call 1: parity suite, then A/B of the verify-kernel variants on the three shapes, then ncu captures (W = 2, W = 4)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; tail -5 gpurun_out/r02_pytest1.log
{
echo "== headline: 1 B x 64-bit, m=4, k=100, batch 4096"; bash tools/ab.sh mih 1000000000 4096 check=4
echo "== C3 shard: 125 M x 128-bit, m=8, k=100, batch 1024"; bash tools/ab.sh mih 125000000 1024 bits=128 m=8 check=4
echo "== C5: 60 M x 256-bit, m=16, k=1000, batch 256, r=1..3"
for r in 1 2 3; do bash tools/ab.sh mih 60000000 256 bits=256 m=16 k=1000 r=$r; done
echo "== C5 batch 1024 r=3 (in-tree only)"; python tools/probe.py mih 60000000 1024 bits=256 m=16 k=1000 r=3 | tail -1
echo "== C2: 100 M x 64-bit"; bash tools/ab.sh mih 100000000 4096
} > gpurun_out/r02_ab1.log 2>&1
tail -60 gpurun_out/r02_ab1.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bmih_verify -s 12 -c 4 -o gpurun_out/r02_verify_w2 -f python tools/probe.py mih 125000000 1024 bits=128 m=8 reps=1 > gpurun_out/r02_ncu_w2.log 2>&1; tail -2 gpurun_out/r02_ncu_w2.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bmih_verify -s 6 -c 3 -o gpurun_out/r02_verify_w4 -f python tools/probe.py mih 60000000 256 bits=256 m=16 k=1000 r=3 reps=1 > gpurun_out/r02_ncu_w4.log 2>&1; tail -2 gpurun_out/r02_ncu_w4.log
