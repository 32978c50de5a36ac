#!/bin/bash
# round 2, GPU call 2 (one GPU): parity suite on the lane-group hit queue, A/B, the refactored bench.py (headline, C2, C4), ncu of the 64-bit kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1; tail -5 gpurun_out/r02_pytest2.log
{
echo "== headline: 1 B x 64-bit, m=4, k=100, batch 4096"; bash tools/ab.sh mih 1000000000 4096 check=4
echo "== C3 shard: 125 M x 128-bit, m=8, k=100, batch 1024"; bash tools/ab.sh mih 125000000 1024 bits=128 m=8 check=4
echo "== C5: 60 M x 256-bit, m=16, k=1000, batch 256, r=3"; bash tools/ab.sh mih 60000000 256 bits=256 m=16 k=1000 r=3
echo "== C2: 100 M x 64-bit"; bash tools/ab.sh mih 100000000 4096
echo "== scan 1 B x 64-bit B=64 / 1024"; python tools/probe.py linear 1000000000 64 | tail -1; python tools/probe.py linear 1000000000 1024 reps=1 | tail -1
} > gpurun_out/r02_ab2.log 2>&1
grep -c kernel_ms gpurun_out/r02_ab2.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err; head -c 600 gpurun_out/r02_bench_a.json; echo
python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; tail -3 gpurun_out/r02_bench_c2.err; head -c 300 gpurun_out/r02_bench_c2.json; echo
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -3 gpurun_out/r02_bench_c4.err; head -c 300 gpurun_out/r02_bench_c4.json; echo
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bmih_verify -s 6 -c 3 -o gpurun_out/r02_verify_w1 -f python tools/probe.py mih 1000000000 4096 reps=1 > gpurun_out/r02_ncu_w1.log 2>&1; tail -2 gpurun_out/r02_ncu_w1.log
