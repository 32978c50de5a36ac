#!/bin/bash
# round 2, GPU call 5 (one GPU): bench.py with the 16 K batch, tensor-core v1 parity + timings, scan batch 1 through the verify kernel
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err; head -c 300 gpurun_out/r02_bench_b.json; echo
python -m pytest tests/test_gpu_tc.py tests/test_gpu_linear.py tests/test_gpu_scan_batched.py tests/test_gpu_persist.py -m gpu -x -q > gpurun_out/r02_pytest5.log 2>&1; tail -4 gpurun_out/r02_pytest5.log
{
echo "== scan 1 B x 64-bit: POPC verify kernel vs tensor-core v4 (scan.tc=1) vs tensor-core v1 (scan.tc=2)"
for b in 64 256 1024 4096; do for tc in 0 1 2; do python tools/probe.py linear 1000000000 $b reps=1 scan.tc=$tc | tail -1; done; done
echo "== scan batch 1 / 2 / 4 / 8: ring kernel vs verify kernel"
for b in 1 2 4 8; do python tools/probe.py linear 1000000000 $b scan.batched_min=1 | tail -1; python tools/probe.py linear 1000000000 $b scan.batched=0 | tail -1; done
echo "== 128-bit scan 125 M batch 1 / 2 / 4"; for b in 1 2 4; do python tools/probe.py linear 125000000 $b bits=128 scan.batched_min=1 | tail -1; python tools/probe.py linear 125000000 $b bits=128 scan.batched=0 | tail -1; done
} > gpurun_out/r02_ab5.log 2>&1
grep -c kernel_ms gpurun_out/r02_ab5.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bmih_verify_tc1 -c 1 -o gpurun_out/r02_tc1_scan1024 -f python tools/probe.py linear 1000000000 1024 reps=1 scan.tc=2 > gpurun_out/r02_ncu_tc1.log 2>&1; tail -2 gpurun_out/r02_ncu_tc1.log
