#!/bin/bash
# round 2, GPU call 3 (one GPU): parity of the per-warp-struct kernel + new host tests, A/B, batch-size sweep of the headline, C5 knobs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bmih.py tests/test_gpu_mih.py tests/test_gpu_sharded.py tests/test_gpu_scan_batched.py tests/test_gpu_host_cli.py tests/test_gpu_rpc_server.py -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; tail -5 gpurun_out/r02_pytest3.log
{
echo "== headline: 1 B x 64-bit, m=4, k=100, batch 4096"; bash tools/ab.sh mih 1000000000 4096 check=4
echo "== C3 shard: 125 M x 128-bit, m=8, k=100, batch 1024"; bash tools/ab.sh mih 125000000 1024 bits=128 m=8 check=4
echo "== C5: 60 M x 256-bit, m=16, k=1000, batch 256, r=3"; bash tools/ab.sh mih 60000000 256 bits=256 m=16 k=1000 r=3
echo "== C5 wide"; python tools/probe.py mih 60000000 256 bits=256 m=16 k=1000 r=3 mih.wide=1 | tail -1
echo "== C2: 100 M x 64-bit"; bash tools/ab.sh mih 100000000 4096
echo "== headline batch sweep (in-tree)"; for q in 1024 2048 8192 16384 32768; do python tools/probe.py mih 1000000000 $q reps=2 | tail -1; done
echo "== C3 shard batch sweep (in-tree)"; for q in 2048 4096; do python tools/probe.py mih 125000000 $q bits=128 m=8 reps=2 | tail -1; done
echo "== shard-size sweep 125 M x 64-bit batch 4096 (what one of 8 GPUs sees, without the exchange)"; bash tools/ab.sh mih 125000000 4096
} > gpurun_out/r02_ab3.log 2>&1
grep -c kernel_ms gpurun_out/r02_ab3.log
./tools/bin/microbench > gpurun_out/r02_microbench.log 2>&1; tail -5 gpurun_out/r02_microbench.log
