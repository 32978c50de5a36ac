#!/bin/bash
# round 2, second 8-GPU call: the headline with the hybrid exchange (results over peer memory, large histogram sums over NCCL) and the
# full bootstrap sample per shard; pure NCCL and pure peer memory beside it; config C4 sharded
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
export VC_BENCH_SKIP_BIG_SCAN=1
timeout 420 $TR --master-port 29611 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench8_hybrid.json 2> gpurun_out/r02_bench8_hybrid.err; tail -1 gpurun_out/r02_bench8_hybrid.err; head -c 250 gpurun_out/r02_bench8_hybrid.json; echo
VC_XCHG=0 timeout 420 $TR --master-port 29612 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench8_nccl2.json 2> gpurun_out/r02_bench8_nccl2.err; head -c 250 gpurun_out/r02_bench8_nccl2.json; echo
VC_BENCH_Q=4096 VC_BENCH_ORACLE_Q=0 timeout 420 $TR --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench8_q4096.json 2> gpurun_out/r02_bench8_q4096.err; head -c 250 gpurun_out/r02_bench8_q4096.json; echo
timeout 420 $TR --master-port 29614 bench.py --gpus 8 --config C4 --steps 2 --warmup 3 > gpurun_out/r02_bench8_c4.json 2> gpurun_out/r02_bench8_c4.err; tail -1 gpurun_out/r02_bench8_c4.err; head -c 250 gpurun_out/r02_bench8_c4.json; echo
