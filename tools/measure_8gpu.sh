#!/bin/bash
# round 2, the 8-GPU call: BASELINE configs C3 / C5 at the sizes and GPU counts they name, the headline at 8 GPUs with the peer
# exchange and with NCCL, C4 sharded.  Every run checks a few queries against the CPU oracle over the full sharded database.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo8.log 2>&1
nproc > gpurun_out/r02_nproc8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
export VC_BENCH_SKIP_BIG_SCAN=1
timeout 420 $TR --master-port 29601 bench.py --gpus 8 --config C3 --steps 5 --warmup 3 > gpurun_out/r02_bench8_c3.json 2> gpurun_out/r02_bench8_c3.err; tail -2 gpurun_out/r02_bench8_c3.err; head -c 300 gpurun_out/r02_bench8_c3.json; echo
VC_BENCH_ORACLE_Q=2 timeout 600 $TR --master-port 29602 bench.py --gpus 8 --config C5 --steps 5 --warmup 3 > gpurun_out/r02_bench8_c5.json 2> gpurun_out/r02_bench8_c5.err; tail -2 gpurun_out/r02_bench8_c5.err; head -c 300 gpurun_out/r02_bench8_c5.json; echo
timeout 420 $TR --master-port 29603 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench8_peer.json 2> gpurun_out/r02_bench8_peer.err; tail -2 gpurun_out/r02_bench8_peer.err; head -c 300 gpurun_out/r02_bench8_peer.json; echo
VC_XCHG=0 timeout 420 $TR --master-port 29604 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench8_nccl.json 2> gpurun_out/r02_bench8_nccl.err; tail -2 gpurun_out/r02_bench8_nccl.err; head -c 300 gpurun_out/r02_bench8_nccl.json; echo
