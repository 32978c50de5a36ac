#!/bin/bash
# scratch: the command list of the current gpurun call (2 GPUs)
mkdir -p gpurun_out
T=r02c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export VC_BENCH_SKIP_BIG_SCAN=1
timeout 300 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/${T}_pytest2.log 2>&1; tail -n 2 gpurun_out/${T}_pytest2.log
timeout 300 $TR --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err; tail -1 gpurun_out/${T}_bench2.err; head -c 300 gpurun_out/${T}_bench2.json; echo
