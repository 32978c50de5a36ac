#!/bin/bash
# round 2, GPU call 11 (two GPUs): the committed library through the sharded parity tests and bench.py --gpus 2 (default exchange, all-peer, all-NCCL)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_sharded.py -m gpu -q > gpurun_out/r02_pytest11.log 2>&1; tail -4 gpurun_out/r02_pytest11.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export VC_BENCH_SKIP_BIG_SCAN=1
$TR --master-port 29711 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench2_final.json 2> gpurun_out/r02_bench2_final.err; tail -1 gpurun_out/r02_bench2_final.err; head -c 250 gpurun_out/r02_bench2_final.json; echo
VC_XCHG_RESULTS=1 $TR --master-port 29712 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench2_peerres.json 2> gpurun_out/r02_bench2_peerres.err; head -c 250 gpurun_out/r02_bench2_peerres.json; echo
VC_XCHG=0 $TR --master-port 29713 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench2_nccl_final.json 2> gpurun_out/r02_bench2_nccl_final.err; head -c 250 gpurun_out/r02_bench2_nccl_final.json; echo
