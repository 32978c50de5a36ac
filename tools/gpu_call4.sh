#!/bin/bash
# round 2, GPU call 4 (two GPUs): NCCL / peer-exchange parity tests, kernel parity, bench at 2 GPUs (headline, C3, C5), single-GPU A/B
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.log 2>&1
python -m pytest tests/test_gpu_nccl.py tests/test_gpu_sharded.py tests/test_gpu_bmih.py tests/test_gpu_mih.py -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1; tail -5 gpurun_out/r02_pytest4.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export VC_BENCH_SKIP_BIG_SCAN=1
$TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench2_peer.json 2> gpurun_out/r02_bench2_peer.err; tail -2 gpurun_out/r02_bench2_peer.err; head -c 400 gpurun_out/r02_bench2_peer.json; echo
VC_XCHG=0 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench2_nccl.json 2> gpurun_out/r02_bench2_nccl.err; tail -2 gpurun_out/r02_bench2_nccl.err; head -c 400 gpurun_out/r02_bench2_nccl.json; echo
VC_XCHG=0 VC_NCCL_DIRECT=0 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench2_torch.json 2> gpurun_out/r02_bench2_torch.err; tail -2 gpurun_out/r02_bench2_torch.err; head -c 400 gpurun_out/r02_bench2_torch.json; echo
$TR --master-port 29514 bench.py --gpus 2 --config C3 --steps 3 --warmup 3 > gpurun_out/r02_bench2_c3.json 2> gpurun_out/r02_bench2_c3.err; tail -2 gpurun_out/r02_bench2_c3.err; head -c 400 gpurun_out/r02_bench2_c3.json; echo
$TR --master-port 29515 bench.py --gpus 2 --config C5 --steps 3 --warmup 3 > gpurun_out/r02_bench2_c5.json 2> gpurun_out/r02_bench2_c5.err; tail -2 gpurun_out/r02_bench2_c5.err; head -c 400 gpurun_out/r02_bench2_c5.json; echo
{
echo "== headline: 1 B x 64-bit, m=4, k=100, batch 4096"; CUDA_VISIBLE_DEVICES=0 bash tools/ab.sh mih 1000000000 4096 check=4
echo "== C3 shard"; CUDA_VISIBLE_DEVICES=0 python tools/probe.py mih 125000000 1024 bits=128 m=8 check=4 | tail -1
echo "== C5 r=3"; CUDA_VISIBLE_DEVICES=0 python tools/probe.py mih 60000000 256 bits=256 m=16 k=1000 r=3 | tail -1
echo "== C2"; CUDA_VISIBLE_DEVICES=0 python tools/probe.py mih 100000000 4096 | tail -1
echo "== scan small batches through the verify kernel"; for b in 2 4; do CUDA_VISIBLE_DEVICES=0 python tools/probe.py linear 1000000000 $b scan.batched_min=2 | tail -1; CUDA_VISIBLE_DEVICES=0 python tools/probe.py linear 1000000000 $b | tail -1; done
} > gpurun_out/r02_ab4.log 2>&1
grep -c kernel_ms gpurun_out/r02_ab4.log
