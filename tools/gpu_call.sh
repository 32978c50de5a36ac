#!/bin/bash
# scratch: the command list of the current gpurun call (2 GPUs)
mkdir -p gpurun_out
T=r02b
timeout 400 python -m pytest tests/test_gpu_nccl.py -q > gpurun_out/${T}_pytest2b.log 2>&1; tail -n 3 gpurun_out/${T}_pytest2b.log
