#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mih.py tests/test_gpu_sharded.py -m gpu -x -q -k "chunk or python_hook or approximate" > gpurun_out/r02_pytest19.log 2>&1; tail -n 5 gpurun_out/r02_pytest19.log
for b in 1 2 4 8; do python tools/probe.py linear 125000000 $b reps=20 2>&1 | tail -n 1; done > gpurun_out/r02_lin125.log; cat gpurun_out/r02_lin125.log
