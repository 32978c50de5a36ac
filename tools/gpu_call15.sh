#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_bmih.py tests/test_gpu_mih.py tests/test_golden.py tests/test_gpu_host_cli.py -m gpu -x -q > gpurun_out/r02_pytest15.log 2>&1; tail -n 15 gpurun_out/r02_pytest15.log
python tools/probe.py mih 1000000000 4096 approx=1 > gpurun_out/r02_approx_b.log 2>&1; tail -n 1 gpurun_out/r02_approx_b.log
python tools/probe.py mih 1000000000 16384 approx=1 >> gpurun_out/r02_approx_b.log 2>&1; tail -n 1 gpurun_out/r02_approx_b.log
python tools/probe.py mih 100000000 4096 approx=1 >> gpurun_out/r02_approx_b.log 2>&1; tail -n 1 gpurun_out/r02_approx_b.log
python tools/probe.py mih 100000000 4096 approx=1 mih.batched=0 >> gpurun_out/r02_approx_b.log 2>&1; tail -n 1 gpurun_out/r02_approx_b.log
