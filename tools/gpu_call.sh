#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_bmih.py -x -q -k "speculation_by_default" > gpurun_out/r02d_pytest2.log 2>&1; tail -n 12 gpurun_out/r02d_pytest2.log
