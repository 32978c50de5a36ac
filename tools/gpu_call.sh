#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_sharded.py -x -q -k "peer_windows" > gpurun_out/r02d_pytest3.log 2>&1; echo rc=$?; tail -n 6 gpurun_out/r02d_pytest3.log | cut -c1-300
