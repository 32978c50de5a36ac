#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
T=r02b
timeout 900 python tools/stress_probe.py reps=12 > gpurun_out/${T}_stress1.log 2>&1; cat gpurun_out/${T}_stress1.log | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_linear.py tests/test_gpu_scan_batched.py tests/test_gpu_tc.py -x -q > gpurun_out/${T}_pytest3.log 2>&1; tail -n 3 gpurun_out/${T}_pytest3.log
