#!/bin/bash
# round 2, GPU call 10 (one GPU): the whole GPU suite on the current library
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest10.log 2>&1; tail -6 gpurun_out/r02_pytest10.log
