#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
T=r02b
{
  timeout 400 python tools/determinism_probe.py 2>&1 | tail -22
  timeout 400 python tools/determinism_probe.py n=125000000 bits=128 m=8 reps=3 2>&1 | tail -22
} > gpurun_out/${T}_determinism1.log 2>&1
cut -c1-260 gpurun_out/${T}_determinism1.log
