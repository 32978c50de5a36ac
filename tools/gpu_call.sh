#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench_final.json 2> gpurun_out/r02d_bench_final.err; echo rc=$?; tail -n 1 gpurun_out/r02d_bench_final.err | cut -c1-200; cut -c1-200 gpurun_out/r02d_bench_final.json
