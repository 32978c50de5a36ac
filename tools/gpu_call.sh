#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest20.log 2>&1; tail -n 4 gpurun_out/r02_pytest20.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -n 2 gpurun_out/r02_bench_final.err; cut -c1-600 gpurun_out/r02_bench_final.json
