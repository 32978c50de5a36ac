#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
T=r02b
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_final.log 2>&1; tail -n 4 gpurun_out/${T}_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_final.json 2> gpurun_out/${T}_bench_final.err; tail -n 2 gpurun_out/${T}_bench_final.err; cut -c1-500 gpurun_out/${T}_bench_final.json
{
  timeout 200 python tools/probe.py mih 125000000 4096 bits=128 m=8 2>&1 | tail -1
  timeout 200 python tools/probe.py mih 60000000 256 bits=256 m=16 k=1000 r=3 2>&1 | tail -1
  timeout 200 python tools/probe.py mih 60000000 1024 bits=256 m=16 k=1000 r=3 2>&1 | tail -1
  timeout 200 python tools/probe.py mih 100000000 4096 2>&1 | tail -1
  timeout 200 python tools/probe.py mih 1000000000 4096 approx=1 2>&1 | tail -1
} > gpurun_out/${T}_configs_probe.log 2>&1
cut -c1-400 gpurun_out/${T}_configs_probe.log
VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 VC_BENCH_ORACLE_Q=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_bench_ncu.out 2>&1
tail -n 3 gpurun_out/${T}_launches_bench.csv | cut -c1-200
