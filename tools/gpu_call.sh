#!/bin/bash
# scratch: the command list of the current gpurun call
mkdir -p gpurun_out
T=r02c
VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_note.json 2> gpurun_out/${T}_bench_note.err; echo rc=$?; tail -n 3 gpurun_out/${T}_bench_note.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02c_bench_note.json').read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["which_bound_binds"])
P
