#!/bin/bash
mkdir -p gpurun_out
python tools/probe.py mih 125000000 16384 shards=8 > gpurun_out/r02_shard8.log 2>&1; tail -n 1 gpurun_out/r02_shard8.log
python tools/probe.py mih 125000000 4096 shards=8 >> gpurun_out/r02_shard8.log 2>&1; tail -n 1 gpurun_out/r02_shard8.log
python tools/probe.py mih 1000000000 16384 >> gpurun_out/r02_shard8.log 2>&1; tail -n 1 gpurun_out/r02_shard8.log
ncu --set full --import-source on --clock-control none -k regex:bmih_verify -s 6 -c 3 -o gpurun_out/r02_shard8_verify -f python tools/probe.py mih 125000000 16384 shards=8 reps=1 > gpurun_out/r02_ncu_shard8.log 2>&1; tail -n 2 gpurun_out/r02_ncu_shard8.log
