#!/bin/bash
mkdir -p gpurun_out
python tools/clustered_probe.py n_centres=200 and_words=5 batch=1024 > gpurun_out/r02_clustered2.json 2> gpurun_out/r02_clustered2.err; tail -n 2 gpurun_out/r02_clustered2.err; cat gpurun_out/r02_clustered2.json
python tools/probe.py mih 125000000 16384 shards=8 > gpurun_out/r02_shard8.log 2>&1; tail -n 3 gpurun_out/r02_shard8.log
python tools/probe.py mih 125000000 4096 shards=8 >> gpurun_out/r02_shard8.log 2>&1; tail -n 1 gpurun_out/r02_shard8.log
python tools/probe.py mih 1000000000 4096 approx=1 > gpurun_out/r02_approx.log 2>&1; tail -n 2 gpurun_out/r02_approx.log
python tools/probe.py mih 1000000000 16384 approx=1 >> gpurun_out/r02_approx.log 2>&1; tail -n 1 gpurun_out/r02_approx.log
ncu --set full --import-source on --clock-control none -k regex:bmih_verify -s 6 -c 3 -o gpurun_out/r02_shard8_verify -f python tools/probe.py mih 125000000 16384 shards=8 reps=1 > gpurun_out/r02_ncu_shard8.log 2>&1; tail -n 2 gpurun_out/r02_ncu_shard8.log
python tools/clustered_probe.py n_codes=1000000000 n_centres=200000 batch=16384 > gpurun_out/r02_clustered_1b.json 2> gpurun_out/r02_clustered_1b.err; tail -n 2 gpurun_out/r02_clustered_1b.err; cat gpurun_out/r02_clustered_1b.json
