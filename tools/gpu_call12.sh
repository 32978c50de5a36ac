#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest12.log 2>&1; tail -3 gpurun_out/r02_pytest12.log
python tools/clustered_probe.py > gpurun_out/r02_clustered.json 2> gpurun_out/r02_clustered.err; tail -2 gpurun_out/r02_clustered.err; cat gpurun_out/r02_clustered.json
python tools/clustered_probe.py n_centres=200 and_words=5 batch=1024 > gpurun_out/r02_clustered2.json 2> gpurun_out/r02_clustered2.err; tail -2 gpurun_out/r02_clustered2.err; cat gpurun_out/r02_clustered2.json
