# Round-end measurements on one B200 (run under gpurun from the repo root): ROUND=02 bash tools/measure.sh
# Outputs in gpurun_out/r<ROUND>_*; tools/collect_r02.py + tools/refresh_profiles.py then copy / summarise them into profiles/.
R=r${ROUND:-02}
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err; tail -2 gpurun_out/${R}_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err
python bench.py --config C2 --steps 5 --warmup 3 > gpurun_out/${R}_bench_c2.json 2> gpurun_out/${R}_bench_c2.err
python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/${R}_bench_c4.json 2> gpurun_out/${R}_bench_c4.err
VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 VC_BENCH_ORACLE_Q=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${R}_bench_ncu.out 2>&1
ncu --set full --import-source on --clock-control none -k regex:bmih_verify -s 6 -c 3 -o gpurun_out/${R}_bmih_verify -f python tools/probe.py mih 1000000000 16384 reps=1 > gpurun_out/${R}_ncu_verify.log 2>&1
tail -2 gpurun_out/${R}_ncu_verify.log
python tools/bench_configs.py > gpurun_out/${R}_configs_misc.json 2> gpurun_out/${R}_configs_misc.err; tail -2 gpurun_out/${R}_configs_misc.err
./tools/bin/microbench > gpurun_out/${R}_microbench.log 2>&1
