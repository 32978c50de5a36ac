# Round-end measurements on one B200 (run under gpurun from the repo root): ROUND=02 bash tools/measure.sh
# Outputs in gpurun_out/r<ROUND>_*; tools/refresh_profiles.py <ROUND> then copies / summarises them into profiles/.
R=r${ROUND:-01}
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err
VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${R}_bench_ncu.out 2>&1
ncu --set full --import-source on --clock-control none -k regex:bmih_verify -s 6 -c 3 -o gpurun_out/${R}_bmih_verify -f python tools/probe.py mih 1000000000 4096 > gpurun_out/${R}_ncu_verify.log 2>&1
tail -2 gpurun_out/${R}_ncu_verify.log
python tools/bench_configs.py > gpurun_out/${R}_configs.json 2> gpurun_out/${R}_configs.err; tail -2 gpurun_out/${R}_configs.err
python tools/bench_configs.py sweep > gpurun_out/${R}_scan_sweep.json 2> gpurun_out/${R}_scan_sweep.err; tail -2 gpurun_out/${R}_scan_sweep.err
