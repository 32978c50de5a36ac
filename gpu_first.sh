timeout 1200 python -m pytest tests/test_gpu_scan_batched.py tests/test_gpu_linear.py -m gpu -x -q 2>&1 | tail -4
python tools/bench_configs.py sweep > gpurun_out/scan_sweep.json 2> gpurun_out/scan_sweep.err; tail -3 gpurun_out/scan_sweep.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/scan_sweep.json'))
for r in d['rows']: print(r['batch'], round(r['queries_per_s_e2e']), round(r['scan_kernel_ms'],2), round(r['hbm_GBps']), '%.2e'%r['tests_per_s'], round(r['frac_of_roofline'],2), r['bound'])
PY
