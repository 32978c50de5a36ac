timeout 900 python -m pytest tests/test_gpu_bmih.py tests/test_gpu_scan_batched.py -x -q -m gpu 2>&1 | tail -3
VC_BENCH_Q=4096 VC_BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'verify_ms', round(d['roofline']['kernel_ms'],2), 'launches', d['gpu_launches'], d['parity_selfcheck'])
print([ (round(s['measured_ms'],2), round(s['popc_ms'],2), s['bound']) for s in d['roofline']['combined']['steps']], d['roofline']['combined']['frac'])
print({k:(round(v['queries_per_s']),round(v['frac_of_peak'],3), round(v['pair_rate_T_per_s'],2)) for k,v in d['scan'].items()})
"
