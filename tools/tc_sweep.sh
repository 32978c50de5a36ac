timeout 200 python tools/scan_probe.py mih 1000000000 4096 2>&1 | head -2
VC_BENCH_Q=4096 VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 timeout 300 python bench.py --steps 4 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'verify_ms', round(d['roofline']['kernel_ms'],2), 'launches', d['gpu_launches'], d['parity_selfcheck'])
"
