python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','parity_selfcheck')})
r=d['roofline']; print({k:r[k] for k in r if k!='combined'}); print(r['combined'])
print(d['integer_pipe']); print(d['cpu_baseline'])
for k,v in d['scan'].items(): print(k, {a:round(b,3) for a,b in v.items()})
PY
