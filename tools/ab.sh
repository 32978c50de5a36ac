# A/B of two builds of the C-ABI library on the bench workload (1 B x 64-bit codes, batch 4096), then the GPU suite and the bench
set -x
for lib in tools/bin/ab_base.so verticut_b200/lib/libverticut_gpu.so; do
  VC_GPU_LIB=$PWD/$lib timeout 300 python tools/scan_probe.py mih 1000000000 4096 2>&1 | tail -1
done
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; tail -2 gpurun_out/r01_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r01_bench.json').readline())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'verify_ms', round(d['roofline']['kernel_ms'],2), 'combined', round(d['roofline']['combined']['frac'],3), [round(s['measured_ms'],2) for s in d['roofline']['combined']['steps']], d['parity_selfcheck'])
PY
