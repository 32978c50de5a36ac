timeout 900 python -m pytest tests/test_gpu_bmih.py tests/test_gpu_mih.py -m gpu -x -q 2>&1 | tail -4
python tools/scan_probe.py mih 1000000000 4096
python tools/scan_probe.py mih 1000000000 4096 mih.cpi_steps=2
python tools/scan_probe.py mih 1000000000 4096 mih.cpi_steps=4
python tools/scan_probe.py mih 1000000000 1024 mih.batched=0
VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 VC_BENCH_Q=4096 python bench.py --steps 5 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','integer_pipe','gpu_launches')})"
