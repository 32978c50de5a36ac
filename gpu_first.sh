for n in 8 4; do
VC_BENCH_SKIP_BIG_SCAN=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 2> gpurun_out/scale_$n.err | tail -1 > gpurun_out/scale_$n.json
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_$n.json').read())
print($n, {k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','parity_selfcheck')}, d['roofline']['kernel_ms'], d['roofline']['probes_per_query'], d['integer_pipe']['frac_of_popc_peak'], {k:(round(v['queries_per_s']),round(v['hbm_GBps_per_gpu'])) for k,v in d['scan'].items()})
PY
done
