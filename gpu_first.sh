set -x
VC_BENCH_SKIP_CPU=1 python bench.py --steps 2 --warmup 3 > gpurun_out/r01_bench_plain.json 2> gpurun_out/r01_bench_plain.err && \
VC_BENCH_SKIP_CPU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r01_bench_ncu.out 2>&1
python tools/scan_probe.py mih 1000000000 4096 > gpurun_out/p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bmih_verify -s 8 -c 1 -o gpurun_out/r01_bmih_verify python tools/scan_probe.py mih 1000000000 4096 > gpurun_out/n1.log 2>&1
python tools/scan_probe.py linear 1000000000 1 > gpurun_out/p2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -o gpurun_out/r01_scan_b1 python tools/scan_probe.py linear 1000000000 1 > gpurun_out/n2.log 2>&1
python tools/scan_probe.py linear 1000000000 1024 scan.waves=8 > gpurun_out/p3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 2 -c 1 -o gpurun_out/r01_scan_b1024 python tools/scan_probe.py linear 1000000000 1024 scan.waves=8 > gpurun_out/n3.log 2>&1
python tools/scan_probe.py mih 1000000000 256 mih.batched=0 > gpurun_out/p4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mih_search -s 2 -c 1 -o gpurun_out/r01_mih_perquery python tools/scan_probe.py mih 1000000000 256 mih.batched=0 > gpurun_out/n4.log 2>&1
cat gpurun_out/p1.log gpurun_out/p2.log gpurun_out/p3.log gpurun_out/p4.log
