"""Copies the outputs of tools/measure.sh (gpurun_out/r<NN>_*) into profiles/ and rebuilds the derived summaries:
ncu_r<NN>.md (verify-kernel sections), traffic_r<NN>.json, launches_step_r<NN>.md.   python tools/refresh_profiles.py [NN]
bench.py reads the newest profiles/traffic_r*.json for `roofline.traffic`."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
RN = "r%02d" % int(sys.argv[1]) if len(sys.argv) > 1 else "r01"
for src, dst in [(RN + "_bench.json", "bench_%s.json" % RN), (RN + "_bench_ref.json", "bench_ref_%s.json" % RN), (RN + "_launches_bench.csv", "launches_bench_%s.csv" % RN),
                 (RN + "_configs.json", "configs_%s.json" % RN), (RN + "_scan_sweep.json", "scan_sweep_%s.json" % RN)]:
    if os.path.getsize(os.path.join(G, src)) > 0:
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# ---- ncu summary of the verify launches; the scan sections of the old file are kept -------------------------------------
tmp = "/tmp/ncu_new.md"
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), tmp, os.path.join(G, RN + "_bmih_verify.ncu-rep")], check=True)
new = open(tmp).read().rstrip("\n").split("\n")
old = open(os.path.join(P, "ncu_%s.md" % RN) if os.path.exists(os.path.join(P, "ncu_%s.md" % RN)) else os.path.join(P, "ncu_r01.md")).read().split("\n")
idx = [i for i, l in enumerate(old) if l.startswith("## ") and "bmih_verify" not in l]
open(os.path.join(P, "ncu_%s.md" % RN), "w").write("\n".join(new + [""] + (old[idx[0]:] if idx else [])))

def g(sec, key):
    m = re.search(r"\| %s \| ([0-9.]+) \| (\w*)" % re.escape(key), sec)
    return float(m.group(1)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(m.group(2), 1)
launches = []
for sec in open(tmp).read().split("## %s_bmih_verify.ncu-rep" % RN)[1:]:
    launches.append(dict(ms=g(sec, "gpu__time_duration.sum"), dram_read_bytes=g(sec, "dram__bytes_read.sum"), dram_write_bytes=g(sec, "dram__bytes_write.sum"),
                         xu_pct=g(sec, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                         alu_pct=g(sec, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                         issue_active_pct=g(sec, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         warps_active_pct=g(sec, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                         registers=g(sec, "launch__registers_per_thread")))
tpath = os.path.join(P, "traffic_%s.json" % RN)
t = json.load(open(tpath if os.path.exists(tpath) else os.path.join(P, "traffic_r01.json")))
t["launches"] = launches
t["traffic_bytes_per_search"] = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches)
json.dump(t, open(tpath, "w"), indent=1)

# ---- launch list of the first timed step -----------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(P, "launches_bench_%s.csv" % RN))) if len(r) > 10]
hdr = rows[0]
ik, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
data = [(int(r[iid]), r[ik], float(r[iv].replace(",", ""))) for r in rows[1:] if r[iid].isdigit()]
inits = [i for i, d in enumerate(data) if "bmih_init" in d[1]]
# order inside bench.py: 3 warm-ups, the 8-query self-check (MIH, then scan), [more searches], then warm-up + timed steps; the timed
# steps are full-size searches: take the first full-size search after the last small one
def verify_ms(i0, i1): return sum(d[2] for d in data[i0:i1] if "bmih_verify" in d[1]) / 1e6
spans = [(inits[i], inits[i + 1]) for i in range(len(inits) - 1)]
small = [i for i, (a, b) in enumerate(spans) if verify_ms(a, b) < 1.0]
first = small[-1] + 1 + 3                     # three warm-up searches follow the self-check
a, b = spans[first]
step = data[a:b]
agg = collections.OrderedDict()
for _, k, ns in step:
    k = re.sub(r"\(.*", "", re.sub(r"^void ", "", k)).replace("vc::", "")
    e = agg.setdefault(k, [0, 0.0]); e[0] += 1; e[1] += ns / 1e6
tot = sum(e[1] for e in agg.values())
bj = json.loads(open(os.path.join(P, "bench_%s.json" % RN)).readline())
out = ["# Launch list of one timed bench step (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n",
       "command: `VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 python bench.py --steps 2 --warmup 3` (1 B x 64-bit codes, batch 4096, k=100, 1 GPU); "
       "full CSV: launches_bench_%s.csv (launches %d .. %d = the first timed step)\n" % (RN, step[0][0], step[-1][0]),
       "| kernel | launches | ms | share |", "|---|---|---|---|"]
for k, (n, ms) in agg.items():
    out.append("| %s | %d | %.3f | %.1f%% |" % (k, n, ms, 100 * ms / tot))
out.append("| total | %d | %.3f | 100%% |" % (len(step), tot))
vms = sum(e[1] for k, e in agg.items() if "verify" in k)
out.append("\nThe same step timed with CUDA events inside bench.py (no profiler, profiles/bench_%s.json): %.1f ms per step, verify kernels %.1f ms = %.0f %% - "
           "the share agrees (ncu: %.1f %%)." % (RN, bj["ms_per_step"], bj["roofline"]["kernel_ms"], 100 * bj["roofline"]["kernel_ms"] / bj["ms_per_step"], 100 * vms / tot))
open(os.path.join(P, "launches_step_%s.md" % RN), "w").write("\n".join(out) + "\n")
print("\n".join(out))
print(json.dumps(launches, indent=0)[:1500])
