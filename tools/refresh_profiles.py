"""Copies the outputs of tools/measure.sh (gpurun_out/r<NN>_*) into profiles/ and rebuilds the derived summaries:
ncu_r<NN>_headline.md (the verify launches of one headline search under ncu --set full), traffic_r<NN>.json (their DRAM traffic; bench.py
reads the newest profiles/traffic_r*.json for `roofline.traffic`), launches_bench_r<NN>.csv + launches_step_r<NN>.md (launch list of one
timed bench step).   python tools/refresh_profiles.py [NN]"""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
RN = "r%02d" % int(sys.argv[1]) if len(sys.argv) > 1 else "r02"


def first_existing(*names):
    for n in names:
        p = os.path.join(G, n)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            return p
    return None


bench_path = first_existing(RN + "_bench_final.json", RN + "_bench.json")
bj = json.loads(open(bench_path).read().strip().splitlines()[-1])
shutil.copy(bench_path, os.path.join(P, "bench_%s.json" % RN))
for src, dst in [(RN + "_bench_ref.json", "bench_ref_%s.json" % RN), (RN + "_launches_bench.csv", "launches_bench_%s.csv" % RN)]:
    p = first_existing(src)
    if p:
        shutil.copy(p, os.path.join(P, dst))

# ---- ncu summary of the verify launches of one headline search ---------------------------------------------------
rep = os.path.join(G, RN + "_bmih_verify.ncu-rep")
t = None
cfg = bj["config"]
if os.path.exists(rep):      # (no new capture: the committed ncu summary and traffic file stay)
    tmp = os.path.join(P, "ncu_%s_headline.md" % RN)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), tmp, os.path.join(G, RN + "_bmih_verify.ncu-rep")], check=True)


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
    cfg = bj["config"]
    t = {"what": "DRAM traffic of the bmih_verify_kernel launches of ONE headline search under `ncu --set full --clock-control none` "
                 "(tools/measure.sh: python tools/probe.py mih %d %d reps=1, launches 7..9 = the third search)" % (cfg["n_codes"], cfg["batch"]),
         "config": {"n_codes": cfg["n_codes"], "batch": cfg["batch"], "k": 100, "n_gpus": 1},
         "launches": launches,
         "traffic_bytes_per_search": sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches),
         "algorithmic_bytes_per_search": bj["roofline"]["algorithmic_bytes_per_launch"]}
    t["traffic_over_algorithmic"] = t["traffic_bytes_per_search"] / t["algorithmic_bytes_per_search"]
    json.dump(t, open(os.path.join(P, "traffic_%s.json" % RN), "w"), indent=1)

# ---- launch list of the first timed step -----------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(P, "launches_bench_%s.csv" % RN))) if len(r) > 10]
hdr = rows[0]
ik, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
data = [(int(r[iid]), r[ik], float(r[iv].replace(",", ""))) for r in rows[1:] if r[iid].isdigit()]
inits = [i for i, d in enumerate(data) if "bmih_init" in d[1]]


def verify_ms(i0, i1):
    return sum(d[2] for d in data[i0:i1] if "bmih_verify" in d[1]) / 1e6


# full-size searches are the ones whose verify launches add up to more than half of the bench's verify time; the timed steps follow
# three warm-up searches: take the fourth full-size search
spans = [(inits[i], inits[i + 1]) for i in range(len(inits) - 1)]
full = [s for s in spans if verify_ms(*s) > 0.5 * bj["roofline"]["kernel_ms"]]
a, b = full[min(3, len(full) - 1)]          # the first timed step follows three warm-up searches
step = data[a:b]
agg = collections.OrderedDict()
for _, k, ns in step:
    k = re.sub(r"\(.*", "", re.sub(r"^void ", "", k)).replace("vc::", "")
    e = agg.setdefault(k, [0, 0.0]); e[0] += 1; e[1] += ns / 1e6
tot = sum(e[1] for e in agg.values())
out = ["# Launch list of one bench step (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n",
       "command: `VC_BENCH_SKIP_CPU=1 VC_BENCH_SKIP_BIG_SCAN=1 VC_BENCH_ORACLE_Q=0 python bench.py --steps 2 --warmup 3` (1 B x 64-bit codes, batch %d, k=100, 1 GPU); "
       "full CSV: launches_bench_%s.csv (launches %d .. %d = one full-size search)\n" % (cfg["batch"], RN, step[0][0], step[-1][0]),
       "| kernel | launches | ms | share |", "|---|---|---|---|"]
for k, (n, ms) in agg.items():
    out.append("| %s | %d | %.3f | %.1f%% |" % (k, n, ms, 100 * ms / tot))
out.append("| total | %d | %.3f | 100%% |" % (len(step), tot))
vms = sum(e[1] for k, e in agg.items() if "verify" in k)
out.append("\nThe same step timed with CUDA events inside bench.py (no profiler, profiles/bench_%s.json): %.1f ms per step, verify kernels %.1f ms = %.0f %% - "
           "the share agrees (ncu: %.1f %%)." % (RN, bj["ms_per_step"], bj["roofline"]["kernel_ms"], 100 * bj["roofline"]["kernel_ms"] / bj["ms_per_step"], 100 * vms / tot))
open(os.path.join(P, "launches_step_%s.md" % RN), "w").write("\n".join(out) + "\n")
print("\n".join(out))
if t is not None:
    print(json.dumps({k: v for k, v in t.items() if k != "launches"}, indent=0))
