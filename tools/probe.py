"""One search workload on one GPU, timed per kernel - for A/B runs (tools/ab.sh) and ncu captures.
    python tools/probe.py <mih|linear> <n_codes> <batch> [bits=64] [m=4] [k=100] [r=-1] [approx=0] [reps=3] [check=0] [shards=1] [param=value ...]
r >= 0: fixed-radius search (config C5).  check=Q: the first Q queries are also answered by the brute-force scan of the same
index and must agree bit for bit (exact mode only).  shards=G: the index behaves like ONE of G id-shards of a G times larger
database - every cross-shard sum is replaced by "times G" (index parameter xchg.emulate; statistically what G shards of uniform codes
give), so the per-shard kernels of an 8-GPU run can be timed and profiled on one GPU; answers are not meaningful then.  Unknown name=value pairs are index parameters (vc_index_set_param)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from verticut_b200 import capi

mode, n, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
opt = {"bits": 64, "m": 4, "k": 100, "r": -1, "approx": 0, "reps": 3, "check": 0, "shards": 1}
knobs = []
for a in sys.argv[4:]:
    name, v = a.split("=")
    if name in opt:
        opt[name] = int(v)
    else:
        knobs.append((name, int(v)))
bits, m, k, r = opt["bits"], opt["m"], opt["k"], opt["r"]
ix = capi.Index(bits, m if mode == "mih" else 0)
if opt["shards"] > 1:
    ix.set_param("id_stride", opt["shards"])     # ids dealt round-robin over the shards, as bench.py --gpus G does
ix.add_synthetic(n, 12345)
ix.build()
ix.set_param("profile", 1)
for name, v in knobs:
    ix.set_param(name, v)
if opt["shards"] > 1:
    ix.set_param("xchg.emulate", opt["shards"])
q = np.random.default_rng(1).integers(0, 256, size=(B, bits // 8), dtype=np.uint8)


def once(stats=False):
    if mode == "mih":
        return ix.search_mih(q, k, approximate=bool(opt["approx"]), max_radius=r, with_stats=stats)
    return ix.search_linear(q, k)


res = once(stats=True)
for _ in range(2):
    once()
t0 = time.perf_counter()
for _ in range(opt["reps"]):
    once()
dt = (time.perf_counter() - t0) / opt["reps"]
ns = ix.get_param("last_kernel_ns")
out = {"mode": mode, "n": n, "B": B, "bits": bits, "m": m, "k": k, "r": r, "approx": opt["approx"], "shards": opt["shards"], "kernel_ms": round(ns / 1e6, 3), "e2e_ms": round(dt * 1e3, 3),
       "qps_e2e": round(B / dt, 1)}
if mode == "mih":
    st = res[3]
    batched = bool(ix.get_param("mih.last_batched"))
    try:
        out["redo"] = ix.get_param("mih.last_redo")
    except capi.VerticutError:      # a library built before the parameter existed (A/B against older builds)
        pass
    out.update(batched=batched, mean_radius=float(st["radius"].mean()),
               probes_per_q=float(st["probes"].mean()), cands_per_q=float(st["candidates"].mean()))
    if batched:
        L = ix.get_param("mih.last_levels")
        out["steps_ms"] = [round(ix.get_param("mih.step_ns.%d" % i) / 1e6, 3) for i in range(L)]
        try:      # probes + work items in front of each verify launch; settle + exchange + decide behind it; the whole device-side search
            out["pre_ms"] = [round(ix.get_param("mih.step_pre_ns.%d" % i) / 1e6, 3) for i in range(L)]
            out["post_ms"] = [round(ix.get_param("mih.step_post_ns.%d" % i) / 1e6, 3) for i in range(L)]
            out["search_ms"] = round(ix.get_param("mih.search_ns") / 1e6, 3)
        except capi.VerticutError:
            pass
        out["step_exec_pairs"] = [ix.get_param("mih.step_exec.%d" % i) for i in range(L)]
        out["step_codes"] = [ix.get_param("mih.step_codes.%d" % i) for i in range(L)]
        ex = float(sum(out["step_exec_pairs"]))
        out["exec_pairs_per_s"] = "%.3e" % (ex / (ns * 1e-9))
        out["bucket_GBps"] = round(sum(out["step_codes"]) * (bits // 8) / ns, 1)
    else:
        out["pairs_per_s"] = "%.3e" % (float(st["candidates"].sum()) / (ns * 1e-9))
    if opt["check"] and r < 0:
        c = opt["check"]
        lid, ld, _ = ix.search_linear(q[:c], k)
        out["equals_scan"] = bool(np.array_equal(lid, res[0][:c]) and np.array_equal(ld, res[1][:c]))
else:
    out.update(pairs_per_s="%.3e" % (n * B / (ns * 1e-9)), GBps=round(n * (bits // 8) / ns, 1), batched=bool(ix.get_param("scan.last_batched")),
               grid=ix.get_param("scan.last_grid"), qt=ix.get_param("scan.last_qt"), occ=ix.get_param("scan.last_occ"), stages=ix.get_param("scan.last_stages"))
out["knobs"] = knobs
print(out)
