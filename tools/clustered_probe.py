"""Clustered (non-uniform) codes: the workload real image hashes look like - many near-duplicates around a limited number of centres,
so buckets are skewed (a centre's substring bucket is huge, most are empty), the k-th distance is small and distance ties are heavy.
Every number elsewhere in profiles/ is on uniform codes; this checks that nothing falls off a cliff (candidate-buffer overflow -> only
those queries are redone, giant buckets -> many work items) and that the answers stay exact.
    python tools/clustered_probe.py [n_codes=100000000] [n_centres=20000] [and_words=4] [batch=4096] [k=100]
Codes = centre[c] with every bit flipped with probability 2^-and_words (the AND of that many random words: 1/16 by default, so two
members of a cluster differ in ~ 8 bits); queries are
drawn the same way.  Prints one JSON line per path: batched MIH, brute-force scan of the same queries, and whether they agree."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from verticut_b200 import capi

opt = {"n_codes": 100_000_000, "n_centres": 20_000, "and_words": 4, "batch": 4096, "k": 100}
for a in sys.argv[1:]:
    name, v = a.split("=")
    opt[name] = int(v)
n, nc, aw, B, k = opt["n_codes"], opt["n_centres"], opt["and_words"], opt["batch"], opt["k"]
flip = 0.5 ** aw
rng = np.random.default_rng(2024)
centres = rng.integers(0, 2 ** 64, size=nc, dtype=np.uint64)


def draw(count):
    out = np.empty(count, dtype=np.uint64)
    step = 10_000_000
    for s in range(0, count, step):
        m = min(step, count - s)
        noise = rng.integers(0, 2 ** 64, size=m, dtype=np.uint64)
        for _ in range(aw - 1):
            noise &= rng.integers(0, 2 ** 64, size=m, dtype=np.uint64)
        out[s:s + m] = centres[rng.integers(0, nc, size=m)] ^ noise
    return out


t0 = time.perf_counter()
codes = draw(n)
queries = draw(B)
t_gen = time.perf_counter() - t0
ix = capi.Index(64, 4)
for s in range(0, n, 20_000_000):
    ix.add(codes[s:s + 20_000_000].view(np.uint8).reshape(-1, 8))
ix.build()
ix.set_param("profile", 1)
q8 = queries.view(np.uint8).reshape(-1, 8)
ids, dists, counts, st = ix.search_mih(q8, k)
for _ in range(2):
    ix.search_mih(q8, k, with_stats=False)
t0 = time.perf_counter()
for _ in range(3):
    ix.search_mih(q8, k, with_stats=False)
dt = (time.perf_counter() - t0) / 3
L = ix.get_param("mih.last_levels") if ix.get_param("mih.last_batched") else 0
out = {"data": "clustered: %d codes around %d centres, bit flip probability %.4f" % (n, nc, flip), "batch": B, "k": k, "host_generation_s": round(t_gen, 1),
       "mih": {"queries_per_s_e2e": B / dt, "ms_per_batch": dt * 1e3, "batched": bool(ix.get_param("mih.last_batched")),
               "queries_redone_by_the_per_query_kernel": ix.get_param("mih.last_redo"),
               "verify_kernel_ms": ix.get_param("last_kernel_ns") / 1e6,
               "steps_ms": [round(ix.get_param("mih.step_ns.%d" % i) / 1e6, 3) for i in range(L)],
               "mean_radius": float(st["radius"].mean()), "max_radius": int(st["radius"].max()),
               "candidates_per_query_mean": float(st["candidates"].mean()), "candidates_per_query_max": int(st["candidates"].max()),
               "kth_distance_mean": float(dists[:, k - 1].mean()), "kth_distance_max": int(dists[:, k - 1].max())}}
chk = min(B, 256)
t0 = time.perf_counter()
lid, ld, lc = ix.search_linear(q8[:chk], k)
out["scan"] = {"queries": chk, "ms": (time.perf_counter() - t0) * 1e3}
out["mih_equals_scan"] = bool(np.array_equal(lid, ids[:chk]) and np.array_equal(ld, dists[:chk]) and np.array_equal(lc, counts[:chk]))
print(json.dumps(out))
