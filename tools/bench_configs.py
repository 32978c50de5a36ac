"""One-GPU measurements that bench.py --config does not cover: C1 (the reference's own CPU-runnable case) on the GPU, the sparse
(s = 32) variant of C5, and the reference's NATIVE shape - 128-bit codes, 4 tables of 32-bit substrings, 100 M codes, k = 100
(src/run_distributed_search.py:8-12, src/image_search_constants.h:14) - as a fixed-radius sweep (on uniform codes its exact k-NN
needs radius 9, i.e. 1.7 * 10^8 bitmap probes per query: not a workload) beside the brute-force scan of the same
index.  Each exact case is checked for self-consistency (MIH == linear scan on the same index) on a few queries.
    python tools/bench_configs.py > profiles/configs_r02.json        (C2 - C5 at full size: bench.py --config ...)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from verticut_b200 import capi  # noqa: E402

PEAK = 6554.6


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def run(name, n, bits, m, k, nq, max_radius=-1, check=8, scan_batches=(1, 256)):
    nbytes = bits // 8
    out = {"config": name, "n_codes": n, "code_bits": bits, "tables": m, "k": k, "batch": nq}
    ix = capi.Index(bits, m)
    t0 = time.perf_counter()
    ix.add_synthetic(n, 12345)
    ix.build()
    out["build_s"] = time.perf_counter() - t0
    out["index_GB"] = ix.info()["device_bytes"] / 1e9
    ix.set_param("profile", 1)
    q = np.random.default_rng(67890).integers(0, 256, size=(nq, nbytes), dtype=np.uint8)
    if m:
        ids, dists, counts, st = ix.search_mih(q, k, max_radius=max_radius)
        dt = timed(lambda: ix.search_mih(q, k, max_radius=max_radius, with_stats=False))
        kern = ix.get_param("last_kernel_ns") * 1e-9
        out["mih"] = {"queries_per_s_e2e": nq / dt, "kernel_ms": kern * 1e3, "batched": bool(ix.get_param("mih.last_batched")),
                      "mean_radius": float(st["radius"].mean()), "probes_per_query": float(st["probes"].mean()),
                      "candidates_per_query": float(st["candidates"].mean()),
                      "pairs_per_s": float(st["candidates"].sum()) / kern,
                      "per_query_formula_GBps": float(st["probes"].sum() * 8 + st["candidates"].sum() * nbytes) / kern / 1e9}
        if max_radius < 0 and check:
            lid, ld, lc = ix.search_linear(q[:check], k)
            out["mih"]["equals_linear_scan"] = bool(np.array_equal(lid, ids[:check]) and np.array_equal(ld, dists[:check]))
    out["scan"] = {}
    for B in scan_batches:
        qq = q[:B] if B <= nq else np.random.default_rng(1).integers(0, 256, size=(B, nbytes), dtype=np.uint8)
        dt = timed(lambda: ix.search_linear(qq, k), reps=2)
        kern = ix.get_param("last_kernel_ns") * 1e-9
        out["scan"]["B=%d" % B] = {"queries_per_s_e2e": len(qq) / dt, "kernel_ms": kern * 1e3,
                                  "hbm_GBps": n * nbytes / kern / 1e9, "frac_of_measured_peak": n * nbytes / kern / 1e9 / PEAK,
                                  "pairs_per_s": len(qq) * n / kern}
    ix.close()
    return out


def scan_sweep(n=1_000_000_000, bits=64, k=100):
    """Config C4: brute-force scan of 1 B x 64-bit codes, batch sweep 1..4096, against the roofline
    max(code bytes / HBM peak, tests / POPC peak)."""
    nbytes = bits // 8
    ix = capi.Index(bits, 0)
    ix.add_synthetic(n, 12345)
    ix.set_param("profile", 1)
    popc_peak = 15.4 * ix.get_param("num_sms") * 1.965e9
    rows = []
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        q = np.random.default_rng(B).integers(0, 256, size=(B, nbytes), dtype=np.uint8)
        ix.search_linear(q, k)
        t0 = time.perf_counter()
        reps = 3 if B <= 64 else 1
        ks = []
        for _ in range(reps):
            ix.search_linear(q, k)
            ks.append(ix.get_param("last_kernel_ns"))
        dt = (time.perf_counter() - t0) / reps
        kern = float(np.mean(ks)) * 1e-9
        bound = max(n * nbytes / (PEAK * 1e9), B * n / popc_peak)
        rows.append({"batch": B, "queries_per_s_e2e": B / dt, "scan_kernel_ms": kern * 1e3, "hbm_GBps": n * nbytes / kern / 1e9,
                     "tests_per_s": B * n / kern, "roofline_ms": bound * 1e3, "frac_of_roofline": bound / kern,
                     "bound": "hbm" if n * nbytes / (PEAK * 1e9) > B * n / popc_peak else "popc"})
    ix.close()
    return {"config": "C4 brute-force scan, 1 B x 64-bit codes, k=100, one GPU, batch sweep", "rows": rows}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        print(json.dumps(scan_sweep(), indent=1))
        sys.exit(0)
    res = []
    res.append(run("C1 linear 1M x 64-bit, 1k queries, k=10 (GPU; MIH m=4 beside it)", 1_000_000, 64, 4, 10, 1000, scan_batches=(1000,)))
    for r in (0, 1, 2, 3):
        res.append(run("reference's native shape: 128-bit, m=4 (s=32 bitmap tables), 100M codes, k=100, fixed radius %d" % r, 100_000_000, 128, 4,
                       100, 1024, max_radius=r, scan_batches=(1, 1024) if r == 0 else ()))
    res.append(run("C5 sparse: MIH 256-bit m=8 (s=32 bitmap tables), 60M codes, k=1000, fixed radius 2", 60_000_000, 256, 8, 1000, 64,
                   max_radius=2, scan_batches=()))
    print(json.dumps(res, indent=1))
