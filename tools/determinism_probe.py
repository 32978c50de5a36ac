"""Full-size repeatability and cross-path parity on one GPU: every search path is run several times on the same queries over the
headline index (1 B x 64-bit codes) - all repetitions must be bit-identical (a race shows as a difference) - and the exact paths
are compared with each other on the same queries: MIH (batched) == brute-force scan through the verify kernel == scan through the
tensor-core kernel == TMA-ring scan, ids, distances and order.     python tools/determinism_probe.py [n=1000000000] [reps=4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from verticut_b200 import capi

opt = {"n": 1_000_000_000, "reps": 4, "bits": 64, "m": 4, "k": 100}
for a in sys.argv[1:]:
    name, v = a.split("=")
    opt[name] = int(v)
n, reps, bits, m, k = opt["n"], opt["reps"], opt["bits"], opt["m"], opt["k"]
ix = capi.Index(bits, m)
ix.add_synthetic(n, 12345)
ix.build()
rng = np.random.default_rng(3)
Q = rng.integers(0, 256, size=(16384, bits // 8), dtype=np.uint8)


def keys(res):
    return (res[1].astype(np.uint64) << np.uint64(32)) | res[0].astype(np.uint64)


def repeat(label, fn):
    first = keys(fn())
    diff = 0
    for _ in range(reps - 1):
        diff += int(not np.array_equal(keys(fn()), first))
    print({"path": label, "reps": reps, "repetitions_that_differ": diff})
    return first, diff


bad = 0
mih, d = repeat("mih exact, batch 16384", lambda: ix.search_mih(Q, k, with_stats=False))
bad += d
mih4k, d = repeat("mih exact, batch 4096", lambda: ix.search_mih(Q[:4096], k, with_stats=False))
bad += d
eq = bool(np.array_equal(mih4k, mih[:4096]))
print({"check": "batch 4096 == first 4096 of batch 16384 (P5: independent of the batch size)", "equal": eq})
bad += int(not eq)
_, d = repeat("mih approximate, batch 16384", lambda: ix.search_mih(Q, k, approximate=True, with_stats=False))
bad += d
for B, what in [(1, "TMA-ring kernel"), (2, "verify kernel, scan mode"), (4, "verify kernel, scan mode"), (64, "verify kernel, scan mode"),
                (512, "tensor-core kernel")]:
    lin, d = repeat("linear scan, batch %d (%s; batched=%d)" % (B, what, -1), lambda: ix.search_linear(Q[:B], k))
    bad += d
    eq = bool(np.array_equal(lin, mih[:B]))
    print({"check": "scan batch %d == MIH on the same queries (P2)" % B, "equal": eq, "scan_batched": ix.get_param("scan.last_batched"),
           "scan_tc": ix.get_param("scan.last_tc")})
    bad += int(not eq)
ix.set_param("scan.batched", 0)
ring, d = repeat("linear scan, batch 8, TMA-ring kernel forced", lambda: ix.search_linear(Q[:8], k))
bad += d
eq = bool(np.array_equal(ring, mih[:8]))
print({"check": "ring scan batch 8 == MIH (P2)", "equal": eq})
bad += int(not eq)
print({"n": n, "bits": bits, "m": m, "k": k, "total_bad": bad, "lib": os.environ.get("VC_GPU_LIB", "in-tree")})
