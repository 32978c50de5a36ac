"""Repeats every search path on small random indexes and compares EVERY repetition with numpy's brute force - a search for races
(the parity tests run each case once; the stage-release race of the ring kernel showed in one call out of four).
    python tools/stress_probe.py [reps=15]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from verticut_b200 import capi

reps = 15
for a in sys.argv[1:]:
    name, v = a.split("=")
    if name == "reps":
        reps = int(v)


def truth(w, qw, k, first_id, stride):
    out = np.empty((qw.shape[0], k), np.uint64)
    ids = first_id + stride * np.arange(w.shape[0], dtype=np.uint64)
    for i in range(qw.shape[0]):
        d = np.bitwise_count(w ^ qw[i]).sum(axis=1).astype(np.uint64)
        out[i] = np.sort((d << np.uint64(32)) | ids)[:k]
    return out


CASES = [
    # bits, m, n, nq, k, knobs
    (64, 4, 2_000_000, 96, 100, {}),
    (64, 4, 2_000_000, 96, 100, {"mih.batched": 1}),
    (64, 4, 2_000_000, 33, 11, {"mih.batched": 1, "scan.batched": 0}),
    (64, 4, 500_000, 300, 100, {"mih.batched": 1, "mih.cap": 256}),
    (64, 4, 3_000_000, 16, 1000, {"mih.batched": 1}),
    (128, 8, 1_000_000, 40, 100, {"mih.batched": 1}),
    (128, 8, 1_000_000, 12, 1000, {}),
    (256, 16, 600_000, 16, 1000, {"mih.batched": 1}),
    (256, 16, 600_000, 16, 100, {}),
    (256, 16, 600_000, 5, 100, {"scan.batched": 0, "mih.batched": 0}),
    (64, 4, 1_000_000, 7, 100, {"scan.batched": 0, "scan.qt": 3}),
    (128, 8, 700_000, 24, 500, {"scan.batched": 0}),
]
rng = np.random.default_rng(21)
total_bad = 0
for bits, m, n, nq, k, knobs in CASES:
    nbytes = bits // 8
    codes = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    queries = codes[rng.integers(0, n, size=nq)].copy()
    queries[:, 0] ^= 1
    queries[nq // 2:] = rng.integers(0, 256, size=(nq - nq // 2, nbytes), dtype=np.uint8)      # half near-duplicates, half fresh codes
    first_id, stride = 3, 2
    ix = capi.Index(bits, m, first_id=first_id)
    ix.set_param("id_stride", stride)
    ix.add(codes)
    ix.build()
    for name, v in knobs.items():
        ix.set_param(name, v)
    want = truth(codes.view(np.uint64), queries.view(np.uint64), k, first_id, stride)
    bad = {"linear": 0, "mih": 0}
    for rep in range(reps):
        ids, dists, _ = ix.search_linear(queries, k)
        if not np.array_equal((dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64), want):
            bad["linear"] += 1
        ids, dists, _, _ = ix.search_mih(queries, k)
        if not np.array_equal((dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64), want):
            bad["mih"] += 1
    total_bad += bad["linear"] + bad["mih"]
    print({"bits": bits, "m": m, "n": n, "nq": nq, "k": k, "knobs": knobs, "reps": reps, "bad": bad,
           "mih_batched": ix.get_param("mih.last_batched"), "scan_batched": ix.get_param("scan.last_batched"), "redo": ix.get_param("mih.last_redo")})
    ix.close()
print({"total_bad": total_bad, "lib": os.environ.get("VC_GPU_LIB", "in-tree")})
