"""Repeats the local half of the 2-GPU parity case 256-bit / k = 1000 on ONE GPU and compares every repetition with numpy:
two id-interleaved shards of random codes as two indexes on the same device, each scanned (and MIH-searched at a fixed radius),
the two top-k lists merged by the merge kernel.     python tools/flake_probe.py [reps=40] [n=600000] [k=1000] [bits=256] [nq=16]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from verticut_b200 import capi

opt = {"reps": 40, "n": 600000, "k": 1000, "bits": 256, "nq": 16, "m": 16, "r": 2}
knobs = []
for a in sys.argv[1:]:
    name, v = a.split("=")
    if name in opt:
        opt[name] = int(v)
    else:
        knobs.append((name, int(v)))
n, k, bits, nq = opt["n"], opt["k"], opt["bits"], opt["nq"]
nbytes = bits // 8
rng = np.random.default_rng(5)
codes = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
queries = codes[:nq].copy()
queries[:, 0] ^= 1
w = codes.view(np.uint64)
qw = queries.view(np.uint64)
want = np.empty((nq, k), np.uint64)
for i in range(nq):
    d = np.bitwise_count(w ^ qw[i]).sum(axis=1).astype(np.uint64)
    keys = (d << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    want[i] = np.sort(keys)[:k]
shards, own = [], []
for g in range(2):
    ix = capi.Index(bits, opt["m"], first_id=g)
    ix.set_param("id_stride", 2)
    ix.add(np.ascontiguousarray(codes[g::2]))
    ix.build()
    for name, v in knobs:
        ix.set_param(name, v)
    shards.append(ix)
    o = np.empty((nq, k), np.uint64)
    for i in range(nq):
        d = np.bitwise_count(w[g::2] ^ qw[i]).sum(axis=1).astype(np.uint64)
        o[i] = np.sort((d << np.uint64(32)) | np.arange(g, n, 2, dtype=np.uint64))[:k]
    own.append(o)
bad = {"linear": 0, "merged": 0}
for rep in range(opt["reps"]):
    lists = []
    for g, ix in enumerate(shards):
        ids, dists, counts = ix.search_linear(queries, k)
        keys = (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
        if not np.array_equal(keys, own[g]):
            bad["linear"] += 1
            for qi in sorted(set(np.argwhere(keys != own[g])[:, 0].tolist())):
                got, wnt = keys[qi], own[g][qi]
                extra = sorted(set(got.tolist()) - set(wnt.tolist()))
                missing = sorted(set(wnt.tolist()) - set(got.tolist()))
                dup = int(k - len(set(got.tolist())))
                print("rep %d shard %d query %d: first diff at %d, duplicates %d, sorted %s, extra %s, missing %s, k-th want %x" % (
                    rep, g, qi, int(np.argwhere(got != wnt)[0][0]), dup, bool(np.all(got[1:] >= got[:-1])),
                    ["%x" % v for v in extra[:4]], ["%x" % v for v in missing[:4]], wnt[-1]))
                for v in extra[:2]:        # is the extra entry a real code of this shard, and at which true distance?
                    idv = v & 0xFFFFFFFF
                    if idv % 2 == g and idv < n:
                        td = int(np.bitwise_count(w[idv] ^ qw[qi]).sum())
                        print("    extra id %d: claimed distance %d, true distance %d" % (idv, v >> 32, td))
                        if td != (v >> 32):      # whose distance is it?  neighbours in the shard's code order
                            L, ws = (idv - g) // 2, w[g::2]
                            hits = [dl for dl in list(range(-2100, 2101)) if dl and 0 <= L + dl < len(ws) and
                                    int(np.bitwise_count(ws[L + dl] ^ qw[qi]).sum()) == (v >> 32)]
                            other_q = [q2 for q2 in range(nq) if int(np.bitwise_count(w[idv] ^ qw[q2]).sum()) == (v >> 32)]
                            print("      codes at offsets %s have that distance to this query; queries %s have it to this code" % (hits[:12], other_q))
        lists.append(keys)
    merged = capi.merge_topk(0, np.stack(lists), k)
    if not np.array_equal(merged, want):
        bad["merged"] += 1
geo = {nm: shards[0].get_param("scan.last_" + nm) for nm in ("grid", "qt", "slices", "occ", "stages", "batched")}
print({"reps": opt["reps"], "bad": bad, "geometry": geo, "lib": os.environ.get("VC_GPU_LIB", "in-tree"), "opt": opt, "knobs": knobs})
