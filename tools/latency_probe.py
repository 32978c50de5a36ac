"""End-to-end latency of one vc_search_mih call (host buffers) vs batch size, per-query kernel vs batched path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from verticut_b200 import capi
n = int(sys.argv[1])
batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4, 16, 64, 256]
ix = capi.Index(64, 4)
ix.add_synthetic(n, 12345)
ix.build()
for B in batches:
    q = np.random.default_rng(B).integers(0, 256, size=(B, 8), dtype=np.uint8)
    row = []
    for mode in (0, 1):
        ix.set_param("mih.batched", mode)
        ix.search_mih(q, 100, with_stats=False)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ix.search_mih(q, 100, with_stats=False)
        row.append((time.perf_counter() - t0) / reps * 1e3)
    print("n=%d B=%d  per-query %.3f ms   batched %.3f ms" % (n, B, row[0], row[1]))
