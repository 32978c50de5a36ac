"""The merge kernel alone (the last step of a sharded search) on one GPU:  python tools/merge_probe.py [lists=8] [nq=16384] [k=100]
G sorted lists of k packed (distance, id) words per query, shaped like the shards' top-k rows of uniform 64-bit codes
(distances 9 .. 15); timed with CUDA events, checked against numpy on the first queries."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from verticut_b200 import capi

opt = {"lists": 8, "nq": 16384, "k": 100, "reps": 20}
for a in sys.argv[1:]:
    name, v = a.split("=")
    opt[name] = int(v)
G, nq, k = opt["lists"], opt["nq"], opt["k"]
rng = np.random.default_rng(7)
dist = rng.choice(np.arange(9, 16), size=(G, nq, k), p=[0.01, 0.02, 0.05, 0.12, 0.2, 0.3, 0.3]).astype(np.uint64)
ids = rng.integers(0, 1 << 30, size=(G, nq, k), dtype=np.uint64) * G + np.arange(G, dtype=np.uint64)[:, None, None]
keys = np.sort((dist << np.uint64(32)) | ids, axis=2)
dev = torch.device("cuda:0")
d_lists = torch.from_numpy(keys.view(np.int64)).to(dev)
d_out = torch.empty((nq, k), dtype=torch.int64, device=dev)
for _ in range(3):
    capi.merge_topk_dev(0, d_lists.data_ptr(), G, nq, k, d_out.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(opt["reps"]):
    capi.merge_topk_dev(0, d_lists.data_ptr(), G, nq, k, d_out.data_ptr())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / opt["reps"]
got = d_out.cpu().numpy().view(np.uint64)
c = min(nq, 256)
want = np.sort(keys[:, :c, :].transpose(1, 0, 2).reshape(c, G * k), axis=1)[:, :k]
print({"merge_ms": round(ms, 4), "lists": G, "nq": nq, "k": k, "equals_numpy": bool(np.array_equal(got[:c], want)),
       "lib": os.environ.get("VC_GPU_LIB", "in-tree")})
