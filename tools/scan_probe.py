"""Small scan / MIH run for ncu captures: python tools/scan_probe.py <mode> <n_codes> <batch> [param=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from verticut_b200 import capi
mode, n, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ix = capi.Index(64, 4 if mode == "mih" else 0)
ix.add_synthetic(n, 12345)
ix.build()
ix.set_param("profile", 1)
for a in sys.argv[4:]:
    name, v = a.split("=")
    ix.set_param(name, int(v))
q = np.random.default_rng(1).integers(0, 256, size=(B, 8), dtype=np.uint8)
for _ in range(3):
    if mode == "mih":
        ix.search_mih(q, 100, with_stats=False)
    else:
        ix.search_linear(q, 100)
    ns = ix.get_param("last_kernel_ns")
if mode == "mih":
    import time
    t0 = time.perf_counter()
    for _ in range(3):
        ix.search_mih(q, 100, with_stats=False)
    dt = (time.perf_counter() - t0) / 3
    print(mode, "n", n, "B", B, "kernel_ms", ns / 1e6, "e2e_ms", dt * 1e3, "steps_ms", [ix.get_param("mih.step_ns.%d" % i) / 1e6 for i in range(ix.get_param("mih.last_levels"))], sys.argv[4:])
else:
    print(mode, "n", n, "B", B, "kernel_ms", ns / 1e6, "pairs/s %.3e" % (n * B / (ns * 1e-9)), "GB/s %.1f" % (n * 8 / ns),
          "grid", ix.get_param("scan.last_grid"), "qt", ix.get_param("scan.last_qt"), "occ", ix.get_param("scan.last_occ"), "stages", ix.get_param("scan.last_stages"),
          "smem", ix.get_param("scan.last_smem"), sys.argv[4:])

