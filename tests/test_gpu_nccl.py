"""The id-sharded search across REAL GPUs, one process per GPU, in its three forms: peer-memory exchange over NVLink (the search
kernels store their histogram rows and top-k rows into the other GPU's window, verticut_b200/csrc/xchg.cuh - the default),
ncclAllReduce issued from C + NCCL all-gather, and torch.distributed.all_reduce through the Python callback - merged answers vs
the CPU oracle, bit-exact, on every rank.  This is the path that replaces src/mpi_coordinator.cc:34-69 (gather_vectors / bcast) and the per-radius exchange of
src/search_worker.cc:177,207.  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped on a one-GPU box, where
tests/test_gpu_sharded.py covers the same kernels with the exchange emulated between two host threads."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        from verticut_b200 import capi
        return capi.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, cfg, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["VC_NCCL_DIRECT"] = "1" if cfg["direct"] else "0"
    os.environ["VC_XCHG"] = "1" if cfg["xchg"] else "0"
    import torch
    import torch.distributed as dist
    from verticut_b200 import capi
    from verticut_b200.sharded import ShardedSearcher, shard_interleaved, shard_range
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    bits, m, n, nq, k = cfg["bits"], cfg["m"], cfg["n"], cfg["nq"], cfg["k"]
    if cfg["interleaved"]:
        first, stride, cnt = shard_interleaved(n, world, rank)
    else:
        b, e = shard_range(n, world, rank)
        first, stride, cnt = b, 1, e - b
    ix = capi.Index(bits, m, device=rank, first_id=first)
    ix.set_param("id_stride", stride)
    ix.add_synthetic(cnt, 12345)
    ix.build()
    ix.set_param("mih.batched", 1)
    s = ShardedSearcher(ix)
    assert s.peer_windows == bool(cfg["xchg"]), s.exchange
    if not cfg["xchg"]:
        assert ("ncclAllReduce" in s.exchange) == bool(cfg["direct"]), s.exchange
    q = torch.from_numpy(cfg["queries"]).to(dev)
    out = {}
    for mode in ("mih", "linear"):
        for rep in range(2):                              # twice: the second batch runs on reused buffers / a warm communicator
            res = s.search(q, k, mode=mode, max_radius=cfg["r"] if mode == "mih" else -1)
            torch.cuda.synchronize()
        out[mode] = res.cpu().numpy().view(np.uint64).copy()
    out["batched"] = ix.get_param("mih.last_batched")
    out["xchg_last"] = ix.get_param("xchg.last")
    out["levels"] = ix.get_param("mih.last_levels")
    ret[rank] = out
    dist.barrier()
    s.close()
    ix.close()
    dist.destroy_process_group()


CASES = [
    dict(bits=64, m=4, n=3_000_001, nq=96, k=100, r=-1, interleaved=True, direct=True, xchg=True),
    dict(bits=64, m=4, n=3_000_001, nq=96, k=100, r=-1, interleaved=True, direct=True, xchg=False),
    dict(bits=64, m=4, n=3_000_001, nq=96, k=100, r=-1, interleaved=True, direct=False, xchg=False),
    dict(bits=128, m=8, n=1_500_000, nq=40, k=100, r=-1, interleaved=True, direct=True, xchg=True),
    dict(bits=64, m=4, n=2_000_000, nq=33, k=11, r=-1, interleaved=False, direct=True, xchg=True),
    dict(bits=256, m=16, n=600_000, nq=16, k=1000, r=2, interleaved=True, direct=True, xchg=True),
    dict(bits=256, m=16, n=600_000, nq=16, k=1000, r=2, interleaved=True, direct=True, xchg=False),
]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("cfg", CASES, ids=lambda c: "%dbit-m%d-k%d-r%d-%s-%s" % (c["bits"], c["m"], c["k"], c["r"], "il" if c["interleaved"] else "range",
                                                                    "peer" if c["xchg"] else ("nccl" if c["direct"] else "torch")))
def test_two_ranks_over_nccl_equal_the_oracle(oracle, cfg):
    import torch.multiprocessing as mp
    world = 2
    cfg = dict(cfg)
    nbytes = cfg["bits"] // 8
    queries = oracle.synth_codes(67890, 0, cfg["nq"], nbytes)
    if cfg["r"] >= 0:                                      # fixed radius: near-duplicates of database codes, so that small radii find something
        queries = oracle.synth_codes(12345, 0, cfg["nq"], nbytes)
        queries[:, 0] ^= 1
    cfg["queries"] = queries
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), cfg, ret), nprocs=world, join=True)
    want_scan = oracle.scan_synth(12345, 0, 1, cfg["n"], nbytes, queries, cfg["k"], n_procs=4)
    want_mih = want_scan if cfg["r"] < 0 else oracle.scan_synth(12345, 0, 1, cfg["n"], nbytes, queries, cfg["k"], m=cfg["m"],
                                                                 max_radius=cfg["r"], n_procs=4)
    for r in range(world):
        assert ret[r]["batched"] == 1
        assert (ret[r]["xchg_last"] > 0) == bool(cfg["xchg"])
        np.testing.assert_array_equal(ret[r]["linear"], want_scan)      # P1 / P5: independent of the sharding, same on every rank
        np.testing.assert_array_equal(ret[r]["mih"], want_mih)          # P2
