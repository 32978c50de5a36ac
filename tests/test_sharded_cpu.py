"""Host-side logic of the id-sharded multi-GPU path on CPU: world_size 2, gloo backend.
The two device steps of ShardedSearcher are replaced by the CPU oracle (the checker standing in for the
kernels, which need a GPU); what is tested is the plumbing the product adds around them: shard ranges,
global ids, the all-gather layout the merge kernel consumes, and that every rank ends with the global top-k."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, nq, k, mode, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import restatement as R
    from verticut_b200.sharded import ShardedSearcher, shard_range

    b, e = shard_range(n_total, world, rank)
    codes = R.synth_codes(12345, b, e - b, 8)            # this rank's shard only
    queries = R.synth_codes(67890, 0, nq, 8)

    class FakeIndex:
        device = 0

    class OracleBacked(ShardedSearcher):
        def local_search(self, d_queries, k_, mode_, approximate, max_radius, out_keys):
            q = d_queries.numpy()
            if mode_ == "linear":
                ids, dists, counts = R.linear_search(codes, q, k_, first_id=b)
            else:
                ids, dists, counts, _ = R.Index(codes, 4, first_id=b).search(q, k_)
            keys = (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
            keys[np.arange(k_)[None, :] >= counts[:, None]] = np.uint64(0xFFFFFFFFFFFFFFFF)
            out_keys.copy_(torch.from_numpy(keys.view(np.int64)))

        def merge(self, gathered, k_, out_keys):
            g = gathered.numpy().view(np.uint64)
            assert g.shape == (world, nq, k_)
            for qi in range(nq):
                m = R.merge_topk(np.ascontiguousarray(g[:, qi, :]), k_)
                row = np.full(k_, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
                row[: m.size] = m
                out_keys[qi].copy_(torch.from_numpy(row.view(np.int64)))

    s = OracleBacked(FakeIndex())
    merged = s.search(torch.from_numpy(queries), k, mode=mode)
    ret[rank] = merged.numpy().view(np.uint64).copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,n_total,k", [("linear", 20001, 10), ("mih", 9000, 100), ("linear", 5, 10)])
def test_sharded_topk_equals_global_topk(oracle, mode, n_total, k):
    world, nq = 2, 5
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, nq, k, mode, ret), nprocs=world, join=True)
    codes = oracle.synth_codes(12345, 0, n_total, 8)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    ids, dists, counts = oracle.linear_search(codes, queries, k)
    want = (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
    want[np.arange(k)[None, :] >= counts[:, None]] = np.uint64(0xFFFFFFFFFFFFFFFF)
    for r in range(world):
        np.testing.assert_array_equal(ret[r], want)      # P5: independent of the sharding, same on every rank


def test_shard_ranges_partition_the_ids():
    from verticut_b200.sharded import gather_layout, shard_range
    for n in (0, 1, 7, 1000, 10**9):
        for g in (1, 2, 3, 4, 8):
            r = [shard_range(n, g, i) for i in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(g - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    assert gather_layout(8, 4096, 100) == (8, 4096, 100)


def test_shard_maps_partition_the_ids():
    sys.path.insert(0, ROOT)
    from verticut_b200.sharded import shard_interleaved, shard_range
    for n in (0, 1, 7, 1000, 1_000_003):
        for G in (1, 2, 3, 8):
            seen_r, seen_i = 0, 0
            prev_end = 0
            for g in range(G):
                b, e = shard_range(n, G, g)
                assert b == prev_end and e >= b
                prev_end = e
                seen_r += e - b
                first, stride, cnt = shard_interleaved(n, G, g)
                assert stride == G and cnt == len(range(g, n, G))
                assert cnt == 0 or (first == g and first + (cnt - 1) * stride < n)
                seen_i += cnt
            assert prev_end == n and seen_r == n and seen_i == n
