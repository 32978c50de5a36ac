"""P5 on one GPU: the database cut into G id-shards (G index objects), per-shard search, merge kernel ==
the unsharded search == the oracle.  (The NCCL all-gather itself is covered by bench.py --gpus N and by the
gloo test of the plumbing.)"""
import numpy as np
import pytest

from verticut_b200 import capi
from verticut_b200.sharded import shard_range

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("mode", ["linear", "mih"])
def test_sharded_equals_unsharded(oracle, G, mode):
    n, bits, m, nq, k = 60_000, 64, 4, 12, 100
    codes = oracle.synth_codes(12345, 0, n, 8)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    lists = []
    for g in range(G):
        b, e = shard_range(n, G, g)
        ix = capi.Index(bits, m, first_id=b)
        ix.add(codes[b:e])
        ix.build()
        if mode == "linear":
            ids, dists, counts = ix.search_linear(queries, k)
        else:
            ids, dists, counts, _ = ix.search_mih(queries, k)
        keys = (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
        keys[np.arange(k)[None, :] >= counts[:, None]] = np.uint64(capi.EMPTY_KEY)
        lists.append(keys)
        ix.close()
    merged = capi.merge_topk(0, np.stack(lists), k)
    ids, dists, counts = capi.unpack_keys(merged)
    oid, od, oc = oracle.linear_search(codes, queries, k)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    np.testing.assert_array_equal(counts, oc)


def test_global_threshold_hook_keeps_results_exact(oracle):
    """The all-reduce hook (vc_index_set_allreduce) lets a shard stop on the k-th distance of the whole database.
    Emulated on one GPU: shard B is searched first; while shard A is searched, the hook adds B's final distance
    histogram (what rank B would contribute).  A must still return every member of the global top-k it owns."""
    import ctypes as C
    import torch
    n, nq, k = 400_000, 40, 100
    codes = oracle.synth_codes(12345, 0, n, 8)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    half = n // 2
    keys = {}
    ixs = {}
    for name, b, e in (("B", half, n), ("A", 0, half)):
        ix = capi.Index(64, 4, first_id=b)
        ix.add(codes[b:e])
        ix.build()
        ix.set_param("mih.batched", 1)
        ixs[name] = ix
    ids, dists, counts, _ = ixs["B"].search_mih(queries, k)
    HB = 96
    hist_b = np.zeros((nq, HB), dtype=np.int32)
    for q in range(nq):
        for d in dists[q, : counts[q]]:
            hist_b[q, d] += 1
    d_hist_b = torch.from_numpy(hist_b).cuda()
    calls = []

    def fake_allreduce(ptr, n_words, stream):
        class _Raw:
            __cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
        view = torch.as_tensor(_Raw(), device="cuda")
        assert n_words == nq * HB
        view += d_hist_b.view(-1)
        torch.cuda.synchronize()
        calls.append(n_words)

    ixs["A"].set_allreduce(fake_allreduce)
    ida, da, ca, sta = ixs["A"].search_mih(queries, k)
    assert calls, "the hook was never called"
    ixs["A"].set_allreduce(None)
    idl, dl, cl, stl = ixs["A"].search_mih(queries, k)          # local rule, for comparison
    assert sta["probes"].sum() <= stl["probes"].sum()
    pack = lambda i, d, c: np.where(np.arange(k)[None, :] < c[:, None], (d.astype(np.uint64) << np.uint64(32)) | i.astype(np.uint64), np.uint64(capi.EMPTY_KEY))
    merged = capi.merge_topk(0, np.stack([pack(ida, da, ca), pack(ids, dists, counts)]), k)
    mi, md, mc = capi.unpack_keys(merged)
    oid, od, oc = oracle.linear_search(codes, queries, k)
    np.testing.assert_array_equal(mi, oid)
    np.testing.assert_array_equal(md, od)
    for ix in ixs.values():
        ix.close()
