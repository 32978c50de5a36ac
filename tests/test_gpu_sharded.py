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
