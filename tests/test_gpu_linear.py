"""GPU parity, linear scan (K6/K5/K7): vc_search_linear through the C ABI vs the CPU oracle.
Contract P1: identical ids, distances and order (ascending (dist, id))."""
import numpy as np
import pytest

from verticut_b200 import capi

pytestmark = pytest.mark.gpu


EXTRA_PARAMS = {}      # module switch (test_gpu_tc.py): knobs for every index built here
SCAN_BATCHED = -1     # module switch (test_gpu_scan_batched.py): 1 forces the verify-kernel scan path, 0 the TMA-ring kernel


def _run(oracle, n, bits, nq, k, first_id=0, params=None, seed=12345):
    nbytes = bits // 8
    codes = oracle.synth_codes(seed, first_id, n, nbytes)
    queries = oracle.synth_codes(67890, 0, nq, nbytes)
    ix = capi.Index(bits, 0, first_id=first_id)
    ix.set_param("scan.batched", SCAN_BATCHED)
    for name, v in {**EXTRA_PARAMS, **(params or {})}.items():
        ix.set_param(name, v)
    ix.add(codes)
    ids, dists, counts = ix.search_linear(queries, k)
    oid, od, oc = oracle.linear_search(codes, queries, k, first_id=first_id)
    ix.close()
    np.testing.assert_array_equal(counts, oc)
    np.testing.assert_array_equal(dists, od)
    np.testing.assert_array_equal(ids, oid)


@pytest.mark.parametrize("bits", [64, 128, 256])
@pytest.mark.parametrize("n,nq,k", [(1000, 3, 10), (5000, 17, 100), (70001, 40, 10), (300000, 5, 100)])
def test_linear_matches_oracle(oracle, bits, n, nq, k):
    _run(oracle, n, bits, nq, k)


def test_linear_c1_shape(oracle):
    # BASELINE config C1: 1M x 64-bit, k = 10 (query count trimmed to keep the CPU oracle in seconds)
    _run(oracle, 1_000_000, 64, 64, 10)


@pytest.mark.parametrize("pf", [0, 1])
def test_linear_prefilter_variants_agree(oracle, pf):
    _run(oracle, 200_000, 64, 33, 100, params={"scan.prefilter": pf})
    _run(oracle, 100_000, 128, 9, 100, params={"scan.prefilter": pf})


def test_linear_first_id_and_small_tiles(oracle):
    _run(oracle, 50_000, 64, 7, 100, first_id=4_000_000_000 - 50_000)
    _run(oracle, 50_000, 64, 70, 10, params={"scan.qt": 3, "scan.waves": 3})


def test_linear_fewer_codes_than_k(oracle):
    _run(oracle, 7, 64, 3, 10)
    _run(oracle, 1, 128, 2, 100)


def test_linear_heavy_ties(oracle):
    # every code identical -> every distance ties; canonical order must fall back to ids
    n, k = 40_000, 100
    codes = np.tile(np.arange(8, dtype=np.uint8), (n, 1))
    queries = oracle.synth_codes(5, 0, 4, 8)
    ix = capi.Index(64, 0)
    ix.set_param("scan.batched", SCAN_BATCHED)
    ix.add(codes)
    ids, dists, counts = ix.search_linear(queries, k)
    oid, od, oc = oracle.linear_search(codes, queries, k)
    ix.close()
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    assert (ids == np.arange(k, dtype=np.uint32)).all()


def test_linear_large_k(oracle):
    _run(oracle, 30_000, 64, 3, 1000)
    _run(oracle, 30_000, 64, 2, 2048)


def test_ring_scan_repeated_256bit_k1000(oracle):
    """The shape that exposed the stage-release race of the TMA-ring kernel (scan.cuh: a stage handed back to cp.async.bulk while
    shared-memory loads of it were still in flight): 256-bit codes, k = 1000, three queries per CTA, two CTAs per SM.  It failed in
    a quarter of the calls, so the same scan is repeated; every repetition must equal the oracle (ids, distances, order)."""
    n, nq, k, nbytes = 300_000, 16, 1000, 32
    codes = oracle.synth_codes(12345, 0, n, nbytes)
    queries = codes[:nq].copy()
    queries[:, 0] ^= 1
    oid, od, oc = oracle.linear_search(codes, queries, k)
    ix = capi.Index(256, 0)
    ix.add(codes)
    for rep in range(24):
        ids, dists, counts = ix.search_linear(queries, k)
        np.testing.assert_array_equal(dists, od, err_msg="repetition %d" % rep)
        np.testing.assert_array_equal(ids, oid, err_msg="repetition %d" % rep)
    assert ix.get_param("scan.last_batched") == 0
    ix.close()


def test_merge_topk_long_lists_heavy_ties(oracle):
    # 8 and 31 lists whose entries crowd into three distances (the shards' top-k rows): the merge kernel selects by a distance
    # histogram before it sorts (topk_compact_block), the k-th distance has more ties than k
    rng = np.random.default_rng(11)
    for n_lists, nq, k in [(8, 7, 100), (31, 3, 100), (8, 3, 40), (2, 4, 1000), (5, 2, 400)]:
        dist = rng.choice(np.array([11, 12, 13], dtype=np.uint64), size=(n_lists, nq, k), p=[0.02, 0.08, 0.9])
        ids = rng.permutation(n_lists * nq * k).astype(np.uint64).reshape(n_lists, nq, k)
        lists = np.sort((dist << np.uint64(32)) | ids, axis=-1)
        out = capi.merge_topk(0, lists, k)
        for q in range(nq):
            ref = oracle.merge_topk(lists[:, q, :], k)
            np.testing.assert_array_equal(out[q, : ref.size], ref)


def test_merge_topk_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    n_lists, nq, k = 8, 5, 100
    lists = np.sort(rng.integers(0, 1 << 40, size=(n_lists, nq, k), dtype=np.uint64), axis=-1)
    lists[3, :, 50:] = np.uint64(capi.EMPTY_KEY)
    out = capi.merge_topk(0, lists, k)
    for q in range(nq):
        ref = oracle.merge_topk(lists[:, q, :], k)
        np.testing.assert_array_equal(out[q, : ref.size], ref)
