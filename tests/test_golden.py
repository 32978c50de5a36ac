"""Golden vectors generated from the reference itself (tests/golden/make_golden.py).
CPU: the oracle reproduces them.  GPU (-m gpu): the CUDA path satisfies the parity contract against them -
the only check available on the GPU box, where /root/reference does not exist."""
import numpy as np
import pytest

from helpers import check_p3, golden_cases, load_golden


@pytest.mark.parametrize("path", golden_cases())
def test_oracle_reproduces_golden(oracle, path):
    g, codes = load_golden(path, oracle)
    m, k, approx = int(g["m"]), int(g["k"]), bool(g["approximate"])
    ix = oracle.Index(codes, m)
    oid, od, oc, ost = ix.search(g["queries"], k, order=oracle.ORDER_REFERENCE, stop=oracle.STOP_REF4, approximate=approx)
    np.testing.assert_array_equal(oid, g["ref_mih_ids"])
    np.testing.assert_array_equal(od, g["ref_mih_dists"])
    np.testing.assert_array_equal(oc, g["ref_mih_counts"])
    assert [s["radius"] for s in ost] == g["ref_mih_radius"].tolist()
    lid, ld, lc = oracle.linear_search(codes, g["queries"], k, reference_order=True)
    np.testing.assert_array_equal(lid, g["ref_lin_ids"])
    np.testing.assert_array_equal(ld, g["ref_lin_dists"])
    for j, key in enumerate(g["probe_keys"]):
        np.testing.assert_array_equal(ix.bucket(0, int(key))[1], g["bucket%d" % j])


@pytest.mark.gpu
@pytest.mark.parametrize("path", golden_cases())
def test_gpu_against_golden(oracle, path):
    from verticut_b200 import capi
    g, codes = load_golden(path, oracle)
    bits, m, k, approx = int(g["bits"]), int(g["m"]), int(g["k"]), bool(g["approximate"])
    queries = g["queries"]
    ix = capi.Index(bits, m)
    ix.add(codes)
    ix.build()
    true_dist = lambda q, i: oracle.hamming(codes[i], queries[q])
    lid, ld, lc = ix.search_linear(queries, k)
    check_p3(lid, ld, lc, g["ref_lin_ids"], g["ref_lin_dists"], g["ref_lin_counts"], true_dist)
    if not approx:
        ids, dists, counts, st = ix.search_mih(queries, k)
        check_p3(ids, dists, counts, g["ref_mih_ids"], g["ref_mih_dists"], g["ref_mih_counts"], true_dist)
        np.testing.assert_array_equal(ids, lid)
        if m == 4:
            for q in range(len(queries)):        # P4
                assert st["radius"][q] in (g["ref_mih_radius"][q], g["ref_mih_radius"][q] + 1)
    else:
        ids, dists, counts, st = ix.search_mih(queries, k, approximate=True)
        # approximate mode: same stop radius and the same distance multiset as the reference's 20k-heap
        np.testing.assert_array_equal(st["radius"], g["ref_mih_radius"])
        for q in range(len(queries)):
            np.testing.assert_array_equal(np.sort(dists[q]), np.sort(g["ref_mih_dists"][q]))
    for j, key in enumerate(g["probe_keys"]):    # tables as the reference's build-tables leaves them
        rc, bids, bcodes = ix.bucket_get(0, int(key))
        assert rc == 0
        np.testing.assert_array_equal(bids, g["bucket%d" % j])
    ix.close()
