"""The streaming oracle used by bench.py's full-size parity checks (vo_scan_synth: codes generated on the fly, so a 1 B-code
shard needs no host memory) against the pinned oracle functions it restates: the linear scan (src/linear_search.cc:39-64) and
the fixed-radius MIH search (src/search_worker.cc:222-264), both themselves pinned to the reference build in
tests/test_oracle_vs_ref.py."""
import numpy as np
import pytest

EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def _keys(ids, dists, counts, k):
    return np.where(np.arange(k)[None, :] < counts[:, None], (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64), EMPTY)


@pytest.mark.parametrize("nbytes", [8, 16, 32])
@pytest.mark.parametrize("first_id,stride", [(0, 1), (3, 8), (4_000_000_000, 1)])
def test_scan_synth_equals_linear_search(oracle, nbytes, first_id, stride):
    n, nq, k = 30_000, 4, 100
    full = oracle.synth_codes(12345, first_id, n * stride, nbytes)[::stride]
    q = oracle.synth_codes(67890, 0, nq, nbytes)
    ids, dists, counts = oracle.linear_search(full, q, k)
    want = np.where(np.arange(k)[None, :] < counts[:, None],
                    (dists.astype(np.uint64) << np.uint64(32)) | (np.uint64(first_id) + np.uint64(stride) * ids.astype(np.uint64)), EMPTY)
    for procs in (1, 3):
        got = oracle.scan_synth(12345, first_id, stride, n, nbytes, q, k, n_procs=procs)
        np.testing.assert_array_equal(got, want)


def test_scan_synth_fewer_codes_than_k(oracle):
    q = oracle.synth_codes(67890, 0, 2, 8)
    got = oracle.scan_synth(12345, 0, 1, 7, 8, q, 10)
    assert (got[:, 7:] == EMPTY).all() and (got[:, :7] != EMPTY).all()
    assert (np.diff(got[:, :7].astype(np.uint64)) > 0).all()


@pytest.mark.parametrize("nbytes,m,r", [(8, 4, 0), (8, 4, 2), (16, 8, 1), (16, 4, 3), (32, 16, 2), (32, 8, 1)])
def test_scan_synth_fixed_radius_equals_mih_search(oracle, nbytes, m, r):
    n, nq, k = 40_000, 4, 200
    codes = oracle.synth_codes(12345, 0, n, nbytes)
    q = codes[:nq].copy()
    q[:, 0] ^= 1
    oid, od, oc, _ = oracle.Index(codes, m).search(q, k, max_radius=r)
    got = oracle.scan_synth(12345, 0, 1, n, nbytes, q, k, m=m, max_radius=r, n_procs=2)
    np.testing.assert_array_equal(got, _keys(oid, od, oc, k))
