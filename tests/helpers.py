"""Shared checks of the parity contract."""
import glob
import os
import zlib

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(path, oracle):
    g = dict(np.load(path))
    n, bits = int(g["n"]), int(g["bits"])
    codes = oracle.synth_codes(int(g["db_seed"]), 0, n, bits // 8)
    assert np.uint32(zlib.crc32(codes.tobytes())) == g["codes_crc"], "synthetic generator drifted from the fixtures"
    return g, codes


def check_p3(canon_ids, canon_dists, canon_counts, ref_ids, ref_dists, ref_counts, true_dist_fn):
    """P3: canonical results vs the reference's own output for one batch.
    - same number of results, same sorted distance multiset;
    - ids identical for every result with dist < d_k;
    - every reference id at dist == d_k really has distance d_k (tie-class membership)."""
    for q in range(canon_ids.shape[0]):
        c = int(canon_counts[q])
        assert c == int(ref_counts[q])
        cd, rd = canon_dists[q, :c].astype(np.int64), ref_dists[q, :c].astype(np.int64)
        np.testing.assert_array_equal(np.sort(cd), np.sort(rd))
        assert (np.diff(rd) <= 0).all(), "reference output must be in descending distance"
        assert (np.diff(cd) >= 0).all(), "canonical output must be in ascending distance"
        if c == 0:
            continue
        dk = cd.max()
        below_c = set(canon_ids[q, :c][cd < dk].tolist())
        below_r = set(ref_ids[q, :c][rd < dk].tolist())
        assert below_c == below_r
        for i in ref_ids[q, :c][rd == dk]:
            assert true_dist_fn(q, int(i)) == dk
