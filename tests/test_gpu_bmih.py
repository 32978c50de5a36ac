"""The bucket-stationary batched MIH path (bmih.cuh) must give exactly the answers and statistics of the
per-query kernel: the whole MIH parity suite is re-run with the batched path forced on, plus batch-sized cases."""
import numpy as np
import pytest

import test_gpu_mih as base
from verticut_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_batched():
    base.BATCHED = 1
    yield
    base.BATCHED = 0


@pytest.mark.parametrize("n,bits,m,nq,k", [
    (20_000, 64, 4, 16, 10), (200_000, 64, 4, 8, 100), (50_000, 128, 8, 6, 100), (20_000, 64, 8, 5, 10), (5_000, 256, 16, 3, 50),
])
def test_batched_exact_equals_linear_scan(oracle, n, bits, m, nq, k):
    base._check_exact(oracle, n, bits, m, nq, k)


def test_batched_first_id_small_and_large_k(oracle):
    base._check_exact(oracle, 30_000, 64, 4, 5, 100, first_id=3_000_000_000)
    base._check_exact(oracle, 50, 64, 4, 3, 100)
    base._check_exact(oracle, 20_000, 64, 4, 2, 1000)


def test_batched_heavy_ties_fall_back_exactly(oracle):
    base.test_mih_heavy_ties(oracle)


@pytest.mark.parametrize("bits,m,r", [(64, 4, 0), (64, 4, 2), (128, 8, 1), (256, 16, 1)])
def test_batched_fixed_radius(oracle, bits, m, r):
    base.test_mih_fixed_radius(oracle, bits, m, r)


def test_batched_large_batch_matches_per_query_kernel(oracle):
    # a real batch: 300 queries over 2M codes (buckets of ~30 codes x 4 tables), both paths, all statistics
    n, nq, k = 2_000_000, 300, 100
    ix = capi.Index(64, 4)
    ix.add_synthetic(n, 12345)
    ix.build()
    queries = oracle.synth_codes(67890, 0, nq, 8)
    ix.set_param("mih.batched", 0)
    a = ix.search_mih(queries, k)
    ix.set_param("mih.batched", 1)
    b = ix.search_mih(queries, k)
    assert ix.get_param("mih.last_batched") == 1
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(x, y)
    for f in ("radius", "n_results"):
        np.testing.assert_array_equal(a[3][f], b[3][f])
    for steps in (0, 1):                      # same stop rhythm -> same probe / candidate counts on both paths
        ix.set_param("mih.table_steps", steps)
        ix.set_param("mih.batched", 0)
        a2 = ix.search_mih(queries[:64], k)
        ix.set_param("mih.batched", 1)
        b2 = ix.search_mih(queries[:64], k)
        for f in ("radius", "n_results", "probes", "candidates"):
            np.testing.assert_array_equal(a2[3][f], b2[3][f])
        np.testing.assert_array_equal(a2[0], b2[0])
    ix.set_param("mih.table_steps", -1)
    lid, ld, lc = ix.search_linear(queries[:16], k)
    np.testing.assert_array_equal(b[0][:16], lid)
    ix.close()


def test_batched_state_is_clean_between_calls(oracle):
    # the per-query device state is reused across calls and batch sizes; answers must not depend on history
    n, k = 300_000, 50
    ix = capi.Index(64, 4)
    ix.add_synthetic(n, 12345)
    ix.build()
    ix.set_param("mih.batched", 1)
    codes = oracle.synth_codes(12345, 0, n, 8)
    for nq, seed in ((200, 1), (7, 2), (200, 3), (1, 4)):
        q = oracle.synth_codes(seed, 0, nq, 8)
        ids, dists, counts, _ = ix.search_mih(q, k)
        oid, od, oc = oracle.linear_search(codes, q[:5], k)
        np.testing.assert_array_equal(ids[:5], oid)
        np.testing.assert_array_equal(dists[:5], od)
    ix.close()


def test_batched_large_k_stays_on_the_batched_path(oracle):
    """k = 1000 (config C5): the candidate buffers scale with k, so no query overflows into the per-query kernel."""
    n, nq, k = 400_000, 24, 1000
    codes = oracle.synth_codes(12345, 0, n, 8)
    ix = capi.Index(64, 4)
    ix.add(codes)
    ix.build()
    ix.set_param("mih.batched", 1)
    q = oracle.synth_codes(67890, 0, nq, 8)
    ids, dists, counts, _ = ix.search_mih(q, k)
    assert ix.get_param("mih.last_batched") == 1 and ix.get_param("mih.last_redo") == 0
    oid, od, oc = oracle.linear_search(codes, q, k)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    ix.close()


@pytest.mark.parametrize("bits,m,r,k", [(256, 16, 2, 1000), (128, 8, 2, 300)])
def test_batched_fixed_radius_large_k(oracle, bits, m, r, k):
    """Config C5's shape in small: fixed radius, k = 1000, 256-bit codes - top-k among the candidates found, vs the oracle."""
    n, nq = 60_000, 6
    codes = oracle.synth_codes(12345, 0, n, bits // 8)
    ix = capi.Index(bits, m)
    ix.add(codes)
    ix.build()
    ix.set_param("mih.batched", 1)
    queries = codes[:nq].copy()
    queries[:, 0] ^= 1
    ids, dists, counts, stats = ix.search_mih(queries, k, max_radius=r)
    assert ix.get_param("mih.last_batched") == 1 and ix.get_param("mih.last_redo") == 0
    oid, od, oc, _ = oracle.Index(codes, m).search(queries, k, max_radius=r)
    np.testing.assert_array_equal(counts, oc)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    ix.close()


@pytest.mark.parametrize("cap", [64, 150, 700, 4096])      # 4096 (the default): long buffers, the settle kernel selects before it sorts
def test_batched_overflow_redoes_only_the_overflowed_queries(oracle, cap):
    """A candidate buffer that overflows (forced here with a tiny mih.cap) takes that query - not the whole batch - through
    the per-query kernel; the answers stay exact either way."""
    n, nq, k = 300_000, 40, 50
    codes = oracle.synth_codes(12345, 0, n, 8)
    ix = capi.Index(64, 4)
    ix.add(codes)
    ix.build()
    ix.set_param("mih.batched", 1)
    ix.set_param("mih.cap", cap)
    ix.set_param("mih.boot_sample", 64)                      # loose first thresholds: plenty of appends
    q = oracle.synth_codes(67890, 0, nq, 8)
    q[:8] = codes[:8]                                        # and a few queries that sit on database codes
    ids, dists, counts, st = ix.search_mih(q, k)
    redo = ix.get_param("mih.last_redo")
    assert ix.get_param("mih.last_batched") == 1
    if cap <= 150:
        assert 0 < redo <= nq
    oid, od, oc = oracle.linear_search(codes, q, k)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    np.testing.assert_array_equal(counts, oc)
    assert (st["n_results"] == k).all()
    ix.close()


def test_batched_settle_folds_more_candidates_than_fit_in_shared_memory(oracle):
    """k = 2000 with a 16 K-entry buffer: the settle kernel sorts 4096 entries at a time and must fold the rest exactly."""
    n, nq, k = 200_000, 5, 2000
    codes = oracle.synth_codes(12345, 0, n, 8)
    ix = capi.Index(64, 4)
    ix.add(codes)
    ix.build()
    ix.set_param("mih.batched", 1)
    ix.set_param("mih.boot_sample", 64)
    q = oracle.synth_codes(67890, 0, nq, 8)
    ids, dists, counts, _ = ix.search_mih(q, k)
    assert ix.get_param("mih.last_batched") == 1
    oid, od, oc = oracle.linear_search(codes, q, k)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    ix.close()


@pytest.mark.parametrize("knobs", [{"mih.r0_first": 1}, {"mih.split_r0": 1}, {"mih.r0_first": 1, "mih.prefilter": 1}, {"mih.prefilter": 0}],
                         ids=["r0_first", "split_r0", "r0_first+lower-bound", "exact-filter"])
@pytest.mark.parametrize("n,bits,m,nq,k", [(300_000, 64, 4, 40, 100), (120_000, 128, 8, 12, 100), (40_000, 256, 16, 6, 50)])
def test_batched_first_step_variants_are_exact(oracle, knobs, n, bits, m, nq, k):
    """How the first search step is organised (the queries' own buckets verified first / radius 0 as a step of its own / which
    filter) changes thresholds and order of discovery, never the answer."""
    base.EXTRA_PARAMS = dict(knobs)
    try:
        base._check_exact(oracle, n, bits, m, nq, k)
    finally:
        base.EXTRA_PARAMS = {}


@pytest.mark.parametrize("bits,m,k,how", [
    (64, 8, 3, "longest bucket alone holds 20 k candidates: nothing is counted"),
    (64, 8, 10, "sum of the buckets decides only after counting"),
    (64, 8, 62, "20 k = 1240 ~ the distinct candidates of radius 0: about half of the queries go on to radius 1"),
    (128, 16, 5, "128-bit codes"),
    (64, 4, 10, "short buckets: every query is answered by the per-query kernel"),
])
def test_batched_approximate(oracle, bits, m, k, how):
    """Approximate mode (search_worker.cc:93-157) through the batched path: radius 0 batched, the queries that need more one by
    one; ids, distances, stop radius and the distinct-candidate count equal the oracle's."""
    n, nq = 40_000, 48
    codes, ix = base._mk(oracle, n, bits, m)
    queries = oracle.synth_codes(67890, 0, nq, bits // 8)
    oix = oracle.Index(codes, m)
    oid, od, oc, ost = oix.search(queries, k, approximate=True)
    for with_stats in (True, False):
        ids, dists, counts, stats = ix.search_mih(queries, k, approximate=True, with_stats=with_stats)
        assert ix.get_param("mih.last_batched") == 1
        np.testing.assert_array_equal(ids, oid)
        np.testing.assert_array_equal(dists, od)
        np.testing.assert_array_equal(counts, oc)
        if with_stats:
            np.testing.assert_array_equal(stats["radius"], [s["radius"] for s in ost])
            np.testing.assert_array_equal(stats["unique"], [s["unique"] for s in ost])
    n_beyond_radius0 = sum(1 for s in ost if s["radius"] > 0)
    assert ix.get_param("mih.last_redo") == n_beyond_radius0
    if "half" in how:
        assert 0 < n_beyond_radius0 < nq
    ix.close()


def test_batched_approximate_equals_per_query_kernel_on_long_buckets(oracle):
    n, nq, k = 3_000_000, 200, 20
    ix = capi.Index(64, 8)                   # m = 8: 256 buckets of ~11 700 codes per table - radius 0 is enough for every query
    ix.add_synthetic(n, 777)
    ix.build()
    queries = oracle.synth_codes(4242, 0, nq, 8)
    ix.set_param("mih.batched", 0)
    a = ix.search_mih(queries, k, approximate=True)
    assert ix.get_param("mih.last_batched") == 0
    ix.set_param("mih.batched", -1)
    b = ix.search_mih(queries, k, approximate=True)
    assert ix.get_param("mih.last_batched") == 1 and ix.get_param("mih.last_redo") == 0
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(x, y)
    for f in ("radius", "n_results", "probes", "candidates", "unique"):
        np.testing.assert_array_equal(a[3][f], b[3][f])
    ix.close()


def _kth_and_answers(oracle, codes, q, k):
    oid, od, oc = oracle.linear_search(codes, q, k)
    return oid, od, od[:, k - 1]


@pytest.mark.parametrize("bits,m,n,nq,k", [(64, 4, 400_000, 48, 50), (128, 8, 200_000, 24, 100)])
def test_speculative_threshold_misses_are_redone_exactly(oracle, bits, m, n, nq, k):
    """Speculative thresholds (bmih_decide_kernel): the starting threshold of every query is capped by a guess of its k-th
    distance.  Whatever the guess - far too small, too small for some queries, exactly the largest k-th distance - the answers
    are the oracle's: a query the guess was too small for is detected (fewer than k codes within it once everything within it has
    been enumerated) and redone by the per-query kernel."""
    codes = oracle.synth_codes(12345, 0, n, bits // 8)
    q = oracle.synth_codes(67890, 0, nq, bits // 8)
    q[:4] = codes[:4]
    oid, od, kth = _kth_and_answers(oracle, codes, q, k)
    ix = capi.Index(bits, m)
    ix.add(codes)
    ix.build()
    ix.set_param("mih.batched", 1)
    for guess in (1, int(np.median(kth)) - 1, int(np.median(kth)), int(kth.max()) - 1, int(kth.max()), int(kth.max()) + 3):
        ix.set_param("mih.spec_tau", guess)
        ids, dists, counts, st = ix.search_mih(q, k)
        assert ix.get_param("mih.last_batched") == 1 and ix.get_param("mih.last_spec_tau") == guess
        misses = int((kth > guess).sum())
        assert ix.get_param("mih.last_spec_fail") == misses, (guess, misses)
        assert ix.get_param("mih.last_redo") == misses
        np.testing.assert_array_equal(dists, od, err_msg="guess %d" % guess)
        np.testing.assert_array_equal(ids, oid, err_msg="guess %d" % guess)
        assert (st["n_results"] == k).all()
    ix.close()


def test_speculative_threshold_is_learned_and_leaves_the_statistics_alone(oracle):
    """The guess is learned from the previous batch on the same index (its largest k-th distance); a batch with a miss is
    followed by one without a guess.  For queries the guess holds for, radius / probes / candidates are those of a search without
    it (fixed step rhythm)."""
    n, nq, k = 1_000_000, 64, 100
    codes = oracle.synth_codes(12345, 0, n, 8)
    qa = oracle.synth_codes(67890, 0, nq, 8)
    qb = oracle.synth_codes(24680, 0, nq, 8)
    ix = capi.Index(64, 4)
    ix.add(codes)
    ix.build()
    ix.set_param("mih.batched", 1)
    ix.set_param("mih.table_steps", 1)
    ix.set_param("mih.speculate", 0)
    plain = ix.search_mih(qb, k)
    assert ix.get_param("mih.last_spec_tau") == -1
    ix.set_param("mih.speculate", 2)                                     # 1 (the default) speculates on batches of >= 1024 queries only
    a = ix.search_mih(qa, k)
    assert ix.get_param("mih.last_spec_tau") == -1                      # nothing learned yet
    kth_a = int(a[1][:, k - 1].max())
    b = ix.search_mih(qb, k)
    assert ix.get_param("mih.last_spec_tau") == kth_a
    misses = int((plain[1][:, k - 1] > kth_a).sum())
    assert ix.get_param("mih.last_spec_fail") == misses
    np.testing.assert_array_equal(b[0], plain[0])
    np.testing.assert_array_equal(b[1], plain[1])
    held = plain[1][:, k - 1] <= kth_a
    for f in ("radius", "n_results", "probes", "candidates"):
        np.testing.assert_array_equal(b[3][f][held], plain[3][f][held])
    oid, od, _ = oracle.linear_search(codes, qb, k)
    np.testing.assert_array_equal(b[0], oid)
    c = ix.search_mih(qa, k)                                              # after a miss: no guess; else the value learned from qb
    assert ix.get_param("mih.last_spec_tau") == (-1 if misses else int(plain[1][:, k - 1].max()))
    np.testing.assert_array_equal(c[0], a[0])
    ix.close()


def test_speculation_by_default_on_large_batches_stays_exact(oracle):
    """mih.speculate = 1 (the default) learns from and applies to batches of >= 1024 queries: three such batches of different queries
    on one index - the second and third start from the previous one's largest k-th distance, and whichever queries that is too small
    for are redone - all equal to the oracle; a small batch in between neither uses nor changes the guess."""
    n, nq, k = 300_000, 1024, 20
    codes = oracle.synth_codes(12345, 0, n, 8)
    ix = capi.Index(64, 4)
    ix.add(codes)
    ix.build()
    ix.set_param("mih.batched", 1)
    assert ix.get_param("mih.speculate") == 1
    prev_max = -1
    for b, seed in enumerate((67890, 24680, 13579)):
        q = oracle.synth_codes(seed, 0, nq, 8)
        oid, od, _ = oracle.linear_search(codes, q, k)
        ids, dists, counts, st = ix.search_mih(q, k)
        assert ix.get_param("mih.last_batched") == 1
        guess = ix.get_param("mih.last_spec_tau")
        assert guess == prev_max, (b, guess, prev_max)
        misses = int((od[:, k - 1] > guess).sum()) if guess >= 0 else 0
        assert ix.get_param("mih.last_spec_fail") == misses
        np.testing.assert_array_equal(dists, od, err_msg="batch %d" % b)
        np.testing.assert_array_equal(ids, oid, err_msg="batch %d" % b)
        prev_max = -1 if misses else int(od[:, k - 1].max())      # a batch with a miss teaches nothing
        if b == 0:                                                 # a small batch: no guess used, none learned
            small = ix.search_mih(q[:8], k)
            assert ix.get_param("mih.last_spec_tau") == -1
            np.testing.assert_array_equal(small[0], oid[:8])
    ix.close()
