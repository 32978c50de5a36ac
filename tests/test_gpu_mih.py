"""GPU parity, index build (K1/K8) and MIH search (K2-K5) through the C ABI vs the CPU oracle.
Contracts: P2 exact MIH == linear scan == canonical oracle, exactly; P4 radius / probe
statistics; build: every bucket has the oracle's members in ascending id order."""
import numpy as np
import pytest

from verticut_b200 import capi

pytestmark = pytest.mark.gpu


BATCHED = 0      # module switch: 1 forces the bucket-stationary batched path wherever it is legal (test_gpu_bmih.py)
EXTRA_PARAMS = {}   # module switch: further knobs for every index built here (test_gpu_tc.py forces the tensor-core verify kernel)


def _mk(oracle, n, bits, m, first_id=0, seed=12345):
    nbytes = bits // 8
    codes = oracle.synth_codes(seed, first_id, n, nbytes)
    ix = capi.Index(bits, m, first_id=first_id)
    ix.set_param("mih.batched", BATCHED)
    for name, v in EXTRA_PARAMS.items():
        ix.set_param(name, v)
    ix.add(codes)
    ix.build()
    return codes, ix


@pytest.mark.parametrize("bits,m", [(64, 4), (64, 8), (128, 8), (64, 2), (128, 4), (256, 16)])
def test_build_buckets_match_oracle(oracle, bits, m):
    n = 20_000
    codes, ix = _mk(oracle, n, bits, m)
    oix = oracle.Index(codes, m)
    sbits = bits // m
    rng = np.random.default_rng(1)
    for t in range(m):
        # buckets of real codes (non-empty) plus random keys (mostly empty when s = 32)
        keys = [oracle.binary_to_int(codes[i, t * sbits // 8:(t + 1) * sbits // 8]) for i in rng.integers(0, n, 6)]
        keys += [int(x) for x in rng.integers(0, 1 << sbits, 6)]
        for key in keys:
            orc, oids = oix.bucket(t, key)
            rc, ids, bcodes = ix.bucket_get(t, key)
            assert rc == orc
            np.testing.assert_array_equal(ids, oids)
            if ids.size:
                assert (np.diff(ids.astype(np.int64)) > 0).all()          # ascending id = insertion order
                np.testing.assert_array_equal(bcodes, codes[ids])
    ix.close()


def test_build_integrity_every_code_in_its_bucket(oracle):
    # the check of the reference's integrity_check.cc:25-35,52-67, through BaseProxy::get
    n, bits, m = 3000, 64, 4
    codes, ix = _mk(oracle, n, bits, m, first_id=100)
    for t in range(m):
        for i in range(0, n, 37):
            key = oracle.binary_to_int(codes[i, 2 * t:2 * t + 2])
            rc, ids, bcodes = ix.bucket_get(t, key)
            assert rc == 0
            j = np.nonzero(ids == 100 + i)[0]
            assert j.size == 1 and (bcodes[j[0]] == codes[i]).all()
    rc, code = ix.code_get(100 + 5)
    assert rc == 0 and (code == codes[5]).all()
    assert ix.code_get(99)[0] == 1 and ix.code_get(100 + n)[0] == 1
    ix.close()


@pytest.mark.parametrize("bits,m", [(64, 4), (128, 8)])
def test_occupancy_bitmap_matches_oracle(oracle, bits, m):
    codes, ix = _mk(oracle, 30_000, bits, m)
    oix = oracle.Index(codes, m)
    for t in (0, m - 1):
        np.testing.assert_array_equal(ix.occupancy_bitmap(t), oix.occupancy_bitmap(t))
    ix.close()


def test_occupancy_bitmap_of_sparse_tables(oracle):
    """s = 32: the 2^32-bit bitmap the device keeps for its sparse tables, exported in the reference's layout (src/bitmap.cc:22-38:
    bit i of 32-bit word i / 32 - the 512 MiB per table generate_bitmap.cc:99 allocates), against the keys of the codes."""
    n, bits, m = 50_000, 64, 2
    codes, ix = _mk(oracle, n, bits, m)
    keys = codes.view(np.uint32).reshape(n, m)               # little-endian: substring t = 32-bit word t (binaryToInt for s = 32)
    for t in range(m):
        words = ix.occupancy_bitmap(t)
        assert words.size == (1 << 32) // 32
        set_bits = np.unique(keys[:, t])
        assert int(np.bitwise_count(words).sum()) == set_bits.size
        assert ((words[set_bits >> 5] >> (set_bits & 31)) & 1).all()
        probe = np.array([0, 1, 2 ** 31, 2 ** 32 - 1], dtype=np.uint64)
        for b in probe:
            assert bool((words[int(b) >> 5] >> np.uint32(int(b) & 31)) & 1) == bool(np.isin(np.uint32(b), set_bits))
    ix.close()


def _check_exact(oracle, n, bits, m, nq, k, first_id=0):
    codes, ix = _mk(oracle, n, bits, m, first_id=first_id)
    queries = oracle.synth_codes(67890, 0, nq, bits // 8)
    lid, ld, lc = oracle.linear_search(codes, queries, k, first_id=first_id)
    oix = oracle.Index(codes, m, first_id=first_id)
    # stop rule tested per radius (0, the reference's rhythm), per table (1), or adaptively (-1, the default):
    # the answers are always the canonical top-k; the statistics follow the oracle run with the same rhythm
    for table_steps, ostop in ((0, oracle.STOP_STRICT_M), (1, oracle.STOP_TABLE_STRICT), (-1, None)):
        ix.set_param("mih.table_steps", table_steps)
        ids, dists, counts, stats = ix.search_mih(queries, k)
        np.testing.assert_array_equal(counts, lc)
        np.testing.assert_array_equal(dists, ld)
        np.testing.assert_array_equal(ids, lid)
        assert (stats["n_results"] == counts).all()
        if ostop is None:
            continue
        oid, od, oc, ost = oix.search(queries, k, order=oracle.ORDER_CANONICAL, stop=ostop)
        np.testing.assert_array_equal(oid, ids)
        for q in range(nq):
            assert stats["radius"][q] == ost[q]["radius"]
            assert stats["candidates"][q] == ost[q]["candidates"]
            if bits // m < 32:
                assert stats["probes"][q] == ost[q]["probes"]
            else:
                assert stats["occupancy_tests"][q] == ost[q]["probes"]
    gid, gd, gc = ix.search_linear(queries, k)
    np.testing.assert_array_equal(gid, lid)
    np.testing.assert_array_equal(gd, ld)
    ix.close()


@pytest.mark.parametrize("n,bits,m,nq,k", [
    (20_000, 64, 4, 16, 10),
    (200_000, 64, 4, 8, 100),
    (50_000, 128, 8, 6, 100),
    (20_000, 64, 8, 5, 10),
    (3_000, 64, 2, 4, 10),       # s = 32: bitmap + rank directory tables
    (5_000, 256, 16, 3, 50),
])
def test_mih_exact_equals_linear_scan(oracle, n, bits, m, nq, k):
    _check_exact(oracle, n, bits, m, nq, k)


def test_mih_first_id_offset(oracle):
    _check_exact(oracle, 30_000, 64, 4, 5, 100, first_id=3_000_000_000)


def test_mih_fewer_codes_than_k(oracle):
    _check_exact(oracle, 50, 64, 4, 3, 100)


def test_mih_large_k(oracle):
    _check_exact(oracle, 20_000, 64, 4, 2, 1000)


def test_mih_heavy_ties(oracle):
    n, k = 5000, 100
    base = oracle.synth_codes(9, 0, 4, 8)
    codes = np.repeat(base, n // 4, axis=0)                # four distinct codes, 1250 copies each
    ix = capi.Index(64, 4)
    ix.set_param("mih.batched", BATCHED)
    for name, v in EXTRA_PARAMS.items():
        ix.set_param(name, v)
    ix.add(codes)
    ix.build()
    queries = oracle.synth_codes(67890, 0, 6, 8)
    queries[0] = base[2]
    ids, dists, counts, _ = ix.search_mih(queries, k)
    lid, ld, lc = oracle.linear_search(codes, queries, k)
    np.testing.assert_array_equal(ids, lid)
    np.testing.assert_array_equal(dists, ld)
    ix.close()


@pytest.mark.parametrize("bits,m,r", [(64, 4, 0), (64, 4, 2), (128, 8, 1), (256, 16, 1), (256, 8, 2)])
def test_mih_fixed_radius(oracle, bits, m, r):
    n, nq, k = 30_000, 4, 100
    codes, ix = _mk(oracle, n, bits, m)
    queries = codes[:nq].copy()
    queries[:, 0] ^= 1                                      # near-duplicates so that small radii find something
    ids, dists, counts, stats = ix.search_mih(queries, k, max_radius=r)
    oix = oracle.Index(codes, m)
    oid, od, oc, ost = oix.search(queries, k, max_radius=r)
    np.testing.assert_array_equal(counts, oc)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    assert (stats["radius"] == r).all()
    ix.close()


@pytest.mark.parametrize("bits,m,k", [(64, 4, 10), (128, 8, 3)])
def test_mih_approximate(oracle, bits, m, k):
    n, nq = 40_000, 6
    codes, ix = _mk(oracle, n, bits, m)
    queries = oracle.synth_codes(67890, 0, nq, bits // 8)
    ids, dists, counts, stats = ix.search_mih(queries, k, approximate=True)
    oix = oracle.Index(codes, m)
    oid, od, oc, ost = oix.search(queries, k, approximate=True)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    for q in range(nq):
        assert stats["radius"][q] == ost[q]["radius"]
        assert stats["unique"][q] == ost[q]["unique"]
    ix.close()


def test_mih_errors():
    ix = capi.Index(64, 4)
    with pytest.raises(capi.VerticutError):
        ix.search_mih(np.zeros((1, 8), np.uint8), 10)      # not built
    ix.build()
    with pytest.raises(capi.VerticutError):
        ix.search_mih(np.zeros((1, 8), np.uint8), 0)
    with pytest.raises(capi.VerticutError):
        capi.Index(64, 3)
    with pytest.raises(capi.VerticutError):
        capi.Index(96, 4)
    ix.close()


def test_host_calls_chunk_long_query_arrays(oracle):
    """vc_search_mih / vc_search_linear take any number of queries and pass them through the device `host.chunk` at a time:
    same answers and statistics as one batch."""
    n, nq, k = 60_000, 70, 10
    codes, ix = _mk(oracle, n, 64, 4)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    whole = ix.search_mih(queries, k)
    whole_lin = ix.search_linear(queries, k)
    ix.set_param("host.chunk", 32)
    parts = ix.search_mih(queries, k)
    parts_lin = ix.search_linear(queries, k)
    for a, b in zip(whole[:3] + whole_lin, parts[:3] + parts_lin):
        np.testing.assert_array_equal(a, b)
    for f in ("radius", "n_results", "probes", "candidates"):
        np.testing.assert_array_equal(whole[3][f], parts[3][f])
    np.testing.assert_array_equal(whole[0], whole_lin[0])
    ix.close()
