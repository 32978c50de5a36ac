"""Pins the plain-C oracle (oracle/verticut_oracle.c) against the reference ITSELF (oracle/_ref: the reference's
search_worker.cc / linear_search.cc / build_hash_tables.cc / integrity_check.cc compiled unmodified).
CPU only.  Skipped only when neither /root/reference nor a prebuilt oracle/_ref exists."""
import subprocess
import sys

import numpy as np
import pytest

import oracle as oracle_pkg
from oracle import reference as F, restatement as R
from helpers import check_p3

if not F.available():
    try:
        oracle_pkg.build()
    except Exception:
        pass
pytestmark = pytest.mark.skipif(not F.available(), reason="oracle/_ref not built (reference tree absent)")


def _data(n, bits, nq, seed=12345):
    return R.synth_codes(seed, 0, n, bits // 8), R.synth_codes(67890, 0, nq, bits // 8)


@pytest.mark.parametrize("n,bits,m,k,nq,approx", [
    (20000, 64, 4, 10, 10, False),
    (30000, 64, 4, 100, 6, False),
    (3000, 128, 8, 10, 3, False),
    (6000, 64, 4, 10, 8, True),
    (6000, 64, 4, 3, 5, True),
    (40, 64, 4, 100, 3, False),       # fewer codes than k: the loop runs to r = s (search_worker.cc:170)
    (5000, 64, 8, 10, 4, False),      # s = 8
    # 256-bit codes: test_restatement_reproduces_reference_search_at_256_bits (the reference's own build overflows there)
])
def test_restatement_reproduces_reference_mih_exactly(n, bits, m, k, nq, approx):
    codes, queries = _data(n, bits, nq)
    store = F.RefStore(codes, m)
    rid, rd, rc, rrad, rsub = store.mih_search(queries, k, approximate=approx)
    store.close()
    ix = R.Index(codes, m)
    oid, od, oc, ost = ix.search(queries, k, order=R.ORDER_REFERENCE, stop=R.STOP_REF4, approximate=approx)
    np.testing.assert_array_equal(oc, rc)
    np.testing.assert_array_equal(od, rd)        # same distances, same (descending) order
    np.testing.assert_array_equal(oid, rid)      # same ids, including the libstdc++ heap tie order
    for q in range(nq):
        assert ost[q]["radius"] == rrad[q]
        assert ost[q]["probes"] == rsub[q].sum()  # n_sub_reads_ (search_worker.cc:245), all ranks


@pytest.mark.parametrize("n,bits,m,k,nq,approx", [(2000, 256, 16, 10, 2, False), (2000, 256, 16, 5, 2, True), (3000, 256, 16, 100, 2, False)])
def test_restatement_reproduces_reference_search_at_256_bits(n, bits, m, k, nq, approx):
    """Config C5's geometry.  The reference's build cannot take 256-bit codes (`char code[17]`, src/build_hash_tables.cc:41 -
    the unmodified build aborts with "buffer overflow detected"), its search can: the tables are filled by the driver's own
    loop (RefMem, oracle/ref_driver.cc) and the unmodified SearchWorker::find runs over them.  With 256-bit codes the
    hard-coded `top.dist <= radius*4` never fires, so the exact search runs to r = s (src/search_worker.cc:170,204)."""
    codes, queries = _data(n, bits, nq)
    rid, rd, rc, rrad, rsub = F.RefMem(codes, m).mih_search(queries, k, approximate=approx)
    oid, od, oc, ost = R.Index(codes, m).search(queries, k, order=R.ORDER_REFERENCE, stop=R.STOP_REF4, approximate=approx)
    np.testing.assert_array_equal(oc, rc)
    np.testing.assert_array_equal(od, rd)
    np.testing.assert_array_equal(oid, rid)
    for q in range(nq):
        assert ost[q]["radius"] == rrad[q] and ost[q]["probes"] == rsub[q].sum()
    if not approx:
        assert (rrad == bits // m).all()


@pytest.mark.parametrize("n,bits,m,k,nq,max_radius", [
    (3000, 256, 16, 1000, 2, 0), (3000, 256, 16, 1000, 2, 1), (3000, 256, 16, 1000, 2, 2), (3000, 256, 16, 1000, 2, 3),   # config C5's sweep
    (20000, 64, 4, 10, 4, 2), (2000, 128, 8, 100, 3, 1), (500, 64, 8, 10, 3, 8),
])
def test_restatement_reproduces_reference_fixed_radius_search(n, bits, m, k, nq, max_radius):
    """Fixed-radius mode (config C5: radii 0 .. r, no stop rule, the k best of what was found, fewer than k when fewer were
    found).  Reference side: its own search_R_neighbors / enumerate_entry / gather_vectors under a subclass that leaves out the
    stop test (FixedRadiusWorker, oracle/ref_driver.cc)."""
    codes, queries = _data(n, bits, nq)
    rid, rd, rc, rsub = F.RefMem(codes, m).fixed_radius_search(queries, k, max_radius)
    oid, od, oc, ost = R.Index(codes, m).search(queries, k, order=R.ORDER_REFERENCE, stop=R.STOP_REF4, max_radius=max_radius)
    np.testing.assert_array_equal(oc, rc)
    np.testing.assert_array_equal(od, rd)
    np.testing.assert_array_equal(oid, rid)
    for q in range(nq):
        assert ost[q]["probes"] == rsub[q].sum()


@pytest.mark.parametrize("n,bits", [(30000, 64), (8000, 128), (8000, 256)])
def test_restatement_reproduces_reference_linear_scan_exactly(n, bits):
    codes, queries = _data(n, bits, 6)
    store = F.RefStore(codes, 0)
    store.put_main_table()
    rid, rd, rc = store.linear_search(queries, 100)
    store.close()
    oid, od, oc = R.linear_search(codes, queries, 100, reference_order=True)
    np.testing.assert_array_equal(oid, rid)
    np.testing.assert_array_equal(od, rd)
    np.testing.assert_array_equal(oc, rc)
    mid, md, mc = F.RefMem(codes, 0).linear_search(queries, 100, n_procs=2)      # the CPU-baseline leg
    np.testing.assert_array_equal(mid, rid)
    np.testing.assert_array_equal(md, rd)


@pytest.mark.parametrize("bits,m", [(64, 4), (128, 8), (64, 8), (64, 2), (128, 4)])      # s = 16, 16, 8, 32, 32
def test_tables_match_reference_build(bits, m):
    n = 5000
    codes, _ = _data(n, bits, 1)
    store = F.RefStore(codes, m)             # the reference's build-tables main(), rank by rank
    mem = F.RefMem(codes, m)
    ix = R.Index(codes, m)
    sub = bits // m // 8
    rng = np.random.default_rng(0)
    for t in range(m):
        for i in rng.integers(0, n, 8):
            key = R.binary_to_int(codes[i, t * sub:(t + 1) * sub])
            rc, ids, bcodes = store.bucket(t, key)
            orc, oids = ix.bucket(t, key)
            mrc, mids, _ = mem.bucket(t, key)
            assert rc == orc == mrc == 0
            np.testing.assert_array_equal(ids, oids)
            np.testing.assert_array_equal(ids, mids)
            np.testing.assert_array_equal(bcodes, codes[ids])
            assert (np.diff(ids.astype(np.int64)) > 0).all()
        key = int(rng.integers(0, 1 << (8 * sub)))
        assert store.bucket(t, key)[0] == ix.bucket(t, key)[0]
    store.close()


def test_reference_integrity_check_passes_on_reference_tables():
    # integrity_check.cc asserts (aborts) on failure, so it runs in a child process
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import reference as F, restatement as R\n"
        "codes = R.synth_codes(12345, 0, 3000, 8)\n"
        "s = F.RefStore(codes, 4)\n"
        "assert s.integrity_check() == 0\n"
        "s.close()\n" % oracle_pkg.HERE.rsplit('/', 1)[0]
    )
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


@pytest.mark.parametrize("n,bits,m,k", [(20000, 64, 4, 10), (30000, 64, 4, 100)])
def test_contract_p3_p4_canonical_vs_reference(n, bits, m, k):
    nq = 8
    codes, queries = _data(n, bits, nq)
    rid, rd, rc, rrad, _ = F.RefMem(codes, m).mih_search(queries, k)
    ix = R.Index(codes, m)
    cid, cd, cc, cst = ix.search(queries, k, order=R.ORDER_CANONICAL, stop=R.STOP_STRICT_M)
    lid, ld, lc = R.linear_search(codes, queries, k)
    np.testing.assert_array_equal(cid, lid)          # P2: canonical MIH == canonical scan
    np.testing.assert_array_equal(cd, ld)
    check_p3(cid, cd, cc, rid, rd, rc, lambda q, i: R.hamming(codes[i], queries[q]))
    for q in range(nq):                              # P4 (m = 4)
        dk = int(cd[q, : cc[q]].max())
        assert cst[q]["radius"] in (rrad[q], rrad[q] + 1)
        if cst[q]["radius"] != rrad[q]:
            assert dk == 4 * (rrad[q] + 1)


def test_bitmap_restatement_matches_reference_bitmap():
    """a10: ImageBitmap::get_idx / set_idx / reset_idx (src/bitmap.cc:22-38, unmodified, through ref_bitmap_ops) against the
    restatement's vo_bitmap_* on a random op sequence - including bit 31 of a word, where the reference shifts a signed 1 -
    and a11: the occupancy bitmap of a table is the reference class fed with every code's key."""
    import ctypes as C
    rng = np.random.default_rng(3)
    n_bits = 1 << 16
    bits = rng.integers(0, n_bits, 6000).astype(np.uint64)
    bits[:64] = np.arange(64, dtype=np.uint64) * 32 + 31          # the sign bit of every one of the first words
    ops = rng.choice([0, 1, 1, 2], size=bits.size).astype(np.int32)
    ref_words = np.zeros(n_bits // 32, dtype=np.uint32)
    ref_get = F.bitmap_ops(ref_words, bits, ops)
    L = R.lib()
    words = np.zeros(n_bits // 32, dtype=np.uint32)
    wp = words.ctypes.data_as(C.POINTER(C.c_uint32))
    for i, (b, op) in enumerate(zip(bits.tolist(), ops.tolist())):
        if op == 0:
            assert L.vo_bitmap_get(wp, b) == ref_get[i]
        elif op == 1:
            L.vo_bitmap_set(wp, b)
        else:
            L.vo_bitmap_reset(wp, b)
    np.testing.assert_array_equal(words, ref_words)
    # occupancy of a table (generate_bitmap.cc:99-125 sets one bit per code and table): s = 16 and s = 8
    for bits_, m in [(64, 4), (64, 8)]:
        codes, _ = _data(4000, bits_, 1)
        ix = R.Index(codes, m)
        sub = bits_ // m // 8
        for t in (0, m - 1):
            keys = np.array([R.binary_to_int(codes[i, t * sub:(t + 1) * sub]) for i in range(codes.shape[0])], dtype=np.uint64)
            ref_occ = np.zeros((1 << (8 * sub)) // 32, dtype=np.uint32)
            F.bitmap_ops(ref_occ, keys, np.ones(keys.size, dtype=np.int32))
            np.testing.assert_array_equal(ix.occupancy_bitmap(t), ref_occ)

