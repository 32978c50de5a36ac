"""The tensor-core verify kernel (tcverify.cuh: int8 tcgen05 GEMM as the distance filter, exact XOR + POPC path
for the rows it flags) must give exactly the answers and statistics of the POPC kernels: the batched-MIH and
linear-scan parity suites are re-run with it forced on, plus cases sized to exercise its pipelines (many tiles per
work item, many items per CTA, query lists longer than one item, ragged tails)."""
import numpy as np
import pytest

import test_gpu_linear as lin
import test_gpu_mih as mih
from verticut_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_tc():
    mih.BATCHED = 1
    mih.EXTRA_PARAMS = {"mih.tc": 1}
    lin.SCAN_BATCHED = 1
    lin.EXTRA_PARAMS = {"scan.tc": 1}
    yield
    mih.BATCHED = 0
    mih.EXTRA_PARAMS = {}
    lin.SCAN_BATCHED = -1
    lin.EXTRA_PARAMS = {}


@pytest.mark.parametrize("bits", [64, 128, 256])
@pytest.mark.parametrize("n,nq,k", [(1000, 3, 10), (5000, 17, 100), (70001, 40, 10), (300000, 5, 100), (40_000, 300, 10)])
def test_tc_scan_matches_oracle(oracle, bits, n, nq, k):
    lin._run(oracle, n, bits, nq, k)


def test_tc_scan_edge_cases(oracle):
    lin._run(oracle, 50_000, 64, 7, 100, first_id=4_000_000_000 - 50_000)
    lin._run(oracle, 7, 64, 3, 10)
    lin._run(oracle, 30_000, 64, 3, 1000)
    lin._run(oracle, 129, 64, 257, 10)           # one more code than a tile, one more query than an item
    lin.test_linear_heavy_ties(oracle)


def test_tc_scan_large(oracle):
    # 4M codes x 520 queries: hundreds of work items per CTA, three query chunks; checked against the oracle on a few
    # queries and against the POPC scan on all of them
    n, nq, k = 4_000_000, 520, 100
    ix = capi.Index(64, 0)
    ix.add_synthetic(n, 12345)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    ix.set_param("scan.batched", 1)
    ix.set_param("scan.tc", 1)
    a = ix.search_linear(queries, k)
    assert ix.get_param("scan.last_tc") == 1 and ix.get_param("scan.last_batched") == 1
    ix.set_param("scan.tc", 0)
    b = ix.search_linear(queries, k)
    assert ix.get_param("scan.last_tc") == 0
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    codes = oracle.synth_codes(12345, 0, n, 8)
    oid, od, oc = oracle.linear_search(codes, queries[:4], k)
    np.testing.assert_array_equal(a[0][:4], oid)
    np.testing.assert_array_equal(a[1][:4], od)
    ix.close()


@pytest.mark.parametrize("n,bits,m,nq,k", [
    (20_000, 64, 4, 16, 10), (200_000, 64, 4, 8, 100), (50_000, 128, 8, 6, 100), (20_000, 64, 8, 5, 10), (5_000, 256, 16, 3, 50),
])
def test_tc_mih_exact_equals_linear_scan(oracle, n, bits, m, nq, k):
    mih._check_exact(oracle, n, bits, m, nq, k)


def test_tc_mih_first_id_small_and_large_k(oracle):
    mih._check_exact(oracle, 30_000, 64, 4, 5, 100, first_id=3_000_000_000)
    mih._check_exact(oracle, 50, 64, 4, 3, 100)
    mih._check_exact(oracle, 20_000, 64, 4, 2, 1000)


def test_tc_mih_heavy_ties(oracle):
    mih.test_mih_heavy_ties(oracle)


@pytest.mark.parametrize("bits,m,r", [(64, 4, 0), (64, 4, 2), (128, 8, 1), (256, 16, 1)])
def test_tc_mih_fixed_radius(oracle, bits, m, r):
    mih.test_mih_fixed_radius(oracle, bits, m, r)


@pytest.mark.parametrize("bits,m,n,nq", [(64, 4, 8_000_000, 600), (128, 8, 2_000_000, 300)])
def test_tc_mih_large_batch_matches_popc_kernel(oracle, bits, m, n, nq):
    # real buckets (~120 / ~30 codes) and real query lists; both verify kernels, every statistic
    k = 100
    ix = capi.Index(bits, m)
    ix.add_synthetic(n, 12345)
    ix.build()
    queries = oracle.synth_codes(67890, 0, nq, bits // 8)
    ix.set_param("mih.batched", 1)
    ix.set_param("mih.tc", 0)
    a = ix.search_mih(queries, k)
    assert ix.get_param("mih.last_tc_steps") == 0
    for tc in (1, -1):
        ix.set_param("mih.tc", tc)
        b = ix.search_mih(queries, k)
        if tc == 1:
            assert ix.get_param("mih.last_tc_steps") == ix.get_param("mih.last_levels")
        for x, y in zip(a[:3], b[:3]):
            np.testing.assert_array_equal(x, y)
        for f in ("radius", "n_results", "probes", "candidates"):
            np.testing.assert_array_equal(a[3][f], b[3][f])
    lid, ld, lc = ix.search_linear(queries[:8], k)
    np.testing.assert_array_equal(a[0][:8], lid)
    np.testing.assert_array_equal(a[1][:8], ld)
    ix.close()


@pytest.mark.parametrize("bits", [64, 128, 256])
@pytest.mark.parametrize("n,nq,k", [(1000, 3, 10), (5000, 17, 100), (70001, 40, 10), (300000, 5, 100), (40_000, 300, 10)])
def test_tc_v1_scan_matches_oracle(oracle, bits, n, nq, k):
    """The first version of the tensor-core kernel (tcverify_v1.cuh, "scan.tc" = 2: A operand through shared memory, up to 256
    queries per tile) - the one that measured faster than the POPC kernel on large scan batches."""
    lin.EXTRA_PARAMS = {"scan.tc": 2}
    lin._run(oracle, n, bits, nq, k)


def test_tc_v1_scan_edge_cases_and_large(oracle):
    lin.EXTRA_PARAMS = {"scan.tc": 2}
    lin._run(oracle, 50_000, 64, 7, 100, first_id=4_000_000_000 - 50_000)
    lin._run(oracle, 7, 64, 3, 10)
    lin._run(oracle, 129, 64, 257, 10)
    lin.test_linear_heavy_ties(oracle)
    n, nq, k = 4_000_000, 520, 100
    ix = capi.Index(64, 0)
    ix.add_synthetic(n, 12345)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    ix.set_param("scan.batched", 1)
    ix.set_param("scan.tc", 2)
    a = ix.search_linear(queries, k)
    assert ix.get_param("scan.last_tc") == 1 and ix.get_param("scan.last_batched") == 1
    ix.set_param("scan.tc", 0)
    b = ix.search_linear(queries, k)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    ix.close()
