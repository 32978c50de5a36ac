"""The two scan kernels must agree: the linear-scan parity suite re-run with the TMA-ring kernel forced (0) and
with the warp-granular verify-kernel path forced (1); the default (-1) picks by batch size."""
import pytest

import test_gpu_linear as base

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], autouse=True)
def _force(request):
    base.SCAN_BATCHED = request.param
    yield
    base.SCAN_BATCHED = -1


@pytest.mark.parametrize("bits", [64, 128, 256])
@pytest.mark.parametrize("n,nq,k", [(1000, 3, 10), (5000, 17, 100), (70001, 40, 10), (300000, 5, 100)])
def test_both_scan_paths_match_oracle(oracle, bits, n, nq, k):
    base._run(oracle, n, bits, nq, k)


def test_both_scan_paths_edge_cases(oracle):
    base._run(oracle, 50_000, 64, 7, 100, first_id=4_000_000_000 - 50_000)
    base._run(oracle, 7, 64, 3, 10)
    base._run(oracle, 30_000, 64, 3, 1000)
    base.test_linear_heavy_ties(oracle)
