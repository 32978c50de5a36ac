"""Index persistence (vc_index_save / vc_index_load): a reloaded index answers exactly like the one that was saved."""
import numpy as np
import pytest

from verticut_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits,m", [(64, 4), (128, 8), (64, 2), (64, 0)])
def test_save_load_round_trip(tmp_path, oracle, bits, m):
    n, nq, k = 30_000, 6, 20
    codes = oracle.synth_codes(12345, 0, n, bits // 8)
    queries = oracle.synth_codes(67890, 0, nq, bits // 8)
    ix = capi.Index(bits, m, first_id=1000)
    ix.add(codes)
    ix.build()
    path = str(tmp_path / "index.vc")
    ix.save(path)
    a_lin = ix.search_linear(queries, k)
    a_mih = ix.search_mih(queries, k) if m else None
    ix.close()
    ix2 = capi.Index.load(path)
    inf = ix2.info()
    assert (inf["n_codes"], inf["code_bits"], inf["n_tables"], inf["first_id"], inf["built"]) == (n, bits, m, 1000, 1)
    b_lin = ix2.search_linear(queries, k)
    for x, y in zip(a_lin, b_lin):
        np.testing.assert_array_equal(x, y)
    if m:
        b_mih = ix2.search_mih(queries, k)
        for x, y in zip(a_mih[:3], b_mih[:3]):
            np.testing.assert_array_equal(x, y)
        sub = bits // m // 8
        key = oracle.binary_to_int(codes[7, :sub])
        rc, ids, bc = ix2.bucket_get(0, key)
        assert rc == 0 and 1007 in ids.tolist()
    ix2.close()


def test_load_rejects_garbage(tmp_path):
    p = tmp_path / "bad.vc"
    p.write_bytes(b"not an index at all")
    with pytest.raises(capi.VerticutError):
        capi.Index.load(str(p))


@pytest.mark.parametrize("bits,m", [(64, 4), (64, 2)])
def test_load_rejects_a_corrupt_bucket_directory(tmp_path, oracle, bits, m):
    """A file whose row_ptr (dense tables) or occupancy bitmap (s = 32) was damaged must be refused at load time: every search
    kernel trusts those arrays for its addresses."""
    n = 20_000
    codes = oracle.synth_codes(12345, 0, n, bits // 8)
    ix = capi.Index(bits, m)
    ix.add(codes)
    ix.build()
    path = tmp_path / "index.vc"
    ix.save(str(path))
    ix.close()
    raw = bytearray(path.read_bytes())
    header, table_header = 40, 16
    first_table = header + n * (bits // 8) + table_header          # first row_ptr entry of table 0
    # dense: a bucket start beyond the code count; sparse: the same (row_ptr comes first in both layouts)
    bad = bytearray(raw)
    bad[first_table + 4 * 100: first_table + 4 * 100 + 4] = (0x7FFFFFFF).to_bytes(4, "little")
    p1 = tmp_path / "bad1.vc"
    p1.write_bytes(bytes(bad))
    with pytest.raises(capi.VerticutError):
        capi.Index.load(str(p1))
    if bits // m == 32:
        # flip one bit of the occupancy bitmap: popcount and rank directory no longer agree with the header
        n_unique = int.from_bytes(raw[header + n * (bits // 8) + 4: header + n * (bits // 8) + 8], "little")
        bitmap_off = first_table + 4 * (n_unique + 1)
        bad = bytearray(raw)
        bad[bitmap_off + 12345] ^= 0x10
        p2 = tmp_path / "bad2.vc"
        p2.write_bytes(bytes(bad))
        with pytest.raises(capi.VerticutError):
            capi.Index.load(str(p2))
    capi.Index.load(str(path)).close()                               # the undamaged file still loads
