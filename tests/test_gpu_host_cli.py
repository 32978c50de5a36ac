"""The C++ host mirror (GpuTableProxy : BaseProxy, SearchWorker, the reference-style CLIs) end to end on the GPU:
raw code / query files in the reference's format, stdout in the reference's formats, results vs the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "verticut_b200", "host", "bin")


def _files(tmp_path, oracle, n, bits, nq):
    codes = oracle.synth_codes(12345, 0, n, bits // 8)
    queries = oracle.synth_codes(67890, 0, nq, bits // 8)
    cf, qf = tmp_path / "lsh.code", tmp_path / "query.code"
    cf.write_bytes(codes.tobytes())
    qf.write_bytes(queries.tobytes())
    return codes, queries, str(cf), str(qf)


def _run(args):
    res = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


@pytest.mark.parametrize("bits,m", [(64, 4), (128, 8)])
def test_cli_mih_matches_oracle(tmp_path, oracle, bits, m):
    n, nq, k = 30_000, 5, 10
    codes, queries, cf, qf = _files(tmp_path, oracle, n, bits, nq)
    out = _run([os.path.join(BIN, "image-search"), "mih", "-s", "gpu", "-f", cf, "-q", qf, "-b", str(bits), "-n", str(m), "-k", str(k)])
    oid, od, oc = oracle.linear_search(codes, queries, k)
    blocks = re.split(r"query \d+\n", out)[1:]
    assert len(blocks) == nq
    for q, blk in enumerate(blocks):
        pairs = [(int(a), int(b)) for a, b in re.findall(r"^(\d+) : (\d+)$", blk, flags=re.M)]
        assert pairs == list(zip(oid[q][::-1].tolist(), od[q][::-1].tolist()))      # descending distance, as the reference prints
    assert "n_sub_reads" in out and "radius" in out and "-------Timings-------" in out


def test_cli_linear_and_accuracy_and_integrity(tmp_path, oracle):
    n, nq, k = 20_000, 3, 10
    codes, queries, cf, qf = _files(tmp_path, oracle, n, 64, nq)
    out = _run([os.path.join(BIN, "image-search"), "linear", "-f", cf, "-q", qf, "-b", "64", "-n", "4", "-k", str(k)])
    got = [(int(a), int(b)) for a, b in re.findall(r"Find image with id=(\d+) and hamming_dist=(\d+)", out)]
    oid, od, oc = oracle.linear_search(codes, queries, k)
    want = []
    for q in range(nq):
        want += list(zip(oid[q][::-1].tolist(), od[q][::-1].tolist()))
    assert got == want
    out = _run([os.path.join(BIN, "image-search"), "accuracy", "-f", cf, "-q", qf, "-b", "64", "-n", "4", "-k", str(k)])
    rows = [ln.split() for ln in out.splitlines() if re.match(r"^[\d.]+ [\d.]+ [\d.]+$", ln)]
    assert len(rows) == nq and all(float(r[1]) >= float(r[0]) for r in rows)    # approximate is never better than exact
    out = _run([os.path.join(BIN, "image-search"), "integrity", "-f", cf, "-b", "64", "-n", "4"])
    assert "%d images, 0 errors" % n in out


def test_build_tables_via_baseproxy_put(tmp_path, oracle):
    # the reference's build loop (get -> append -> put per code and table) against GpuTableProxy, then the
    # reference's integrity check through BaseProxy::get
    n = 1500
    codes, _, cf, _ = _files(tmp_path, oracle, n, 64, 1)
    out = _run([os.path.join(BIN, "build-tables"), "--via-put", "--check", "-f", cf, "-b", "64", "-n", "4"])
    assert "images : %d" % n in out and "%d images, 0 errors" % n in out
    out = _run([os.path.join(BIN, "build-tables"), "--check", "-f", cf, "-b", "64", "-n", "4"])
    assert "%d images, 0 errors" % n in out


def test_cli_query_by_id(tmp_path, oracle):
    # search_image_by_id of the reference's image_search_client, served from the GPU index (main table kept)
    n, k = 20_000, 10
    codes, _, cf, _ = _files(tmp_path, oracle, n, 64, 1)
    qid = 1234
    out = _run([os.path.join(BIN, "image-search"), "byid", "-f", cf, "-I", str(qid), "-b", "64", "-n", "4", "-k", str(k)])
    pairs = [(int(a), int(b)) for a, b in re.findall(r"^(\d+) : (\d+)$", out, flags=re.M)]
    oid, od, oc = oracle.linear_search(codes, codes[qid:qid + 1], k)
    assert pairs == list(zip(oid[0][::-1].tolist(), od[0][::-1].tolist()))
    assert pairs[-1] == (qid, 0)                                  # the image itself is its own nearest neighbour
    res = subprocess.run([os.path.join(BIN, "image-search"), "byid", "-f", cf, "-I", str(n + 5), "-b", "64", "-n", "4"],
                         capture_output=True, text=True)
    assert res.returncode != 0 and "Can't find match" in res.stderr


def test_cli_build_once_search_from_index_file(tmp_path, oracle):
    # build-tables -o writes the index; the search tool reads it with -x instead of rebuilding from the code file
    n, nq, k = 20_000, 3, 10
    codes, queries, cf, qf = _files(tmp_path, oracle, n, 64, nq)
    idx = str(tmp_path / "tables.vc")
    _run([os.path.join(BIN, "build-tables"), "-f", cf, "-b", "64", "-n", "4", "-o", idx])
    out = _run([os.path.join(BIN, "image-search"), "mih", "-x", idx, "-q", qf, "-b", "64", "-n", "4", "-k", str(k)])
    oid, od, oc = oracle.linear_search(codes, queries, k)
    blocks = re.split(r"query \d+\n", out)[1:]
    assert len(blocks) == nq
    for q, blk in enumerate(blocks):
        pairs = [(int(a), int(b)) for a, b in re.findall(r"^(\d+) : (\d+)$", blk, flags=re.M)]
        assert pairs == list(zip(oid[q][::-1].tolist(), od[q][::-1].tolist()))


def test_reference_positional_command_line(tmp_path, oracle):
    """distributed-image-search <config> <count> <bits> <substring bits> <k> <server> <read mode> <approximate> <query id> [query file]
    (src/distributed_image_search.cc:141-156; what run_distributed_search.py:74-79 and the RPC server's popen line start)."""
    n, nq, k = 20_000, 3, 10
    codes, queries, cf, qf = _files(tmp_path, oracle, n, 128, nq)
    idx = str(tmp_path / "tables.vc")
    _run([os.path.join(BIN, "build-tables"), "-f", cf, "-b", "128", "-n", "8", "-o", idx])
    exe = os.path.join(BIN, "distributed-image-search")
    for line in ("index " + idx, "codes " + cf):
        cfg = tmp_path / "gpu.cnf"
        cfg.write_text("0\n# tables of this server\n" + line + "\n")
        # query by id: "id : dist" lines, descending distance (the format image_search_server.cc:94 parses)
        qid = 17
        out = _run([exe, str(cfg), str(n), "128", "16", str(k), "gpu", "0", "0", str(qid)])
        oid, od, _ = oracle.linear_search(codes, codes[qid:qid + 1], k)
        pairs = [(int(a), int(b)) for a, b in re.findall(r"^(\d+) : (\d+)$", out, flags=re.M)]
        assert pairs == list(zip(oid[0][::-1].tolist(), od[0][::-1].tolist()))
        # query file: the averaged statistics line of :87-93
        out = _run([exe, str(cfg), str(n), "128", "16", str(k), "gpu", "0", "0", "0", qf])
        m = re.search(r"Averate result : \n0  n_main_reads : 0 , n_sub_reads : (\d+), n_local_reads : (\d+), radius : (\d+), rdma : 0", out)
        assert m and int(m.group(1)) > 0
        assert "-------Timings-------" in out
    # the reference's other servers are not served by this build; too few arguments die as in the reference
    res = subprocess.run([exe, str(cfg), str(n), "128", "16", str(k), "pilaf", "0", "0", "0"], capture_output=True, text=True)
    assert res.returncode != 0 and "Unrecognized server type." in res.stderr
    res = subprocess.run([exe, str(cfg), str(n), "128"], capture_output=True, text=True)
    assert res.returncode != 0 and "Incorrect number of arguments!" in res.stderr


def test_build_tables_writes_generate_bitmap_files(tmp_path, oracle):
    """build-tables --bitmaps: one raw occupancy bitmap per table, named and laid out as src/generate_bitmap.cc:84-87,119-125."""
    n = 30_000
    codes, _, cf, _ = _files(tmp_path, oracle, n, 64, 1)
    _run([os.path.join(BIN, "build-tables"), "-f", cf, "-b", "64", "-n", "4", "--bitmaps"])
    oix = oracle.Index(codes, 4)
    for t in range(4):
        words = np.fromfile(cf + "_bmp_%d_2b_4k.raw" % (t + 1), dtype=np.uint32)
        assert words.size == (1 << 16) // 32
        np.testing.assert_array_equal(words, oix.occupancy_bitmap(t))
