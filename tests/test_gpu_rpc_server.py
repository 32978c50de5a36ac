"""image-search-server (msgpack-rpc in front of the GPU index) end to end: a msgpack client against the server, results vs
the CPU oracle (the wire format itself is covered on CPU by tests/test_rpc_wire.py).  Replaces the reference's
src/image_search_server.cc:58-83 / src/image_search_client.cc:24-34 round trip."""
import os
import signal
import socket
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "verticut_b200", "host", "bin")


def test_server_answers_like_the_oracle(tmp_path, oracle):
    msgpack = pytest.importorskip("msgpack")
    n, k = 30_000, 10
    codes = oracle.synth_codes(12345, 0, n, 8)
    cf = tmp_path / "lsh.code"
    cf.write_bytes(codes.tobytes())
    p = subprocess.Popen([os.path.join(BIN, "image-search-server"), "-f", str(cf), "-b", "64", "-t", "4", "-p", "0", "-i", "127.0.0.1", "-n", "2"],
                         stdout=subprocess.PIPE, text=True)
    try:
        port = None
        for _ in range(50):
            line = p.stdout.readline()
            if line.startswith("port "):
                port = int(line.split()[1])
                break
        assert port
        s = socket.create_connection(("127.0.0.1", port), timeout=120)
        u = msgpack.Unpacker(raw=False)

        def call(msgid, method, params):
            s.sendall(msgpack.packb([0, msgid, method, params], use_bin_type=False))
            while True:
                for obj in u:
                    return obj
                u.feed(s.recv(1 << 16))

        assert call(1, "ping", ["hi"]) == [1, 1, None, "hi"]
        for msgid, qid in enumerate((0, 17, n - 1), start=2):
            oid, od, _ = oracle.linear_search(codes, codes[qid:qid + 1], k)
            want = [[int(a), int(b)] for a, b in zip(oid[0][::-1], od[0][::-1])]          # descending distance, as the reference lists them
            assert call(msgid, "search_image_by_id", [qid, k, False]) == [1, msgid, None, want]
        assert call(9, "search_image_by_id", [n + 5, k, False]) == [1, 9, "Can't find match", None]
        # a bad knn is the caller's error, not the server's death: the process holds the only copy of the GPU index
        for msgid, knn in enumerate((0, 2049, 3000, 2 ** 31, 2 ** 32 - 1), start=20):
            reply = call(msgid, "search_image_by_id", [0, knn, False])
            assert reply[:2] == [1, msgid] and isinstance(reply[2], str) and "knn" in reply[2] and reply[3] is None
        assert p.poll() is None
        oid, od, _ = oracle.linear_search(codes, codes[5:6], 2048)                          # the largest legal k still works
        got = call(30, "search_image_by_id", [5, 2048, False])
        assert got[2] is None and [g[0] for g in got[3]] == [int(a) for a in oid[0][::-1]]
    finally:
        p.send_signal(signal.SIGINT)
        p.wait(timeout=30)
