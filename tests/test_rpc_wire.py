"""The query RPC on the wire: the hand-written msgpack-rpc front end of
verticut_b200/host/include/image_search_rpc.h against the third-party `msgpack` package as an independent codec.

Reference: src/image_search_client.cc:19-34 (calls), src/image_search_server.cc:22-50 (dispatch and error behaviour).
No GPU: the server under test is host/bin/rpc-stub, the same front end over a deterministic stand-in service."""
import os
import socket
import subprocess

import pytest

msgpack = pytest.importorskip("msgpack")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "verticut_b200", "host")
STUB = os.path.join(HOST, "bin", "rpc-stub")


@pytest.fixture(scope="module")
def stub():
    res = subprocess.run(["make", "-C", HOST, "rpc-stub"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    return STUB


@pytest.fixture()
def server(stub):
    p = subprocess.Popen([stub, "serve", "0"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
    line = p.stdout.readline().split()
    assert line[0] == "port"
    yield int(line[1])
    p.stdin.close()
    assert p.wait(timeout=20) == 0


class Conn:
    def __init__(self, port):
        self.s = socket.create_connection(("127.0.0.1", port), timeout=20)
        self.u = msgpack.Unpacker(raw=False, strict_map_key=False)

    def send(self, obj, **kw):
        self.s.sendall(msgpack.packb(obj, **kw))

    def recv(self):
        while True:
            for obj in self.u:
                return obj
            data = self.s.recv(1 << 16)
            assert data, "server closed the connection"
            self.u.feed(data)

    def call(self, msgid, method, params, **kw):
        self.send([0, msgid, method, params], **kw)
        return self.recv()


def test_selftest_binary(stub):
    res = subprocess.run([stub, "selftest"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == "ok", res.stdout + res.stderr


def test_request_bytes_match_the_msgpack_package():
    # what vcrpc::request_ping / request_search emit (pinned byte for byte inside `rpc-stub selftest`) is what the
    # msgpack package produces in its old-raw mode - the format of the reference's msgpack 0.5 peers
    assert msgpack.packb([0, 1, "ping", ["hi"]], use_bin_type=False) == b"\x94\x00\x01\xa4ping\x91\xa2hi"
    assert msgpack.packb([0, 7, "search_image_by_id", [300, 70000, True]], use_bin_type=False) == \
        b"\x94\x00\x07\xb2search_image_by_id\x93\xcd\x01\x2c\xce\x00\x01\x11\x70\xc3"


def test_ping_and_search(server):
    c = Conn(server)
    assert c.call(1, "ping", ["hello"]) == [1, 1, None, "hello"]
    assert c.call(2, "search_image_by_id", [40, 5, False]) == [1, 2, None, [[40, 4], [41, 3], [42, 2], [43, 1], [44, 0]]]
    assert c.call(3, "search_image_by_id", [4000000000, 2, True]) == [1, 3, None, [[4000000000, 1001], [4000000001, 1000]]]
    big = c.call(4, "search_image_by_id", [0, 70000, False])          # array32 of pairs, ids and distances up to uint32
    assert big[:3] == [1, 4, None] and len(big[3]) == 70000 and big[3][0] == [0, 69999] and big[3][-1] == [69999, 0]
    # extra parameters are ignored, as msgpack::type::tuple's convert does
    assert c.call(5, "ping", ["a", "b"]) == [1, 5, None, "a"]


def test_error_behaviour(server):
    c = Conn(server)
    assert c.call(1, "no_such_method", []) == [1, 1, 1, None]               # NO_METHOD_ERROR
    assert c.call(2, "ping", [7]) == [1, 2, 2, None]                        # ARGUMENT_ERROR: not a string
    assert c.call(3, "ping", []) == [1, 3, 2, None]                         # ARGUMENT_ERROR: too few parameters
    assert c.call(4, "search_image_by_id", [1, 2]) == [1, 4, 2, None]
    assert c.call(5, "search_image_by_id", [1, 2, 1]) == [1, 5, 2, None]    # bool must be a bool
    assert c.call(6, "search_image_by_id", [-1, 2, False]) == [1, 6, 2, None]
    assert c.call(7, "search_image_by_id", [1 << 32, 2, False]) == [1, 7, 2, None]
    assert c.call(8, "search_image_by_id", [0xFFFFFFFF, 2, False]) == [1, 8, "Can't find match", None]   # what() of the exception
    assert c.call(9, "ping", ["still there"]) == [1, 9, None, "still there"]


def test_streaming_notifications_and_new_style_strings(server):
    c = Conn(server)
    # a notification gets no answer; two requests in one segment are answered in order
    c.s.sendall(msgpack.packb([2, "ping", ["ignored"]]) + msgpack.packb([0, 10, "ping", ["x"]]) + msgpack.packb([0, 11, "ping", ["y"]]))
    assert c.recv() == [1, 10, None, "x"]
    assert c.recv() == [1, 11, None, "y"]
    # a request that arrives byte by byte
    for b in msgpack.packb([0, 12, "search_image_by_id", [9, 1, False]]):
        c.s.sendall(bytes([b]))
    assert c.recv() == [1, 12, None, [[9, 0]]]
    # str8 (current msgpack) and bin parameters are accepted; the answer uses raw16, which every generation reads
    s = "z" * 200
    assert msgpack.packb(s, use_bin_type=True)[0] == 0xd9
    assert c.call(13, "ping", [s], use_bin_type=True) == [1, 13, None, s]
    c.s.sendall(msgpack.packb([0, 14, "ping", [s.encode()]], use_bin_type=True))
    assert c.recv() == [1, 14, None, s]


def test_garbage_closes_the_connection_only(server):
    c = Conn(server)
    c.s.sendall(b"\xc1")                                   # the one byte MessagePack never uses
    assert c.s.recv(16) == b""
    assert Conn(server).call(1, "ping", ["next client"]) == [1, 1, None, "next client"]


def test_item_flood_and_oversized_requests_close_the_connection_only(server):
    """A request is seven items and tens of bytes: a message that declares thousands of items (each decoded Value costs ~100
    bytes of server memory) or keeps streaming without completing an object is dropped, the server lives on."""
    c = Conn(server)
    c.s.sendall(b"\xdc\x10\x00" + b"\xc0" * 4096)          # array16 of 4096 nils
    assert c.s.recv(16) == b""
    c = Conn(server)
    c.s.sendall(b"\x94\x00\x01\xa4ping\x91\xdb\x00\x40\x00\x00" + b"a" * ((1 << 20) + 4096))     # a ping whose str32 announces 4 MB
    assert c.s.recv(16) == b""
    assert Conn(server).call(1, "ping", ["next client"]) == [1, 1, None, "next client"]


def test_client_cli_against_the_stub(server, stub):
    cli = os.path.join(HOST, "bin", "image-search-client")
    r = subprocess.run([cli, "127.0.0.1", str(server), "ping", "hello"], capture_output=True, text=True, timeout=30)
    assert r.returncode == 0 and r.stdout == "hello\n"
    r = subprocess.run([cli, "127.0.0.1", str(server), "byid", "40", "3"], capture_output=True, text=True, timeout=30)
    assert r.returncode == 0 and r.stdout == "40 : 2\n41 : 1\n42 : 0\n"
    r = subprocess.run([cli, "127.0.0.1", str(server), "byid", "40", "2", "-a"], capture_output=True, text=True, timeout=30)
    assert r.returncode == 0 and r.stdout == "40 : 1001\n41 : 1000\n"
    r = subprocess.run([cli, "127.0.0.1", str(server), "byid", str(0xFFFFFFFF), "2"], capture_output=True, text=True, timeout=30)
    assert r.returncode == 1 and r.stderr == "Can't find match\n"
    r = subprocess.run([cli, "127.0.0.1", "1", "ping", "x"], capture_output=True, text=True, timeout=30)    # nobody listens
    assert r.returncode == 1 and "cannot connect" in r.stderr
