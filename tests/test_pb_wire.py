"""The KV schema on the wire: the hand-written messages of
verticut_b200/host/include/image_search.pb.h against the real protobuf runtime (the `protobuf` Python package) on the
schema of the reference's src/image_search.proto:3-27 - bytes written by one side must be the other side's bytes."""
import os
import random
import subprocess

import pytest

pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "verticut_b200", "host")
F = descriptor_pb2.FieldDescriptorProto


def _schema():
    """src/image_search.proto, built as a descriptor (protoc is not available here)."""
    fd = descriptor_pb2.FileDescriptorProto(name="image_search.proto", syntax="proto2")

    def msg(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, number, ftype, label, tname in fields:
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if tname:
                f.type_name = tname
    msg("ID", ("id", 1, F.TYPE_UINT32, F.LABEL_REQUIRED, None))
    msg("BinaryCode", ("code", 1, F.TYPE_BYTES, F.LABEL_REQUIRED, None))
    msg("HashIndex", ("table_id", 1, F.TYPE_UINT32, F.LABEL_REQUIRED, None), ("index", 2, F.TYPE_UINT32, F.LABEL_REQUIRED, None))
    msg("ID_Code_Pair", ("id", 1, F.TYPE_UINT32, F.LABEL_REQUIRED, None), ("code", 2, F.TYPE_BYTES, F.LABEL_REQUIRED, None))
    msg("ImageList", ("images", 1, F.TYPE_UINT32, F.LABEL_REPEATED, None))
    msg("Image_List", ("images", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".ID_Code_Pair"))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName(n)) for n in ("ID", "BinaryCode", "HashIndex", "ID_Code_Pair", "ImageList", "Image_List")}


@pytest.fixture(scope="module")
def tool():
    res = subprocess.run(["make", "-C", HOST, "pb-wire"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    p = subprocess.Popen([os.path.join(HOST, "bin", "pb-wire")], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)

    def ask(line):
        p.stdin.write(line + "\n")
        p.stdin.flush()
        return p.stdout.readline().strip()
    yield ask
    p.stdin.close()
    p.wait(timeout=20)


def _hex(b):
    return b.hex() if b else "-"


U32 = [0, 1, 127, 128, 255, 300, 16383, 16384, 65535, 65536, (1 << 28) - 1, 1 << 28, (1 << 31) - 1, 1 << 31, (1 << 32) - 1]
CODES = [b"", b"\x00", b"\xff" * 8, bytes(range(16)), bytes(range(32)), bytes(200)]     # 64 / 128 / 256-bit codes and odd sizes


def test_scalar_messages_both_directions(tool):
    M = _schema()
    rng = random.Random(5)
    for v in U32 + [rng.randrange(1 << 32) for _ in range(50)]:
        want = M["ID"](id=v).SerializeToString()
        assert tool("enc ID %d" % v) == _hex(want)
        assert tool("dec ID %s" % _hex(want)) == str(v)
        w = rng.choice(U32)
        want = M["HashIndex"](table_id=v, index=w).SerializeToString()
        assert tool("enc HashIndex %d %d" % (v, w)) == _hex(want)
        assert tool("dec HashIndex %s" % _hex(want)) == "%d %d" % (v, w)
    for c in CODES + [bytes(rng.randrange(256) for _ in range(rng.choice((8, 16, 32)))) for _ in range(30)]:
        want = M["BinaryCode"](code=c).SerializeToString()
        assert tool("enc BinaryCode %s" % _hex(c)) == _hex(want)
        assert tool("dec BinaryCode %s" % _hex(want)) == _hex(c)
        v = rng.choice(U32)
        want = M["ID_Code_Pair"](id=v, code=c).SerializeToString()
        assert tool("enc ID_Code_Pair %d %s" % (v, _hex(c))) == _hex(want)
        assert tool("dec ID_Code_Pair %s" % _hex(want)) == "%d %s" % (v, _hex(c))


def test_bucket_values_both_directions(tool):
    """Image_List is the value of a bucket (src/build_hash_tables.cc:53-64): from empty to a few thousand members."""
    M = _schema()
    rng = random.Random(6)
    for n, nbytes in [(0, 8), (1, 8), (2, 16), (17, 32), (300, 8), (3000, 16)]:
        pairs = [(rng.choice(U32) if i % 7 == 0 else rng.randrange(1 << 32), bytes(rng.randrange(256) for _ in range(nbytes))) for i in range(n)]
        m = M["Image_List"]()
        for i, c in pairs:
            m.images.add(id=i, code=c)
        want = m.SerializeToString()
        text = " ".join("%d:%s" % (i, _hex(c)) for i, c in pairs)
        assert tool(("enc Image_List " + text).strip()) == _hex(want)
        assert tool("dec Image_List %s" % _hex(want)) == text
        ids = [i for i, _ in pairs]
        want = M["ImageList"](images=ids).SerializeToString()          # proto2 repeated uint32: not packed
        assert tool(("enc ImageList " + " ".join(map(str, ids))).strip()) == _hex(want)
        assert tool("dec ImageList %s" % _hex(want)) == " ".join(map(str, ids))


def test_truncated_bytes_are_rejected(tool):
    M = _schema()
    m = M["Image_List"]()
    m.images.add(id=5, code=b"\x01\x02\x03\x04\x05\x06\x07\x08")
    good = m.SerializeToString()
    for cut in range(1, len(good)):
        assert tool("dec Image_List %s" % _hex(good[:cut])) == "ERROR"
    assert tool("dec ID_Code_Pair 0805") == "ERROR"                    # required field `code` missing
    assert tool("dec ID 08ffffffffffffffffffffff01") == "ERROR"          # varint longer than 10 bytes
