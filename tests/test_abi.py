"""The C-ABI library loads and exports every symbol include/verticut_gpu.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "verticut_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vc_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from verticut_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "libverticut_gpu.so does not export %s" % name
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS and the header disagree"
    assert lib.vc_abi_version() == 1


def test_no_cpu_fallback_without_device():
    # Without a CUDA device the product must fail loudly, not compute on the host.
    from verticut_b200 import capi
    try:
        n = capi.device_count()
    except capi.VerticutError:
        n = 0
    if n > 0:
        pytest.skip("a device is present")
    with pytest.raises(capi.VerticutError):
        capi.Index(64, 4)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "verticut_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace("oracle/verticut_oracle.c", ""), f


def test_synth_generator_matches_oracle(oracle):
    from verticut_b200 import capi
    for seed, idx, w in [(12345, 0, 0), (12345, 999_999_999, 0), (7, 123456, 3), (67890, 5, 1)]:
        assert capi.synth_word(seed, idx, w) == oracle.lib().vo_synth_word(seed, idx, w)


def test_nccl_hook_marshals_an_in_place_uint32_sum():
    """vc_nccl_allreduce_hook (include/verticut_gpu.h) against a stand-in for ncclAllReduce: in place, count = words,
    ncclUint32 (3), ncclSum (0), communicator and stream passed through; a non-zero ncclResult_t is an error."""
    import ctypes as C
    from verticut_b200 import capi
    L = capi.lib()
    seen = {}
    FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p)

    def fake(send, recv, count, dtype, op, comm, stream):
        seen.update(send=send, recv=recv, count=count, dtype=dtype, op=op, comm=comm, stream=stream)
        return seen.get("rc", 0)
    cb = FN(fake)
    hook = capi.NcclHook(C.cast(cb, C.c_void_p), C.c_void_p(0xC0FFEE))
    user = C.cast(C.pointer(hook), C.c_void_p)
    assert L.vc_nccl_allreduce_hook(user, C.c_void_p(0x7000), 12345, C.c_void_p(0x55)) == 0
    assert seen == dict(send=0x7000, recv=0x7000, count=12345, dtype=3, op=0, comm=0xC0FFEE, stream=0x55)
    seen["rc"] = 5
    assert L.vc_nccl_allreduce_hook(user, C.c_void_p(0x7000), 1, None) != 0 and b"ncclAllReduce" in L.vc_last_error()
    assert L.vc_nccl_allreduce_hook(None, C.c_void_p(0x7000), 1, None) != 0


def test_tools_and_bench_compile_and_stay_off_the_oracle():
    """tools/*.py run on the GPU box only: at least they must parse here, and - like the product - the measurement tools compare
    with numpy or with another GPU path, never with oracle/ (only tests/, smoke() and bench.py's CPU legs may use it)."""
    import glob
    import py_compile
    paths = sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(paths) >= 10
    for p in paths:
        py_compile.compile(p, doraise=True)
    for p in glob.glob(os.path.join(ROOT, "tools", "*.py")):
        src = open(p).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), "%s imports the oracle" % p
