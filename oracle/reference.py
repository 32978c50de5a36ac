"""TEST INFRASTRUCTURE ONLY - ctypes wrapper of oracle/_ref/libverticut_ref.so, i.e. the
reference's own search_worker.cc / linear_search.cc / build_hash_tables.cc / integrity_check.cc
compiled unmodified (see oracle/ref_driver.cc for the shims and the three documented deviations)."""
import ctypes as C
import os
import tempfile

import numpy as np

from . import REF_SO

_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.ref_build_tables.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.ref_integrity_check.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.ref_put_main_table.argtypes = [u8p, C.c_uint64, C.c_int, C.c_uint32]
        L.ref_bucket_get.argtypes = [C.c_uint32, C.c_uint32, C.c_int, u32p, u8p, C.c_uint32, u32p]
        L.ref_mem_bucket_get.argtypes = L.ref_bucket_get.argtypes
        L.ref_mih_search.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32p, u32p, u32p, u32p, u64p]
        L.ref_mem_mih_search.argtypes = L.ref_mih_search.argtypes
        L.ref_linear_search.argtypes = [u8p, C.c_int, C.c_int, C.c_uint32, u32p, u32p, u32p]
        L.ref_bitmap_ops.argtypes = [u32p, C.c_uint64, u64p, C.POINTER(C.c_int), C.c_uint64, u8p]
        L.ref_mem_fixed_radius_search.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32p, u32p, u32p, u64p]
        L.ref_mem_build.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.c_uint32]
        L.ref_mem_linear_search.argtypes = [u8p, C.c_uint64, C.c_int, u8p, C.c_int, C.c_int, C.c_int, u32p, u32p, u32p]
        L.ref_kv_entries.restype = C.c_uint64
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


class RefStore:
    """The reference stack over the in-memory stand-in for the Pilaf store (byte KV): tables are built by
    the reference's build-tables main(), searched by SearchWorker::find through the reference's PilafProxy."""

    def __init__(self, codes, n_tables):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.n, self.nbytes = self.codes.shape
        self.m = n_tables
        self._tmp = tempfile.NamedTemporaryFile(prefix="vc_ref_codes_", suffix=".code", delete=False)
        self._tmp.write(self.codes.tobytes())   # raw records, id = ordinal (build_hash_tables.cc:40-45)
        self._tmp.close()
        lib().ref_reset()
        if n_tables > 0:
            rc = lib().ref_build_tables(self._tmp.name.encode(), self.nbytes * 8, n_tables)
            if rc != 0:
                raise RuntimeError("reference build-tables failed")

    def put_main_table(self):
        if lib().ref_put_main_table(_u8(self.codes), self.n, self.nbytes, 0) != 0:
            raise RuntimeError("reference main-table fill failed")

    def integrity_check(self):
        """Runs the reference's integrity-check main(); a failure is an assert() -> process abort."""
        return lib().ref_integrity_check(self._tmp.name.encode(), self.nbytes * 8, self.m)

    def close(self):
        lib().ref_reset()
        try:
            os.unlink(self._tmp.name)
        except OSError:
            pass

    def bucket(self, table, index):
        cap = max(1, self.n)
        ids = np.empty(cap, dtype=np.uint32)
        codes = np.empty((cap, self.nbytes), dtype=np.uint8)
        n = C.c_uint32(0)
        rc = lib().ref_bucket_get(table, index, self.nbytes, _u32(ids), _u8(codes), cap, C.byref(n))
        return rc, ids[: n.value].copy(), codes[: n.value].copy()

    def mih_search(self, queries, k, approximate=False):
        return _mih(lib().ref_mih_search, queries, self.nbytes, self.m, k, approximate, self.n)

    def linear_search(self, queries, k):
        queries = np.ascontiguousarray(queries, dtype=np.uint8)
        nq = queries.shape[0]
        ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        dists = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        counts = np.zeros(nq, dtype=np.uint32)
        for q in range(nq):
            c = C.c_uint32(0)
            lib().ref_linear_search(_u8(queries[q]), self.nbytes, k, self.n, _u32(ids[q]), _u32(dists[q]), C.byref(c))
            counts[q] = c.value
        return ids, dists, counts


def _mih(fn, queries, nbytes, m, k, approximate, n):
    queries = np.ascontiguousarray(queries, dtype=np.uint8)
    nq = queries.shape[0]
    ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    dists = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    counts = np.zeros(nq, dtype=np.uint32)
    radius = np.zeros(nq, dtype=np.uint32)
    sub_reads = np.zeros((nq, m), dtype=np.uint64)
    rc = fn(_u8(queries), nq, nbytes, m, k, int(approximate), int(n), _u32(ids), _u32(dists), _u32(counts), _u32(radius),
            sub_reads.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc != 0:
        raise RuntimeError("reference search failed")
    return ids, dists, counts, radius, sub_reads


class RefMem:
    """The same reference search code over MemProxy (parsed messages in memory) - the fast CPU-baseline leg."""

    def __init__(self, codes, n_tables=0, first_id=0):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.n, self.nbytes = self.codes.shape
        self.m = n_tables
        if lib().ref_mem_build(_u8(self.codes), self.n, self.nbytes, n_tables, first_id) != 0:
            raise RuntimeError("ref_mem_build failed")

    def bucket(self, table, index):
        cap = max(1, self.n)
        ids = np.empty(cap, dtype=np.uint32)
        codes = np.empty((cap, self.nbytes), dtype=np.uint8)
        n = C.c_uint32(0)
        rc = lib().ref_mem_bucket_get(table, index, self.nbytes, _u32(ids), _u8(codes), cap, C.byref(n))
        return rc, ids[: n.value].copy(), codes[: n.value].copy()

    def mih_search(self, queries, k, approximate=False):
        return _mih(lib().ref_mem_mih_search, queries, self.nbytes, self.m, k, approximate, self.n)

    def fixed_radius_search(self, queries, k, max_radius):
        """Radii 0 .. max_radius, no stop rule (config C5): the reference's own probing and verification under a subclass that
        leaves out the stop test (FixedRadiusWorker, ref_driver.cc).  Returns ids, dists (descending), counts, sub_reads."""
        queries = np.ascontiguousarray(queries, dtype=np.uint8)
        nq = queries.shape[0]
        ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        dists = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        counts = np.zeros(nq, dtype=np.uint32)
        sub_reads = np.zeros((nq, self.m), dtype=np.uint64)
        rc = lib().ref_mem_fixed_radius_search(_u8(queries), nq, self.nbytes, self.m, k, int(max_radius), int(self.n), _u32(ids), _u32(dists),
                                               _u32(counts), sub_reads.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc != 0:
            raise RuntimeError("reference fixed-radius search failed")
        return ids, dists, counts, sub_reads

    def linear_search(self, queries, k, n_procs=1):
        queries = np.ascontiguousarray(queries, dtype=np.uint8)
        nq = queries.shape[0]
        ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        dists = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        counts = np.zeros(nq, dtype=np.uint32)
        rc = lib().ref_mem_linear_search(_u8(self.codes), self.n, self.nbytes, _u8(queries), nq, k, n_procs,
                                         _u32(ids), _u32(dists), _u32(counts))
        if rc != 0:
            raise RuntimeError("reference linear search failed")
        return ids, dists, counts


def bitmap_ops(words, bits, ops):
    """Runs get (0) / set (1) / reset (2) on the reference's own ImageBitmap (src/bitmap.cc) over `words` (uint32 array, modified
    in place); returns the get results (0 / 1; undefined where the op was not a get)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint64)
    ops = np.ascontiguousarray(ops, dtype=np.int32)
    out = np.zeros(bits.size, dtype=np.uint8)
    rc = lib().ref_bitmap_ops(_u32(words), words.nbytes, bits.ctypes.data_as(C.POINTER(C.c_uint64)), ops.ctypes.data_as(C.POINTER(C.c_int)),
                              bits.size, _u8(out))
    if rc != 0:
        raise RuntimeError("reference bitmap ops failed")
    return out

