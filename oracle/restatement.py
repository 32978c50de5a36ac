"""TEST INFRASTRUCTURE ONLY - ctypes wrapper of the plain-C oracle (verticut_oracle.c)."""
import ctypes as C
import os

import numpy as np

from . import ORACLE_SO, build

ORDER_CANONICAL, ORDER_REFERENCE = 0, 1
STOP_REF4, STOP_STRICT_M, STOP_REF_M, STOP_TABLE_STRICT = 0, 1, 2, 3
APPROX_FACTOR = 20

_lib = None


class Stats(C.Structure):
    _fields_ = [("radius", C.c_uint32), ("probes", C.c_uint64), ("candidates", C.c_uint64), ("unique", C.c_uint64)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(os.path.dirname(ORACLE_SO), "verticut_oracle.c")
        ):
            build()
        L = C.CDLL(ORACLE_SO)
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.vo_binary_to_int.restype = C.c_uint32
        L.vo_binary_to_int.argtypes = [u8p, C.c_int]
        L.vo_hamming.restype = C.c_int
        L.vo_hamming.argtypes = [u8p, u8p, C.c_int]
        L.vo_index_create.restype = C.c_void_p
        L.vo_index_create.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, C.c_uint32]
        L.vo_index_destroy.argtypes = [C.c_void_p]
        L.vo_bucket_get.restype = C.c_int
        L.vo_bucket_get.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, u32p, C.c_uint32, u32p]
        L.vo_occupancy_bitmap.argtypes = [C.c_void_p, C.c_uint32, u32p]
        L.vo_linear_search.restype = C.c_uint32
        L.vo_linear_search.argtypes = [u8p, C.c_uint64, C.c_int, C.c_uint32, u8p, C.c_uint32, u32p, u32p]
        L.vo_linear_search_ref.restype = C.c_uint32
        L.vo_linear_search_ref.argtypes = L.vo_linear_search.argtypes
        L.vo_mih_search.restype = C.c_uint32
        L.vo_mih_search.argtypes = [C.c_void_p, u8p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, u32p, u32p,
                                    C.POINTER(Stats)]
        L.vo_merge_topk.restype = C.c_uint32
        L.vo_merge_topk.argtypes = [u64p, C.c_uint32, C.c_uint32, u64p]
        L.vo_synth_codes.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u8p]
        L.vo_scan_synth.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, u8p, C.c_uint32, C.c_uint32,
                                    C.c_int, C.c_int, u64p]
        L.vo_synth_word.restype = C.c_uint64
        L.vo_synth_word.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.vo_bitmap_get.restype = C.c_int
        L.vo_bitmap_get.argtypes = [u32p, C.c_uint64]
        L.vo_bitmap_set.argtypes = [u32p, C.c_uint64]
        L.vo_bitmap_reset.argtypes = [u32p, C.c_uint64]
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u32(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def synth_codes(seed, first_id, n, nbytes):
    """Seeded synthetic codes, [n, nbytes] uint8 (same generator as the product's on-device one)."""
    out = np.empty((n, nbytes), dtype=np.uint8)
    lib().vo_synth_codes(seed, first_id, n, nbytes, _u8(out))
    return out


def binary_to_int(buf):
    a = np.ascontiguousarray(buf, dtype=np.uint8)
    return lib().vo_binary_to_int(_u8(a), a.size)


def hamming(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    return lib().vo_hamming(_u8(a), _u8(b), a.size)


def linear_search(codes, queries, k, first_id=0, reference_order=False):
    """Returns (ids [nq,k] uint32, dists [nq,k] uint32, counts [nq]).  Canonical: ascending (dist,id);
    reference_order: the reference's heap order (descending distance)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    queries = np.ascontiguousarray(queries, dtype=np.uint8)
    n, nbytes = codes.shape
    nq = queries.shape[0]
    ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    dists = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
    counts = np.zeros(nq, dtype=np.uint32)
    fn = lib().vo_linear_search_ref if reference_order else lib().vo_linear_search
    for q in range(nq):
        counts[q] = fn(_u8(codes), n, nbytes, first_id, _u8(queries[q]), k, _u32(ids[q]), _u32(dists[q]))
    return ids, dists, counts


class Index:
    """CPU MIH index (restatement of build_hash_tables.cc + search_worker.cc)."""

    def __init__(self, codes, n_tables, first_id=0):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.n, self.nbytes = self.codes.shape
        self.m = n_tables
        self.first_id = first_id
        self.h = lib().vo_index_create(_u8(self.codes), self.n, self.nbytes, n_tables, first_id)
        if not self.h:
            raise ValueError("bad index shape")

    def __del__(self):
        if getattr(self, "h", None):
            lib().vo_index_destroy(self.h)
            self.h = None

    def bucket(self, table, key):
        n = C.c_uint32(0)
        cap = max(1, self.n)
        ids = np.empty(cap, dtype=np.uint32)
        rc = lib().vo_bucket_get(self.h, table, key, _u32(ids), cap, C.byref(n))
        return rc, ids[: n.value].copy()

    def occupancy_bitmap(self, table):
        bits = 1 << (8 * self.nbytes // self.m)
        data = np.zeros(bits // 32, dtype=np.uint32)
        lib().vo_occupancy_bitmap(self.h, table, _u32(data))
        return data

    def search(self, queries, k, order=ORDER_CANONICAL, stop=STOP_STRICT_M, approximate=False, max_radius=-1):
        queries = np.ascontiguousarray(queries, dtype=np.uint8)
        nq = queries.shape[0]
        ids = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        dists = np.full((nq, k), 0xFFFFFFFF, dtype=np.uint32)
        counts = np.zeros(nq, dtype=np.uint32)
        stats = []
        for q in range(nq):
            st = Stats()
            counts[q] = lib().vo_mih_search(self.h, _u8(queries[q]), k, order, stop, int(approximate), max_radius,
                                            _u32(ids[q]), _u32(dists[q]), C.byref(st))
            stats.append(dict(radius=st.radius, probes=st.probes, candidates=st.candidates, unique=st.unique))
        return ids, dists, counts, stats


def merge_topk(lists, k):
    """lists: [n_lists, k] uint64 packed keys (UINT64_MAX = empty).  Returns the merged ascending keys."""
    lists = np.ascontiguousarray(lists, dtype=np.uint64)
    out = np.full(k, np.iinfo(np.uint64).max, dtype=np.uint64)
    n = lib().vo_merge_topk(lists.ctypes.data_as(C.POINTER(C.c_uint64)), lists.shape[0], k,
                            out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out[:n]


def _scan_synth_range(args):
    seed, first_id, stride, i0, i1, nbytes, queries, k, m, max_radius = args
    out = np.empty((queries.shape[0], k), dtype=np.uint64)
    lib().vo_scan_synth(seed, first_id, stride, i0, i1, nbytes, _u8(queries), queries.shape[0], k, m, max_radius,
                        out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def scan_synth(seed, first_id, stride, n, nbytes, queries, k, m=0, max_radius=-1, n_procs=1):
    """Canonical top-k of every query over the n synthetic codes with ids first_id + i * stride, generated on the fly
    (vo_scan_synth): [nq, k] uint64 packed words, ascending, UINT64_MAX padded.  m > 0 and max_radius >= 0: among the
    candidates of the fixed-radius MIH search only.  n_procs > 1 forks workers over ranges of i and merges their lists."""
    queries = np.ascontiguousarray(queries, dtype=np.uint8).reshape(-1, nbytes)
    n_procs = max(1, min(int(n_procs), int(n) // 1_000_000 or 1))
    bounds = [n * p // n_procs for p in range(n_procs + 1)]
    jobs = [(seed, first_id, stride, bounds[p], bounds[p + 1], nbytes, queries, k, m, max_radius) for p in range(n_procs)]
    if n_procs == 1:
        parts = [_scan_synth_range(jobs[0])]
    else:
        import multiprocessing as mp
        lib()                                            # built before the fork
        with mp.get_context("fork").Pool(n_procs) as pool:
            parts = pool.map(_scan_synth_range, jobs)
    out = np.empty((queries.shape[0], k), dtype=np.uint64)
    for q in range(queries.shape[0]):
        merged = merge_topk(np.stack([p[q] for p in parts]), k)
        out[q, : len(merged)] = merged
        out[q, len(merged):] = np.iinfo(np.uint64).max
    return out
