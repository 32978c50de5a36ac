"""ctypes binding of libverticut_gpu.so - the C ABI declared in include/verticut_gpu.h.

This module is plumbing only: argument marshalling around the exported C functions.  It never
computes anything itself and it has NO fallback: if the CUDA library is missing or no device is
visible, the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# VC_GPU_LIB: another build of the same ABI (A/B measurements); the default is the in-tree library
LIB_PATH = os.environ.get("VC_GPU_LIB") or os.path.join(HERE, "lib", "libverticut_gpu.so")

VC_OK, VC_NOT_FOUND = 0, 1
VC_ERR_ARG, VC_ERR_STATE, VC_ERR_CUDA, VC_ERR_NOMEM = -1, -2, -3, -4
EMPTY_KEY = 0xFFFFFFFFFFFFFFFF
EMPTY_U32 = 0xFFFFFFFF
MAX_K = 2048

# every symbol include/verticut_gpu.h declares (tests check the library exports all of them)
EXPORTS = [
    "vc_last_error", "vc_abi_version", "vc_device_count", "vc_index_create", "vc_index_destroy", "vc_index_get_info",
    "vc_index_add", "vc_index_add_device", "vc_index_add_synthetic", "vc_synth_word", "vc_index_build",
    "vc_bucket_get", "vc_code_get", "vc_occupancy_bitmap_get", "vc_search_linear", "vc_search_mih",
    "vc_search_linear_dev", "vc_search_mih_dev", "vc_merge_topk_dev", "vc_merge_topk",
    "vc_index_set_param", "vc_index_get_param", "vc_index_set_allreduce", "vc_index_save", "vc_index_load",
    "vc_nccl_allreduce_hook", "vc_xchg_create", "vc_xchg_local_window", "vc_xchg_open", "vc_xchg_open_ptrs", "vc_search_sharded_dev",
]
XCHG_HANDLE_BYTES = 64


class VerticutError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("verticut_gpu error %d: %s" % (code, msg))
        self.code = code


class QueryStats(C.Structure):
    _fields_ = [("radius", C.c_uint32), ("n_results", C.c_uint32), ("probes", C.c_uint64),
                ("occupancy_tests", C.c_uint64), ("candidates", C.c_uint64), ("unique", C.c_uint64)]


STATS_DTYPE = np.dtype([("radius", np.uint32), ("n_results", np.uint32), ("probes", np.uint64),
                        ("occupancy_tests", np.uint64), ("candidates", np.uint64), ("unique", np.uint64)])


class NcclHook(C.Structure):          # vc_nccl_hook
    _fields_ = [("nccl_allreduce", C.c_void_p), ("comm", C.c_void_p)]


class IndexInfo(C.Structure):
    _fields_ = [("n_codes", C.c_uint64), ("code_bits", C.c_uint32), ("n_tables", C.c_uint32),
                ("substring_bits", C.c_uint32), ("first_id", C.c_uint32), ("device", C.c_int32), ("built", C.c_int32),
                ("device_bytes", C.c_uint64)]


# int fn(void* user, uint32_t* d_words, uint64_t n_words, void* stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)

_lib = None


def lib():
    """Loads the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing - build it with `make -C verticut_b200/csrc` (or __graft_entry__.build()); "
                          "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.vc_last_error.restype = C.c_char_p
    L.vc_synth_word.restype = C.c_uint64
    L.vc_synth_word.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.vc_index_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.vc_index_destroy.argtypes = [vp]
    L.vc_index_destroy.restype = None
    L.vc_index_get_info.argtypes = [vp, C.POINTER(IndexInfo)]
    L.vc_index_add.argtypes = [vp, vp, C.c_uint64]
    L.vc_index_add_device.argtypes = [vp, vp, C.c_uint64]
    L.vc_index_add_synthetic.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.vc_index_build.argtypes = [vp]
    L.vc_index_save.argtypes = [vp, C.c_char_p]
    L.vc_index_load.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.vc_bucket_get.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, u32p]
    L.vc_code_get.argtypes = [vp, C.c_uint32, vp]
    L.vc_occupancy_bitmap_get.argtypes = [vp, C.c_uint32, vp, C.c_uint64]
    L.vc_search_linear.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.vc_search_mih.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp, vp, vp, vp]
    L.vc_search_linear_dev.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, vp]
    L.vc_search_mih_dev.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp, vp, vp]
    L.vc_merge_topk_dev.argtypes = [C.c_int, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    L.vc_merge_topk.argtypes = [C.c_int, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.vc_index_set_allreduce.argtypes = [vp, ALLREDUCE_FN, vp]
    L.vc_nccl_allreduce_hook.argtypes = [vp, vp, C.c_uint64, vp]
    L.vc_xchg_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint64, vp]
    L.vc_xchg_local_window.argtypes = [vp, C.POINTER(vp)]
    L.vc_xchg_open.argtypes = [vp, vp]
    L.vc_xchg_open_ptrs.argtypes = [vp, C.POINTER(vp)]
    L.vc_search_sharded_dev.argtypes = [vp, C.c_int, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp, vp]
    L.vc_index_set_param.argtypes = [vp, C.c_char_p, C.c_int64]
    L.vc_index_get_param.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
    _lib = L
    return L


def check(rc, allow=(VC_OK,)):
    if rc not in allow:
        raise VerticutError(rc, lib().vc_last_error().decode("utf-8", "replace"))
    return rc


def device_count():
    return check(lib().vc_device_count(), allow=tuple(range(0, 1024)))


def synth_word(seed, idx, word):
    return lib().vc_synth_word(seed, idx, word)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Index:
    """One id-shard of the code database, resident in the HBM of one GPU (vc_index)."""

    def __init__(self, code_bits, n_tables, device=0, first_id=0):
        self.h = C.c_void_p()
        self.code_bits, self.n_tables, self.device, self.first_id = code_bits, n_tables, device, first_id
        self.nbytes = code_bits // 8
        check(lib().vc_index_create(device, code_bits, n_tables, first_id, C.byref(self.h)))

    def save(self, path):
        check(lib().vc_index_save(self.h, os.fsencode(path)))

    @classmethod
    def load(cls, path, device=0):
        """Reads an index written by save(); the tables come back built."""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        check(lib().vc_index_load(device, os.fsencode(path), C.byref(self.h)))
        inf = self.info()
        self.code_bits, self.n_tables, self.device, self.first_id = inf["code_bits"], inf["n_tables"], device, inf["first_id"]
        self.nbytes = self.code_bits // 8
        return self

    def close(self):
        if self.h:
            lib().vc_index_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- contents ------------------------------------------------------------------------------------
    def add(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1, self.nbytes)
        check(lib().vc_index_add(self.h, _ptr(codes), codes.shape[0]))

    def add_device(self, dev_ptr, n):
        check(lib().vc_index_add_device(self.h, C.c_void_p(dev_ptr), n))

    def add_synthetic(self, n, seed):
        check(lib().vc_index_add_synthetic(self.h, n, seed))

    def build(self):
        check(lib().vc_index_build(self.h))

    def info(self):
        inf = IndexInfo()
        check(lib().vc_index_get_info(self.h, C.byref(inf)))
        return {f: getattr(inf, f) for f, _ in IndexInfo._fields_}

    def set_allreduce(self, fn):
        """fn(device_ptr, n_words, stream) -> None must sum the n_words uint32 words at device_ptr over all shards in
        place (see vc_index_set_allreduce); fn = None removes the hook."""
        if fn is None:
            self._allreduce_cb = None
            self._nccl_hook = None
            check(lib().vc_index_set_allreduce(self.h, C.cast(None, ALLREDUCE_FN), None))
            return

        def tramp(user, ptr, n_words, stream):
            try:
                fn(ptr, n_words, stream)
                return 0
            except Exception:          # an exception must not unwind through the C frames
                import traceback
                traceback.print_exc()
                return 1

        self._allreduce_cb = ALLREDUCE_FN(tramp)      # keep the trampoline alive as long as the index
        check(lib().vc_index_set_allreduce(self.h, self._allreduce_cb, None))

    def set_allreduce_nccl(self, nccl_allreduce_addr, comm):
        """The per-step exchanges go straight from the library to ncclAllReduce (vc_nccl_allreduce_hook): address of
        ncclAllReduce in the loaded NCCL library and an ncclComm_t of this rank, both as integers."""
        self._nccl_hook = NcclHook(C.c_void_p(nccl_allreduce_addr), C.c_void_p(comm))      # must outlive the index's use of it
        self._allreduce_cb = None
        fn = C.cast(lib().vc_nccl_allreduce_hook, ALLREDUCE_FN)
        check(lib().vc_index_set_allreduce(self.h, fn, C.cast(C.pointer(self._nccl_hook), C.c_void_p)))

    # ---- peer exchange over NVLink peer memory (vc_xchg_*) ------------------------------------------------
    def xchg_create(self, rank, world, slot_bytes):
        """Allocates this rank's exchange window; returns its CUDA IPC handle (bytes) for the other processes."""
        h = (C.c_ubyte * XCHG_HANDLE_BYTES)()
        check(lib().vc_xchg_create(self.h, rank, world, slot_bytes, C.cast(h, C.c_void_p)))
        return bytes(h)

    def xchg_local_window(self):
        p = C.c_void_p()
        check(lib().vc_xchg_local_window(self.h, C.byref(p)))
        return p.value

    def xchg_open(self, handles):
        """handles: the IPC handles of all ranks back to back (world x 64 bytes, rank order)."""
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        check(lib().vc_xchg_open(self.h, C.cast(buf, C.c_void_p)))

    def xchg_open_ptrs(self, windows):
        """windows: device pointers of all ranks' windows (all shards in one process)."""
        arr = (C.c_void_p * len(windows))(*[C.c_void_p(w) for w in windows])
        check(lib().vc_xchg_open_ptrs(self.h, arr))

    def search_sharded_dev(self, mih, d_queries, nq, k, d_out_keys, approximate=False, max_radius=-1, stream=0):
        check(lib().vc_search_sharded_dev(self.h, int(bool(mih)), C.c_void_p(d_queries), nq, k, int(approximate), int(max_radius),
                                          C.c_void_p(d_out_keys), C.c_void_p(stream)))

    def set_param(self, name, value):
        check(lib().vc_index_set_param(self.h, name.encode(), int(value)))

    def get_param(self, name):
        v = C.c_int64(0)
        check(lib().vc_index_get_param(self.h, name.encode(), C.byref(v)))
        return v.value

    # ---- BaseProxy::get ------------------------------------------------------------------------------
    def bucket_get(self, table, index, cap=None):
        """Returns (rc, ids, codes): rc 0 = found, 1 = not found (src/base_proxy.h:10-11)."""
        n = C.c_uint32(0)
        rc = check(lib().vc_bucket_get(self.h, table, index, None, None, 0, C.byref(n)), allow=(VC_OK, VC_NOT_FOUND))
        if rc == VC_NOT_FOUND:
            return rc, np.empty(0, np.uint32), np.empty((0, self.nbytes), np.uint8)
        cap = n.value if cap is None else min(cap, n.value)
        ids = np.empty(cap, np.uint32)
        codes = np.empty((cap, self.nbytes), np.uint8)
        rc = check(lib().vc_bucket_get(self.h, table, index, _ptr(ids), _ptr(codes), cap, C.byref(n)), allow=(VC_OK, VC_NOT_FOUND))
        return rc, ids, codes

    def code_get(self, idx):
        code = np.empty(self.nbytes, np.uint8)
        rc = check(lib().vc_code_get(self.h, idx, _ptr(code)), allow=(VC_OK, VC_NOT_FOUND))
        return rc, code

    def occupancy_bitmap(self, table):
        sbits = self.code_bits // self.n_tables
        words = np.empty((1 << sbits) // 32, np.uint32)
        check(lib().vc_occupancy_bitmap_get(self.h, table, _ptr(words), words.size))
        return words

    # ---- search, host buffers ---------------------------------------------------------------------------
    def _out(self, nq, k):
        return (np.empty((nq, k), np.uint32), np.empty((nq, k), np.uint32), np.empty(nq, np.uint32))

    def search_linear(self, queries, k):
        queries = np.ascontiguousarray(queries, dtype=np.uint8).reshape(-1, self.nbytes)
        nq = queries.shape[0]
        ids, dists, counts = self._out(nq, k)
        check(lib().vc_search_linear(self.h, _ptr(queries), nq, k, _ptr(ids), _ptr(dists), _ptr(counts)))
        return ids, dists, counts

    def search_mih(self, queries, k, approximate=False, max_radius=-1, with_stats=True):
        queries = np.ascontiguousarray(queries, dtype=np.uint8).reshape(-1, self.nbytes)
        nq = queries.shape[0]
        ids, dists, counts = self._out(nq, k)
        stats = np.zeros(nq, STATS_DTYPE) if with_stats else None
        check(lib().vc_search_mih(self.h, _ptr(queries), nq, k, int(approximate), int(max_radius), _ptr(ids), _ptr(dists),
                                  _ptr(counts), _ptr(stats)))
        return ids, dists, counts, stats

    # ---- search, device pointers (raw addresses, e.g. torch.Tensor.data_ptr()) ---------------------------
    def search_linear_dev(self, d_queries, nq, k, d_out_keys, stream=0):
        check(lib().vc_search_linear_dev(self.h, C.c_void_p(d_queries), nq, k, C.c_void_p(d_out_keys), C.c_void_p(stream)))

    def search_mih_dev(self, d_queries, nq, k, d_out_keys, approximate=False, max_radius=-1, d_stats=0, stream=0):
        check(lib().vc_search_mih_dev(self.h, C.c_void_p(d_queries), nq, k, int(approximate), int(max_radius),
                                      C.c_void_p(d_out_keys), C.c_void_p(d_stats) if d_stats else None, C.c_void_p(stream)))


def merge_topk_dev(device, d_lists, n_lists, nq, k, d_out, stream=0):
    check(lib().vc_merge_topk_dev(device, C.c_void_p(d_lists), n_lists, nq, k, C.c_void_p(d_out), C.c_void_p(stream)))


def merge_topk(device, lists, k):
    """lists: [n_lists, nq, k] uint64 packed words -> [nq, k] merged."""
    lists = np.ascontiguousarray(lists, dtype=np.uint64)
    n_lists, nq, kk = lists.shape
    assert kk == k
    out = np.empty((nq, k), np.uint64)
    check(lib().vc_merge_topk(device, _ptr(lists), n_lists, nq, k, _ptr(out)))
    return out


def unpack_keys(keys):
    """[.., k] uint64 packed words -> (ids, dists, counts) with EMPTY_U32 padding."""
    keys = np.asarray(keys, dtype=np.uint64)
    valid = keys != np.uint64(EMPTY_KEY)
    ids = np.where(valid, keys & np.uint64(0xFFFFFFFF), np.uint64(EMPTY_U32)).astype(np.uint32)
    dists = np.where(valid, keys >> np.uint64(32), np.uint64(EMPTY_U32)).astype(np.uint32)
    return ids, dists, valid.sum(axis=-1).astype(np.uint32)
