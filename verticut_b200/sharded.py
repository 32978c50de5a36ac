"""Id-sharded search over the GPUs of one box: one process per GPU, NCCL all-gather of the local top-k.

Replaces the reference's distribution layer for this path - `mpirun -n m` with one rank per TABLE
(src/search_worker.cc:58,99-101,225), MPI_Gatherv of every candidate to rank 0 and MPI_Bcast of the stop
flag once per radius step (src/mpi_coordinator.cc:26-69, src/search_worker.cc:177,207) - by:

  * rank g owns an id-shard - a contiguous range (shard_range) or, what bench.py uses, every G-th id
    (shard_interleaved) - and ALL m tables over it;
  * queries are replicated; every rank answers them on its shard (exact local top-k);
  * ONE all-gather of [nq][k] packed words (dist<<32|id) per batch over NVLink, then every rank folds
    the G lists with merge_topk_kernel (shards are id-disjoint, so no de-duplication is needed);
  * inside the MIH search, one small all-reduce per search step sums the per-query distance histograms of
    the shards (vc_index_set_allreduce), so that every GPU filters and stops on the k-th distance of the
    WHOLE database: a shard then does 1/G of the single-GPU work instead of searching to its own, larger,
    local k-th distance.  (The reference exchanges all candidates and a stop flag per radius step.)
  * on NVLink-connected GPUs neither exchange goes through NCCL: every rank opens the others' exchange windows (CUDA IPC,
    handles carried by torch.distributed) and the search kernels store their histogram rows and top-k rows straight into
    the peers' memory (verticut_b200/csrc/xchg.cuh; one call, Index.search_sharded_dev, = search + exchange + merge).
    VC_XCHG=0 turns that off; the exchanges then go to ncclAllReduce issued by the library from C on the search stream
    (vc_nccl_allreduce_hook, a communicator created here; VC_NCCL_DIRECT=0: torch.distributed.all_reduce through a Python
    callback), and the results to torch's all_gather_into_tensor + the merge kernel.  Payloads larger than a window slot
    take that route too.

torch.distributed is plumbing here (process group, the all-gather, device buffers); searching and
merging are the library's CUDA kernels, called through the C ABI with raw device pointers.
"""
import os

import numpy as np

from . import capi


def shard_range(n_total, world_size, rank):
    """[begin, end) of the ids owned by `rank`: contiguous, sizes differ by at most one code."""
    base, rem = divmod(int(n_total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_interleaved(n_total, world_size, rank):
    """(first_id, id_stride, count) of the ids owned by `rank` when ids are dealt round-robin: rank, rank + G, ...
    Every shard then spans the whole id range, which is what lets all of them leave a bucket at the k-th id of the whole
    database (bmih.cuh) instead of only the shards that own small ids; create the index with first_id = rank and
    set_param("id_stride", G) before adding codes."""
    n_total, world_size, rank = int(n_total), int(world_size), int(rank)
    count = (n_total - rank + world_size - 1) // world_size if rank < n_total else 0
    return rank, world_size, count


def gather_layout(world_size, nq, k):
    """Shape of the all-gathered buffer the merge kernel consumes: [n_lists = G][nq][k] packed words."""
    return (world_size, nq, k)


class ShardedSearcher:
    """One rank's view of the sharded database.

    `index` is this rank's capi.Index (first_id = start of its id range).  The two device steps are methods so
    that the host-side plumbing (shard ranges, gather layout, collective) can be exercised on CPU with the
    `gloo` backend by overriding them; the product path always runs the CUDA kernels.
    """

    def __init__(self, index, group=None, global_threshold=True):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.index = index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._bufs = {}
        self._views = {}
        self._nccl = None
        self.peer_windows = False
        self.exchange = "none"
        if self.world > 1 and global_threshold and hasattr(index, "set_allreduce"):
            index.set_allreduce(self._allreduce_words)
            self.exchange = "torch.distributed.all_reduce (Python callback)"
            if os.environ.get("VC_NCCL_DIRECT", "1") != "0" and hasattr(index, "set_allreduce_nccl") and dist.get_backend(group) == "nccl":
                self._nccl_direct()
                self.exchange = "ncclAllReduce from C on the search stream"
            if os.environ.get("VC_XCHG", "1") != "0" and hasattr(index, "xchg_create") and dist.get_backend(group) == "nccl":
                self._open_peer_windows()

    def _nccl_direct(self):
        """The library's per-step exchanges call ncclAllReduce themselves (vc_nccl_allreduce_hook), on the search stream, on a
        communicator of their own, instead of coming back to Python for torch.distributed.all_reduce - half a dozen callbacks
        per search.  torch.distributed stays the plumbing: it carries the ncclUniqueId to the other ranks."""
        import ctypes as C
        t, dist = self.torch, self.dist
        nccl = C.CDLL("libnccl.so.2")                     # by soname: the copy torch has already loaded

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]

        uid = UniqueId()
        if self.rank == 0 and nccl.ncclGetUniqueId(C.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        dev = t.device("cuda", self.index.device)
        buf = t.tensor(list(bytes(uid)), dtype=t.uint8, device=dev)
        dist.broadcast(buf, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        C.memmove(C.byref(uid), bytes(buf.cpu().tolist()), 128)
        comm = C.c_void_p()
        nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        with t.cuda.device(dev):
            rc = nccl.ncclCommInitRank(C.byref(comm), self.world, uid, self.rank)
        if rc != 0:
            raise RuntimeError("ncclCommInitRank failed (ncclResult_t %d)" % rc)
        self._nccl = (nccl, comm)                          # keep the library handle and the communicator alive
        self.index.set_allreduce_nccl(C.cast(nccl.ncclAllReduce, C.c_void_p).value, comm.value)

    XCHG_RESULT_EGRESS = 8 << 20    # result rows go over peer memory while payload x peers stays below this
    XCHG_SLOT_BYTES = 32 << 20      # largest payload of one rank in one exchange that goes over peer memory (else: NCCL)

    def _open_peer_windows(self):
        """Exchange windows over NVLink peer memory: allocate mine, hand its IPC handle to the peers, open theirs.  If any rank
        cannot (no peer access between the devices), every rank stays on the NCCL path."""
        t, dist = self.torch, self.dist
        dev = t.device("cuda", self.index.device)
        ok, handle = 1, bytes(capi.XCHG_HANDLE_BYTES)
        try:
            handle = self.index.xchg_create(self.rank, self.world, self.XCHG_SLOT_BYTES)
        except capi.VerticutError:
            ok = 0
        mine = t.tensor(list(handle), dtype=t.uint8, device=dev)
        everyone = t.empty(self.world * capi.XCHG_HANDLE_BYTES, dtype=t.uint8, device=dev)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        flag = t.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()):
            try:
                self.index.xchg_open(bytes(everyone.cpu().tolist()))
            except capi.VerticutError:
                ok = 0
            flag = t.tensor([ok], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self.peer_windows = bool(int(flag.item()))
        if self.peer_windows:
            self.exchange = ("stores into the peers' memory over NVLink (xchg.cuh) while payload x peers <= 8 MB (result rows) / 2 MB "
                             "(histogram sums), else " + self.exchange + " and all_gather_into_tensor + merge kernel")
        else:
            self.index.set_param("xchg", 0)
        dist.barrier(group=self.group)

    def close(self):
        """Releases the library's own NCCL communicator (if any); the index is the caller's."""
        if self._nccl is not None:
            nccl, comm = self._nccl
            self.index.set_allreduce(None)
            nccl.ncclCommDestroy.argtypes = [type(comm)]
            nccl.ncclCommDestroy(comm)
            self._nccl = None

    def _allreduce_words(self, ptr, n_words, stream):
        """Sum n_words int32 words at device address ptr over the ranks, ordered after the work already in `stream` (the
        library's search stream) and before what the library enqueues there next."""
        t = self.torch
        key = (ptr, n_words)
        view = self._views.get(key)
        if view is None:
            class _Raw:            # zero-copy view of library-owned device memory
                __cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
            view = t.as_tensor(_Raw(), device=t.device("cuda", self.index.device))
            self._views[key] = view
        handle = int(stream or 0)
        if view.device.type == "cuda" and handle != t.cuda.current_stream(view.device).cuda_stream:
            # handle 0 is the legacy default stream = torch's default stream; ExternalStream(0) is NOT ordered with it (measured:
            # work launched under it overtook the library's kernels on stream 0)
            target = t.cuda.default_stream(view.device) if handle == 0 else t.cuda.ExternalStream(handle, device=view.device)
            with t.cuda.stream(target):
                self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group)
        else:
            self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM, group=self.group)

    def _buffers(self, nq, k, device):
        key = (nq, k, str(device))
        if key not in self._bufs:
            t = self.torch
            local = t.empty((nq, k), dtype=t.int64, device=device)
            gathered = t.empty(gather_layout(self.world, nq, k), dtype=t.int64, device=device)
            merged = t.empty((nq, k), dtype=t.int64, device=device)
            self._bufs[key] = (local, gathered, merged)
        return self._bufs[key]

    def _stream(self, device):
        return self.torch.cuda.current_stream().cuda_stream if device.type == "cuda" else 0

    # ---- device steps (CUDA kernels through the C ABI) --------------------------------------------------
    def local_search(self, d_queries, k, mode, approximate, max_radius, out_keys):
        nq = d_queries.shape[0]
        stream = self._stream(d_queries.device)
        if mode == "mih":
            self.index.search_mih_dev(d_queries.data_ptr(), nq, k, out_keys.data_ptr(), approximate=approximate,
                                      max_radius=max_radius, stream=stream)
        elif mode == "linear":
            self.index.search_linear_dev(d_queries.data_ptr(), nq, k, out_keys.data_ptr(), stream=stream)
        else:
            raise ValueError("mode must be 'mih' or 'linear'")

    def merge(self, gathered, k, out_keys):
        n_lists, nq, _ = gathered.shape
        capi.merge_topk_dev(self.index.device, gathered.data_ptr(), n_lists, nq, k, out_keys.data_ptr(),
                            stream=self._stream(gathered.device))

    # ---- the batch -----------------------------------------------------------------------------------------
    def search(self, d_queries, k, mode="mih", approximate=False, max_radius=-1):
        """d_queries: uint8 tensor [nq, code_bytes] on this rank's device.  Returns an int64 tensor [nq, k] of
        packed words (bit pattern of uint64, ascending), identical on every rank."""
        nq = d_queries.shape[0]
        local, gathered, merged = self._buffers(nq, k, d_queries.device)
        # Results over peer memory while the egress (payload x peers) is small - the latency-bound regime; large result sets go
        # through NCCL, whose all-gather is replicated inside the NVSwitch (measured at 8 GPUs, 13 MB per rank: 0.12 ms per
        # batch faster than the stores; level at 2 GPUs; profiles/scale_r02.json).  VC_XCHG_RESULTS=1 / 0 force.
        peer_results = os.environ.get("VC_XCHG_RESULTS", "")
        small = nq * k * 8 * (self.world - 1) <= self.XCHG_RESULT_EGRESS
        if (self.peer_windows and nq * k * 8 <= self.XCHG_SLOT_BYTES and mode in ("mih", "linear")
                and (peer_results == "1" or (peer_results != "0" and small))):
            # search + exchange + merge in one call, the exchange being the search kernels' own stores into the peers' windows
            self.index.search_sharded_dev(mode == "mih", d_queries.data_ptr(), nq, k, merged.data_ptr(), approximate=approximate,
                                          max_radius=max_radius, stream=self._stream(d_queries.device))
            return merged
        self.local_search(d_queries, k, mode, approximate, max_radius, local)
        if self.world == 1:
            return local
        self.dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        self.merge(gathered, k, merged)
        return merged
