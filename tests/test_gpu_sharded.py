"""P5 on one GPU: the database cut into G id-shards (G index objects), per-shard search, merge kernel ==
the unsharded search == the oracle.  (The NCCL all-gather itself is covered by bench.py --gpus N and by the
gloo test of the plumbing.)"""
import os

import numpy as np
import pytest

from verticut_b200 import capi
from verticut_b200.sharded import shard_interleaved, shard_range

pytestmark = pytest.mark.gpu


def test_python_hook_on_the_default_stream_equals_the_in_library_emulation():
    """A Python all-reduce callback that multiplies by G on the library's stream (handle 0 = the legacy default stream) must give
    exactly what the library's own "times G" (`xchg.emulate`, tools/probe.py shards=G) gives - the work the callback enqueues is
    ordered with the library's kernels.  (torch.cuda.ExternalStream(0) is not; sharded.py maps handle 0 to torch's default stream.)"""
    import torch
    G, n, nq, k = 8, 3_000_000, 512, 100
    queries = np.random.default_rng(5).integers(0, 256, size=(nq, 8), dtype=np.uint8)
    res = {}
    for how in ("library", "python"):
        ix = capi.Index(64, 4)
        ix.set_param("id_stride", G)
        ix.add_synthetic(n, 12345)
        ix.build()
        ix.set_param("mih.batched", 1)
        calls = []
        if how == "library":
            ix.set_param("xchg.emulate", G)
        else:
            def times_g(ptr, n_words, stream):
                class _Raw:
                    __cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
                view = torch.as_tensor(_Raw(), device=torch.device("cuda", 0))
                handle = int(stream or 0)
                target = torch.cuda.default_stream() if handle == 0 else torch.cuda.ExternalStream(handle)
                with torch.cuda.stream(target):
                    view.mul_(G)
                calls.append(n_words)
            ix.set_allreduce(times_g)
        for _ in range(2):
            res[how] = ix.search_mih(queries, k)
        res[how + ".levels"] = ix.get_param("mih.last_levels")
        assert how == "library" or calls
        ix.close()
    assert res["library.levels"] == res["python.levels"]
    for a, b in zip(res["library"][:3], res["python"][:3]):
        np.testing.assert_array_equal(a, b)
    for f in ("radius", "probes", "candidates"):
        np.testing.assert_array_equal(res["library"][3][f], res["python"][3][f])


@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("mode", ["linear", "mih"])
def test_sharded_equals_unsharded(oracle, G, mode):
    n, bits, m, nq, k = 60_000, 64, 4, 12, 100
    codes = oracle.synth_codes(12345, 0, n, 8)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    lists = []
    for g in range(G):
        b, e = shard_range(n, G, g)
        ix = capi.Index(bits, m, first_id=b)
        ix.add(codes[b:e])
        ix.build()
        if mode == "linear":
            ids, dists, counts = ix.search_linear(queries, k)
        else:
            ids, dists, counts, _ = ix.search_mih(queries, k)
        keys = (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
        keys[np.arange(k)[None, :] >= counts[:, None]] = np.uint64(capi.EMPTY_KEY)
        lists.append(keys)
        ix.close()
    merged = capi.merge_topk(0, np.stack(lists), k)
    ids, dists, counts = capi.unpack_keys(merged)
    oid, od, oc = oracle.linear_search(codes, queries, k)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    np.testing.assert_array_equal(counts, oc)


def test_global_threshold_hook_keeps_results_exact(oracle):
    """The all-reduce hook (vc_index_set_allreduce) lets a shard stop on the k-th distance of the whole database.
    Emulated on one GPU: shard B is searched first; while shard A is searched, the hook adds B's final distance
    histogram (what rank B would contribute).  A must still return every member of the global top-k it owns."""
    import ctypes as C
    import torch
    n, nq, k = 400_000, 40, 100
    codes = oracle.synth_codes(12345, 0, n, 8)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    half = n // 2
    keys = {}
    ixs = {}
    for name, b, e in (("B", half, n), ("A", 0, half)):
        ix = capi.Index(64, 4, first_id=b)
        ix.add(codes[b:e])
        ix.build()
        ix.set_param("mih.batched", 1)
        ixs[name] = ix
    ids, dists, counts, _ = ixs["B"].search_mih(queries, k)
    HB = 96
    hist_b = np.zeros((nq, HB), dtype=np.int32)
    for q in range(nq):
        for d in dists[q, : counts[q]]:
            hist_b[q, d] += 1
    d_hist_b = torch.from_numpy(hist_b).cuda()
    calls = []

    def fake_allreduce(ptr, n_words, stream):
        class _Raw:
            __cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
        view = torch.as_tensor(_Raw(), device="cuda")
        assert n_words == nq * HB
        view += d_hist_b.view(-1)
        torch.cuda.synchronize()
        calls.append(n_words)

    ixs["A"].set_allreduce(fake_allreduce)
    ixs["A"].set_param("mih.global_key", 0)     # the second exchange (id histograms) is covered by the two-thread test below
    ida, da, ca, sta = ixs["A"].search_mih(queries, k)
    assert calls, "the hook was never called"
    ixs["A"].set_allreduce(None)
    idl, dl, cl, stl = ixs["A"].search_mih(queries, k)          # local rule, for comparison
    assert sta["probes"].sum() <= stl["probes"].sum()
    pack = lambda i, d, c: np.where(np.arange(k)[None, :] < c[:, None], (d.astype(np.uint64) << np.uint64(32)) | i.astype(np.uint64), np.uint64(capi.EMPTY_KEY))
    merged = capi.merge_topk(0, np.stack([pack(ida, da, ca), pack(ids, dists, counts)]), k)
    mi, md, mc = capi.unpack_keys(merged)
    oid, od, oc = oracle.linear_search(codes, queries, k)
    np.testing.assert_array_equal(mi, oid)
    np.testing.assert_array_equal(md, od)
    for ix in ixs.values():
        ix.close()


@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("mode", ["linear", "mih"])
def test_interleaved_shards_equal_unsharded(oracle, G, mode):
    """Ids dealt round-robin over the shards (first_id = g, id_stride = G): same answers, and vc_code_get / the synthetic
    generator follow the same id map."""
    n, bits, m, nq, k = 60_001, 64, 4, 12, 100
    codes = oracle.synth_codes(12345, 0, n, 8)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    lists = []
    for g in range(G):
        first, stride, cnt = shard_interleaved(n, G, g)
        assert (first, stride, cnt) == (g, G, len(range(g, n, G)))
        ix = capi.Index(bits, m, first_id=first)
        ix.set_param("id_stride", stride)
        if g % 2:
            ix.add(codes[g::G])
        else:
            ix.add_synthetic(cnt, 12345)             # generated on the device from the global ids
        ix.build()
        rc, code = ix.code_get(g + 5 * G)
        assert rc == 0 and (code == codes[g + 5 * G]).all()
        assert ix.code_get(g + 5 * G + 1)[0] == (1 if G > 1 else 0)
        if mode == "linear":
            ids, dists, counts = ix.search_linear(queries, k)
        else:
            ix.set_param("mih.batched", g % 2)
            ids, dists, counts, _ = ix.search_mih(queries, k)
        assert ((ids[:, :1] - g) % G == 0).all()
        keys = (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
        keys[np.arange(k)[None, :] >= counts[:, None]] = np.uint64(capi.EMPTY_KEY)
        lists.append(keys)
        ix.close()
    merged = capi.merge_topk(0, np.stack(lists), k)
    ids, dists, counts = capi.unpack_keys(merged)
    oid, od, oc = oracle.linear_search(codes, queries, k)
    np.testing.assert_array_equal(ids, oid)
    np.testing.assert_array_equal(dists, od)
    np.testing.assert_array_equal(counts, oc)


@pytest.mark.parametrize("table_steps,spec_tau", [(-1, -1), (1, -1), (-1, 16), (1, 5)])
def test_two_shards_with_live_exchange(oracle, table_steps, spec_tau):
    """Both collectives of the sharded batched search, matched between two shards that search at the same time (two
    host threads, one GPU): the per-step sum of the distance histograms and, before table-granular steps, the sum of the
    id histograms that bounds the k-th key of the whole database.  Merged result == oracle, exactly."""
    import threading
    import torch
    n, nq, k, G = 3_000_000, 64, 100, 2
    codes = oracle.synth_codes(12345, 0, n, 8)
    queries = oracle.synth_codes(67890, 0, nq, 8)
    ixs, views, results, errors = [], [None] * G, [None] * G, []
    barrier = threading.Barrier(G)
    n_calls = [0] * G
    for g in range(G):
        first, stride, cnt = shard_interleaved(n, G, g)
        ix = capi.Index(64, 4, first_id=first)
        ix.set_param("id_stride", stride)
        ix.add_synthetic(cnt, 12345)
        ix.build()
        ix.set_param("mih.batched", 1)
        ix.set_param("mih.table_steps", table_steps)
        # a speculative threshold that is too small for some (16) or all (5) queries: the miss is decided on the SUMMED histograms,
        # so both shards take the query out at the same step and redo it; the merged answer stays exact
        ix.set_param("mih.spec_tau", spec_tau)
        ixs.append(ix)

    def make_hook(g):
        def hook(ptr, n_words, stream):
            class _Raw:
                __cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
            views[g] = torch.as_tensor(_Raw(), device="cuda")
            torch.cuda.synchronize()
            barrier.wait(timeout=60)
            if g == 0:
                total = views[0] + views[1]
                views[0].copy_(total)
                views[1].copy_(total)
                torch.cuda.synchronize()
            barrier.wait(timeout=60)
            n_calls[g] += 1
        return hook

    def run(g):
        try:
            ixs[g].set_allreduce(make_hook(g))
            results[g] = ixs[g].search_mih(queries, k)
        except Exception as ex:      # a thread that dies must not leave the other one waiting
            errors.append(ex)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(g,)) for g in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert n_calls[0] == n_calls[1] and n_calls[0] >= 2
    lists = []
    for g in range(G):
        ids, dists, counts, _ = results[g]
        keys = (dists.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
        keys[np.arange(k)[None, :] >= counts[:, None]] = np.uint64(capi.EMPTY_KEY)
        lists.append(keys)
        ixs[g].close()
    mi, md, mc = capi.unpack_keys(capi.merge_topk(0, np.stack(lists), k))
    oid, od, oc = oracle.linear_search(codes, queries, k)
    np.testing.assert_array_equal(mi, oid)
    np.testing.assert_array_equal(md, od)


def _peer_rank(rank, G, n, nq, k, conns, ret):
    """One id-shard in its own process, on GPU 0: window handles travel through pipes, searches through vc_search_sharded_dev."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from oracle import restatement as R
    from verticut_b200 import capi as C2
    queries = R.synth_codes(67890, 0, nq, 8)
    dq = torch.from_numpy(queries).cuda()
    first, stride, cnt = shard_interleaved(n, G, rank)
    ix = C2.Index(64, 4, first_id=first)
    ix.set_param("id_stride", stride)
    ix.add_synthetic(cnt, 12345)
    ix.build()
    ix.set_param("mih.batched", 1)
    ix.set_param("mih.speculate", 2)                    # learned speculative thresholds for these small batches too (the default: >= 1024 queries)
    mine = ix.xchg_create(rank, G, 4 << 20)
    for c in conns:                                     # full exchange of the 64-byte IPC handles
        c.send((rank, mine))
    handles = {rank: mine}
    for c in conns:
        r, h = c.recv()
        handles[r] = h
    ix.xchg_open(b"".join(handles[r] for r in range(G)))
    outs = [torch.empty((nq, k), dtype=torch.int64, device="cuda") for _ in range(4)]
    ix.search_sharded_dev(True, dq.data_ptr(), nq, k, outs[0].data_ptr())
    n_x = ix.get_param("xchg.last")
    ix.search_sharded_dev(False, dq.data_ptr(), nq, k, outs[1].data_ptr())
    ix.search_sharded_dev(True, dq.data_ptr(), nq, k, outs[2].data_ptr(), max_radius=1)
    # other queries: this search starts from the guess the first one taught (the largest k-th distance of the WHOLE database's
    # answers - both ranks learn the same value from the summed histograms); queries it is too small for are redone on both ranks
    dq2 = torch.from_numpy(R.synth_codes(24680, 0, nq, 8)).cuda()
    ix.search_sharded_dev(True, dq2.data_ptr(), nq, k, outs[3].data_ptr())
    spec = (ix.get_param("mih.last_spec_tau"), ix.get_param("mih.last_spec_fail"))
    torch.cuda.synchronize()
    ret[rank] = ([o.cpu().numpy().view(np.uint64).copy() for o in outs], n_x, spec)
    for c in conns:                                     # nobody closes its window while a peer may still store into it
        c.send("done")
    for c in conns:
        c.recv()
    ix.close()


def test_two_shards_over_peer_windows(oracle):
    """The peer-memory exchange (verticut_b200/csrc/xchg.cuh) between two shards that live in two PROCESSES sharing one GPU: each
    rank opens the other's window through its CUDA IPC handle (vc_xchg_open) - the mechanism of the multi-GPU path, minus
    NVLink -; the settle kernel's histogram rows and the finish kernel's top-k rows go through the windows, and
    vc_search_sharded_dev returns the merged answer on both ranks.  Merged result == oracle, exactly, for the MIH search, the
    linear scan and a fixed-radius search.  (Two contexts on one GPU are time-sliced, so a rank's wait kernel spins through its
    slice until the peer's slice has produced - slow, which is fine here; tests/test_gpu_nccl.py is the same over two GPUs.)"""
    import multiprocessing as mp
    n, nq, k, G = 1_200_000, 48, 100, 2
    queries = oracle.synth_codes(67890, 0, nq, 8)
    ctx = mp.get_context("spawn")
    a, b = ctx.Pipe()
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_peer_rank, args=(0, G, n, nq, k, [a], ret)), ctx.Process(target=_peer_rank, args=(1, G, n, nq, k, [b], ret))]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    for p in procs:
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0, "rank process failed (exit code %s)" % p.exitcode
    want = oracle.scan_synth(12345, 0, 1, n, 8, queries, k, n_procs=4)
    want_r1 = oracle.scan_synth(12345, 0, 1, n, 8, queries, k, m=4, max_radius=1, n_procs=4)
    want2 = oracle.scan_synth(12345, 0, 1, n, 8, oracle.synth_codes(24680, 0, nq, 8), k, n_procs=4)
    kth1 = int((want[:, k - 1] >> np.uint64(32)).max())
    misses = int(((want2[:, k - 1] >> np.uint64(32)) > kth1).sum())
    for g in range(G):
        outs, n_x, spec = ret[g]
        assert n_x >= 3                                  # bootstrap + one per step + the result rows
        np.testing.assert_array_equal(outs[0], want)
        np.testing.assert_array_equal(outs[1], want)
        np.testing.assert_array_equal(outs[2], want_r1)
        assert spec == (kth1, misses), (spec, kth1, misses)      # the same guess and the same misses on both ranks
        np.testing.assert_array_equal(outs[3], want2)
