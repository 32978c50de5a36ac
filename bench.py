#!/usr/bin/env python
"""bench.py - headline benchmark of verticut_b200 (see BASELINE.json).

Metric : exact k-NN queries/s at k = 100 over 1 B 64-bit codes (uniform synthetic), database sharded by id
         over the N GPUs of one box (strong scaling: total work fixed), plus scan HBM GB/s vs peak.
Step   : one batch of Q queries answered exactly by the MIH path (m = 4 tables) on every shard, followed
         (N > 1) by one NCCL all-gather of the local top-k and the merge kernel.
value  : whole-job queries/s with the query batch already resident in HBM.
e2e    : the same through the host-buffer call (vc_search_mih, pinned staging, H2D + D2H inside the timed region).
roofline: dominant kernel (bmih_verify_kernel, all launches of a search) - algorithmic bytes /
         CUDA-event time against the measured HBM peak of MEASURED_PEAKS.json; roofline.combined = the slower of HBM
         time and POPC time per search step, summed, over the measured time; integer_pipe = tests/s vs POPC peak.
breakdown: where a batch's time goes (CUDA events inside the library): verify kernels / probes + work items / settle +
         cross-shard exchange + decide / bootstrap + finish / the rest = host round trips, all-gather + merge.
scan   : the brute-force path (config C4): passes/s at small batches as HBM GB/s vs peak, queries/s at a large batch.
cpu_baseline: the reference's own linear scan (oracle/_ref, unmodified sources) on a bounded sample, host cores.
oracle_check: a few queries of the batch answered by the CPU oracle over the SAME (sharded) database at full size - every
         rank scans its own shard with oracle.restatement.scan_synth (codes generated on the fly), rank 0 merges the lists -
         and compared bit for bit with the merged GPU answer; outside the timed region, the oracle is only the checker.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config headline|C2|C3|C4|C5]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
--config (BASELINE.json `configs`): headline = 1 B x 64-bit MIH m=4 k=100 (the default, what the driver runs);
  C2 = MIH 64-bit m=4, 100 M codes, k=100;  C3 = MIH 128-bit m=8, 1 B codes, k=100 (needs the HBM of >= 2 GPUs; on one GPU a
  125 M-code shard, flagged as reduced);  C4 = brute-force scan of 1 B x 64-bit codes, batch sweep 1..4096;  C5 = MIH 256-bit,
  500 M codes (needs >= 4 GPUs; fewer: 62.5 M codes per GPU, flagged), fixed-radius sweep r = 0..3, k = 1000, m = 16
  (VC_BENCH_TABLES=8: the sparse s = 32 tables).
Environment overrides for quick runs: VC_BENCH_N (total codes), VC_BENCH_Q (batch), VC_BENCH_MODE (mih|linear), VC_BENCH_K,
VC_BENCH_BITS, VC_BENCH_TABLES, VC_BENCH_ORACLE_Q (queries checked against the oracle; 0 = skip).  VC_BENCH_WEAK=1 (N > 1): also time a
batch of N x the batch size (`batch_x_gpus`: the work per GPU of the one-GPU run).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line: whatever native libraries write to file descriptor 1 (NCCL_DEBUG=VERSION/INFO print
# "NCCL version ..." there) is sent to stderr, the JSON line goes to a private copy of the original stdout
RESULT_OUT = sys.stdout


def _private_stdout():
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DB_SEED, QUERY_SEED = 12345, 67890
POPC_PER_CLK_PER_SM = 15.4      # measured on this pool's B200s by tools/microbench.cu (profiles/microbench_r02.log); nominal 16

# workloads (BASELINE.json `configs`); n = total codes, Q = batch, radii = fixed-radius sweep (-1: the exact stop rule)
CONFIGS = {
    # batch: the search is bucket-stationary, so a larger batch puts more queries on every bucket read (7.5 per probed bucket at
    # radius 2 with 4096 queries, 30 with 16384) and amortises the per-step work; round 1 ran 4096 - kept as `batch_4096`
    "headline": dict(n=1_000_000_000, bits=64, m=4, k=100, Q=16384, mode="mih", radii=[-1], min_gpus=1),
    "C2": dict(n=100_000_000, bits=64, m=4, k=100, Q=4096, mode="mih", radii=[-1], min_gpus=1),
    "C3": dict(n=1_000_000_000, bits=128, m=8, k=100, Q=4096, mode="mih", radii=[-1], min_gpus=2, per_gpu=125_000_000),
    "C4": dict(n=1_000_000_000, bits=64, m=0, k=100, Q=4096, mode="linear", radii=[-1], min_gpus=1,
               sweep=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]),
    "C5": dict(n=500_000_000, bits=256, m=16, k=1000, Q=1024, mode="mih", radii=[0, 1, 2, 3], min_gpus=4, per_gpu=62_500_000),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU sampled through NVML while the timed region runs."""

    def __init__(self, gpu_index, period_s=0.05):   # a few samples per timed region (NVML queries can hold up launches: not more often)
        self.gpu, self.period = gpu_index, period_s
        self.samples, self.reasons = [], set()
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        self.err = None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            while not self._stop.is_set():
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for n, bit in names.items():
                        if mask & bit:
                            self.reasons.add(n)
                except Exception:
                    pass
                self._stop.wait(self.period)
        except Exception as ex:   # NVML missing: report it, never fake numbers
            self.err = repr(ex)

    def start(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        if self.err or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable: %s" % self.err], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_queries(seed, nq, nbytes):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(nq, nbytes), dtype=np.uint8)


def resolve_config(name, world):
    cfg = dict(CONFIGS[name])
    cfg["name"] = name
    env = os.environ
    cfg["reduced"] = None
    if world < cfg["min_gpus"] and "VC_BENCH_N" not in env:
        cfg["reduced"] = "%s needs the HBM of >= %d GPUs; this run holds %d codes per GPU instead (%d of %d codes)" % (
            name, cfg["min_gpus"], cfg["per_gpu"], cfg["per_gpu"] * world, cfg["n"])
        cfg["n"] = cfg["per_gpu"] * world
    cfg["n"] = int(env.get("VC_BENCH_N", cfg["n"]))
    cfg["Q"] = int(env.get("VC_BENCH_Q", cfg["Q"]))
    cfg["k"] = int(env.get("VC_BENCH_K", cfg["k"]))
    cfg["bits"] = int(env.get("VC_BENCH_BITS", cfg["bits"]))
    cfg["mode"] = env.get("VC_BENCH_MODE", cfg["mode"])
    if cfg["mode"] == "mih":
        cfg["m"] = int(env.get("VC_BENCH_TABLES", cfg["m"] or 4))
    if "VC_BENCH_RADII" in env:
        cfg["radii"] = [int(x) for x in env["VC_BENCH_RADII"].split(",")]
    return cfg


def metric_name(cfg):
    n = cfg["n"]
    return "queries/sec (k=%d, %s %d-bit codes)" % (cfg["k"], "1B" if n == 10**9 else str(n), cfg["bits"])


def reference_arm(args, cfg, world, rank):
    """--impl reference: the reference's own CPU linear scan (oracle/_ref) on a bounded sample."""
    if rank != 0:
        return
    from oracle import reference as ref, restatement as R
    cores = os.cpu_count() or 1
    n_total, bits, k = cfg["n"], cfg["bits"], cfg["k"]
    sample_n = int(os.environ.get("VC_BENCH_REF_N", 64_000_000 * 8 // (bits // 8)))
    nq = max(cores, 8) * 4
    codes = R.synth_codes(DB_SEED, 0, sample_n, bits // 8)
    mem = ref.RefMem(codes, 0)
    times = []
    for step in range(args.warmup + args.steps):
        q = make_queries(QUERY_SEED + step, nq, bits // 8)
        t0 = time.perf_counter()
        mem.linear_search(q, k, n_procs=cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    qps_sample = nq / per_step
    qps = qps_sample * sample_n / n_total          # the scan is linear in N
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 xor+popcount", "data": "synthetic",
        "config": {"workload": "exact k-NN, k=%d, %d x %d-bit uniform codes" % (k, n_total, bits)},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "reference",
                         "sample": "reference src/linear_search.cc:39-64 (unmodified, oracle/_ref) over an in-memory "
                                   "%d-code sample, %d queries/step on %d forked processes; scaled by sample/N (scan is linear in N)"
                                   % (sample_n, nq, cores)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    # the reference's MIH (SearchWorker::find, m threads as its m ranks) beside its scan, at the size it is feasible at on a
    # host: on uniform codes it is slower than its own scan already at 10^6 codes, which is why the
    # scan, not the MIH, is the CPU arm
    if not os.environ.get("VC_BENCH_SKIP_REF_MIH") and cfg["m"]:
        try:
            n_small = int(os.environ.get("VC_BENCH_REF_MIH_N", 1_000_000))
            small = codes[:n_small]
            qs = make_queries(QUERY_SEED, 8, bits // 8)
            rix = ref.RefMem(small, cfg["m"])                     # (replaces the in-memory store of the scan leg above)
            t0 = time.perf_counter()
            rix.mih_search(qs, k)
            dt_mih = (time.perf_counter() - t0) / qs.shape[0]
            t0 = time.perf_counter()
            rix.linear_search(qs, k, n_procs=1)
            dt_scan = (time.perf_counter() - t0) / qs.shape[0]
            line["reference_mih"] = {"n_codes": n_small, "queries_per_s": 1.0 / dt_mih, "threads": cfg["m"],
                                     "reference_scan_queries_per_s_1_core": 1.0 / dt_scan,
                                     "what": "reference SearchWorker::find (src/search_worker.cc:65-264, unmodified, oracle/_ref, one thread per "
                                             "table as its MPI ranks) on %d codes, 8 queries; not extrapolated - its cost is not linear in N" % n_small}
        except Exception as ex:
            line["reference_mih"] = {"unavailable": repr(ex)}
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def oracle_check(cfg, searcher, world, rank, dev, queries_np, shard, max_radius, n_check):
    """GPU answer (through the whole sharded path: local search, exchange, merge) of the first n_check queries vs the CPU oracle
    over the same database at full size.  Every rank scans its own shard; rank 0 merges and compares."""
    import torch
    import torch.distributed as dist
    from oracle import restatement as R
    first, stride, n_shard = shard
    k, nbytes, mode = cfg["k"], cfg["bits"] // 8, cfg["mode"]
    chk_np = np.ascontiguousarray(queries_np[:n_check])
    chk = torch.from_numpy(chk_np).to(dev)
    got = searcher.search(chk, k, mode=mode, max_radius=max_radius).clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    procs = max(1, (os.cpu_count() or 1) // world)
    m, r = (cfg["m"], max_radius) if (mode == "mih" and max_radius >= 0) else (0, -1)
    mine = R.scan_synth(DB_SEED, first, stride, n_shard, nbytes, chk_np, k, m=m, max_radius=r, n_procs=procs)
    if world > 1:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
    else:
        parts = [mine]
    res = None
    if rank == 0:
        want = np.empty((n_check, k), dtype=np.uint64)
        for q in range(n_check):
            mg = R.merge_topk(np.stack([p[q] for p in parts]), k)
            want[q, : len(mg)] = mg
            want[q, len(mg):] = np.iinfo(np.uint64).max
        ok = bool(np.array_equal(got.cpu().numpy().view(np.uint64), want))
        res = {"queries": int(n_check), "equals_oracle": ok, "oracle_s": time.perf_counter() - t0, "oracle_procs_per_rank": procs,
               "what": "merged GPU top-k (ids, distances, order) == oracle/restatement.scan_synth over all %d codes%s" % (
                   cfg["n"], "" if r < 0 else " reachable at radius <= %d (m=%d)" % (r, m))}
    if world > 1:
        flag = torch.tensor([1 if (res is None or res["equals_oracle"]) else 0], device=dev)
        dist.broadcast(flag, src=0)
        if int(flag.item()) == 0 and res is None:
            res = {"equals_oracle": False}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = resolve_config(args.config, world)
    n_total, Q, mode, K_NN, CODE_BITS, N_TABLES = cfg["n"], cfg["Q"], cfg["mode"], cfg["k"], cfg["bits"], cfg["m"]
    nbytes = CODE_BITS // 8

    if args.impl == "reference":
        reference_arm(args, cfg, world, rank)
        return

    import torch
    import torch.distributed as dist
    from verticut_b200 import capi
    from verticut_b200.sharded import ShardedSearcher, shard_interleaved

    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak_gbs, peak_src = measured_peaks()

    # ---- this rank's shard: synthetic codes generated on the device, all m tables built in HBM --------
    # ids are dealt round-robin over the ranks (rank, rank + G, ...): every shard spans the whole id range
    b, id_stride, n_shard = shard_interleaved(n_total, world, rank)
    t0 = time.perf_counter()
    ix = capi.Index(CODE_BITS, N_TABLES if mode == "mih" else 0, device=local_rank, first_id=b)
    ix.set_param("id_stride", id_stride)
    ix.add_synthetic(n_shard, DB_SEED)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    ix.build()
    t_build = time.perf_counter() - t0
    ix.set_param("profile", 1)
    searcher = ShardedSearcher(ix)
    fixed_radius = cfg["radii"] != [-1]

    n_batches = args.warmup + args.steps
    host_batches = [make_queries(QUERY_SEED + i, Q, nbytes) for i in range(n_batches)]      # fresh uniform codes
    pinned = [torch.from_numpy(hb).pin_memory() for hb in host_batches]
    dev_batches = [p.to(dev) for p in pinned]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            tt = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return x

    n_oracle_q = int(os.environ.get("VC_BENCH_ORACLE_Q", 4 if (args.config != "headline" or world > 1) else 2))
    sweep_rows = []
    headline = None
    for radius in (cfg["radii"] if "sweep" not in cfg else [-1]):
        def step(i, B=None):
            qd = dev_batches[i] if B is None else dev_batches[i][:B].contiguous()
            return searcher.search(qd, K_NN, mode=mode, max_radius=radius)

        if "sweep" in cfg:
            break
        for i in range(args.warmup):
            step(i)
        barrier()
        # ---- parity outside the timed region -------------------------------------------------------------------------
        ocheck = None
        if n_oracle_q > 0:
            ocheck = oracle_check(cfg, searcher, world, rank, dev, host_batches[0], (b, id_stride, n_shard), radius, min(n_oracle_q, Q))
            if ocheck is not None and not ocheck["equals_oracle"]:
                raise SystemExit("bench.py: the GPU answer differs from the CPU oracle - refusing to report a number")
        parity_ok = None
        if radius < 0 and mode == "mih":
            # self-check: the MIH answer of a few queries equals the brute-force scan of the same (sharded) database
            chk = dev_batches[0][:8].contiguous()
            a = searcher.search(chk, K_NN, mode="mih").clone()
            bscan = searcher.search(chk, K_NN, mode="linear").clone()
            parity_ok = bool(torch.equal(a, bscan))
            if not parity_ok:
                raise SystemExit("bench.py: MIH and linear-scan results differ - refusing to report a number")
        barrier()
        launches0 = ix.get_param("launches")
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ns, prof = [], []
        barrier()
        ev0.record()
        for i in range(args.warmup, n_batches):
            step(i)
            kernel_ns.append(ix.get_param("last_kernel_ns"))
            if mode == "mih" and ix.get_param("mih.last_batched"):
                L = ix.get_param("mih.last_levels")
                prof.append((ix.get_param("mih.search_ns"), sum(ix.get_param("mih.step_pre_ns.%d" % j) for j in range(L)),
                             sum(ix.get_param("mih.step_post_ns.%d" % j) for j in range(L))))
        ev1.record()
        barrier()
        clocks = sampler.stop()
        launches = ix.get_param("launches") - launches0
        ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
        value = Q / (ms_step * 1e-3)
        row = {"radius": radius, "queries_per_s": value, "ms_per_batch": ms_step, "verify_kernel_ms": float(np.mean(kernel_ns)) * 1e-6,
               "oracle_check": ocheck, "gpu_launches": int(launches)}
        breakdown = None
        if prof:
            s_ns, pre_ns, post_ns = (float(np.mean([p[j] for p in prof])) for j in range(3))
            v_ns = float(np.mean(kernel_ns))
            breakdown = {"unit": "ms per batch (this rank; CUDA events inside the library)", "batch_total": ms_step,
                         "verify_kernels": v_ns * 1e-6, "probes_and_work_items": pre_ns * 1e-6,
                         "settle_exchange_decide": post_ns * 1e-6,
                         "bootstrap_finish_and_host_round_trips_between_steps": (s_ns - v_ns - pre_ns - post_ns) * 1e-6,
                         "allgather_merge_and_call_overhead": ms_step - s_ns * 1e-6,
                         "exchange": searcher.exchange}
            row["breakdown"] = breakdown
        sweep_rows.append(row)
        headline = dict(value=value, ms_step=ms_step, kernel_ns=kernel_ns, clocks=clocks, launches=launches, parity_ok=parity_ok,
                        ocheck=ocheck, breakdown=breakdown, radius=radius)

    # ---- config C4: the brute-force scan, batch sweep ------------------------------------------------------------------
    if "sweep" in cfg:
        sampler = ClockSampler(local_rank)
        sampler.start()
        sm_hz = 1965.0e6
        popc_peak = POPC_PER_CLK_PER_SM * ix.get_param("num_sms") * sm_hz
        ocheck = None
        if n_oracle_q > 0:
            ocheck = oracle_check(cfg, searcher, world, rank, dev, host_batches[0], (b, id_stride, n_shard), -1, min(n_oracle_q, Q))
            if ocheck is not None and not ocheck["equals_oracle"]:
                raise SystemExit("bench.py: the GPU answer differs from the CPU oracle - refusing to report a number")
        total_launches = 0
        for B in cfg["sweep"]:
            if B > Q:
                continue
            qd = dev_batches[0][:B].contiguous()
            reps = max(1, min(args.steps, 5 if B <= 64 else 2))
            for _ in range(2 if B <= 64 else 1):
                searcher.search(qd, K_NN, mode="linear")
            barrier()
            l0 = ix.get_param("launches")
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ks = []
            s0.record()
            for _ in range(reps):
                searcher.search(qd, K_NN, mode="linear")
                ks.append(ix.get_param("last_kernel_ns"))
            s1.record()
            barrier()
            total_launches += ix.get_param("launches") - l0
            ms = max_over_ranks(s0.elapsed_time(s1) / reps)
            k_s = float(np.mean(ks)) * 1e-9
            t_hbm, t_popc = n_shard * nbytes / (peak_gbs * 1e9), B * n_shard * (nbytes // 8) / popc_peak
            sweep_rows.append({"batch": B, "queries_per_s": B / (ms * 1e-3), "ms_per_batch": ms, "scan_kernel_ms": k_s * 1e3,
                               "hbm_GBps_per_gpu": n_shard * nbytes / k_s / 1e9, "frac_of_measured_hbm_peak": n_shard * nbytes / k_s / 1e9 / peak_gbs,
                               "frac_of_nominal_8TBps": n_shard * nbytes / k_s / 1e9 / 8000.0,
                               "tests_per_s_per_gpu": B * n_shard / k_s, "roofline_ms": max(t_hbm, t_popc) * 1e3,
                               "frac_of_roofline": max(t_hbm, t_popc) / k_s, "bound": "hbm" if t_hbm > t_popc else "popc",
                               "kernel": "bmih_verify_kernel (scan mode)" if ix.get_param("scan.last_batched") else "scan_topk_kernel"})
        clocks = sampler.stop()
        best = max(sweep_rows, key=lambda r: r["queries_per_s"])
        b1 = sweep_rows[0]
        headline = dict(value=best["queries_per_s"], ms_step=best["ms_per_batch"], kernel_ns=[b1["scan_kernel_ms"] * 1e6], clocks=clocks,
                        launches=total_launches, parity_ok=None, ocheck=ocheck, breakdown=None, radius=-1)

    # ---- the round-1 batch size beside it (same index, same kernels): 4096 queries per batch ------------------------------
    batch_4096 = None
    if args.config == "headline" and mode == "mih" and Q > 4096:
        small = [db[:4096].contiguous() for db in dev_batches]
        for i in range(2):
            searcher.search(small[i], K_NN, mode=mode)
        barrier()
        reps = max(3, min(args.steps, 5))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = []
        s0.record()
        for i in range(reps):
            searcher.search(small[i % len(small)], K_NN, mode=mode)
            ks.append(ix.get_param("last_kernel_ns"))
        s1.record()
        barrier()
        ms = max_over_ranks(s0.elapsed_time(s1) / reps)
        batch_4096 = {"queries_per_s": 4096 / (ms * 1e-3), "ms_per_batch": ms, "verify_kernel_ms": float(np.mean(ks)) * 1e-6,
                      "what": "the same measurement with round 1's batch of 4096 queries (device-resident), %d batches" % reps}

    # ---- opt-in (VC_BENCH_WEAK=1), N > 1: the batch grown with the GPU count, i.e. the same work per GPU as the one-GPU run --------
    # (`value` is strong scaling at a fixed batch; the search is bucket-stationary, so what a shard of N / G codes loses is the
    # number of queries that share a bucket read - a batch of G x Q gives it back.  Not part of `value`.)
    batch_x_gpus = None
    if args.config == "headline" and mode == "mih" and world > 1 and os.environ.get("VC_BENCH_WEAK"):
        QW = Q * world
        big = torch.cat([dev_batches[i % len(dev_batches)] for i in range(world)])[:QW].contiguous()
        for i in range(2):
            searcher.search(big, K_NN, mode=mode)
        barrier()
        reps = 3
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks, ss = [], []
        s0.record()
        for i in range(reps):
            searcher.search(big, K_NN, mode=mode)
            ks.append(ix.get_param("last_kernel_ns"))
            ss.append(ix.get_param("mih.search_ns"))
        s1.record()
        barrier()
        ms = max_over_ranks(s0.elapsed_time(s1) / reps)
        batch_x_gpus = {"batch": QW, "queries_per_s": QW / (ms * 1e-3), "ms_per_batch": ms, "verify_kernel_ms": float(np.mean(ks)) * 1e-6,
                        "search_ms_this_rank": float(np.mean(ss)) * 1e-6,
                        "what": "batch = %d x %d GPUs (work per GPU as in the one-GPU run: weak scaling), device-resident, %d batches" % (Q, world, reps)}
        del big

    # ---- the reference's approximate mode (search_worker.cc:93-157: stop at 20 k distinct candidates) on the same index, one GPU ------
    approximate_mode = None
    if args.config == "headline" and mode == "mih" and world == 1:
        keys_a = torch.empty((Q, K_NN), dtype=torch.int64, device=dev)
        cur = torch.cuda.current_stream().cuda_stream
        for i in range(2):
            ix.search_mih_dev(dev_batches[i % len(dev_batches)].data_ptr(), Q, K_NN, keys_a.data_ptr(), approximate=True, stream=cur)
        torch.cuda.synchronize()
        reps = max(3, min(args.steps, 5))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(reps):
            ix.search_mih_dev(dev_batches[i % len(dev_batches)].data_ptr(), Q, K_NN, keys_a.data_ptr(), approximate=True, stream=cur)
        s1.record()
        torch.cuda.synchronize()
        ms = s0.elapsed_time(s1) / reps
        approximate_mode = {"queries_per_s": Q / (ms * 1e-3), "ms_per_batch": ms, "batch": Q, "batched": bool(ix.get_param("mih.last_batched")),
                            "queries_answered_by_the_per_query_kernel": int(ix.get_param("mih.last_redo")),
                            "what": "vc_search_mih_dev(approximate = 1), device-resident, %d batches; not part of `value`" % reps}
        del keys_a

    value, ms_step, kernel_ns, clocks = headline["value"], headline["ms_step"], headline["kernel_ns"], headline["clocks"]
    launches, radius = headline["launches"], headline["radius"]

    # ---- roofline of the dominant kernel: algorithmic bytes from the kernel's own statistics ------------
    stats_t = torch.zeros((Q, capi.STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    keys_t = torch.empty((Q, K_NN), dtype=torch.int64, device=dev)
    roofline = None
    integer_pipe = None
    exec_pairs = 0
    popc_per_pair_pf = max(1, nbytes // 8)           # one POPC per 64 bits with the lower-bound filter (64- and 128-bit codes)
    if mode == "mih":
        if world > 1:
            barrier()                                # the per-step exchanges inside the search are collective: all ranks together
        ix.search_mih_dev(dev_batches[-1].data_ptr(), Q, K_NN, keys_t.data_ptr(), max_radius=radius, d_stats=stats_t.data_ptr(),
                          stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        st = stats_t.cpu().numpy().view(capi.STATS_DTYPE).reshape(-1)
        probes, cands = float(st["probes"].sum()), float(st["candidates"].sum())
        batched = bool(ix.get_param("mih.last_batched"))
        k_s = float(np.mean(kernel_ns)) * 1e-9
        per_query_bytes = probes * 8 + cands * nbytes     # what one-query-at-a-time MIH must move
        if batched:
            # bucket-stationary: every probed bucket is read once per radius level for all its queries
            algo_bytes = float(ix.get_param("mih.last_bucket_codes")) * nbytes + probes * 8
            kernel = "bmih_verify_kernel (sum over radius levels)"
        else:
            algo_bytes = per_query_bytes
            kernel = "mih_search_kernel"
        achieved = algo_bytes / k_s / 1e9
        roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                    "frac": achieved / peak_gbs, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": k_s * 1e3,
                    "probes_per_query": probes / Q, "candidates_per_query": cands / Q,
                    "mean_radius": float(st["radius"].mean()), "queries_redone_by_the_per_query_kernel": ix.get_param("mih.last_redo"),
                    "per_query_formula_GBps": per_query_bytes / k_s / 1e9,
                    "survey_formula_GBps": (probes * 8 + cands * (4 + nbytes)) / k_s / 1e9}
        # POPCs per test: the lower-bound filter (one per 64 bits) for 64- and 128-bit codes, the exact distance (two per 64 bits) for 256-bit
        popc_per_pair = popc_per_pair_pf if nbytes <= 16 else nbytes // 4
        if batched:
            # per search step the lower bound is the slower of its HBM time and its POPC time (north_star's roofline);
            # the sum over the steps against the measured verify-kernel time is the fraction of that combined roofline
            n_steps = ix.get_param("mih.last_levels")
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            popc_peak = POPC_PER_CLK_PER_SM * ix.get_param("num_sms") * sm_hz          # POPC/s of the GPU
            steps, bound_s, meas_s = [], 0.0, 0.0
            exec_pairs = 0
            for i in range(n_steps):
                sc, sp = ix.get_param("mih.step_codes.%d" % i), ix.get_param("mih.step_pairs.%d" % i)
                sx = ix.get_param("mih.step_exec.%d" % i)      # tests really executed (queries leave buckets at their k-th id)
                sn = ix.get_param("mih.step_ns.%d" % i) * 1e-9
                t_hbm, t_popc = sc * nbytes / (peak_gbs * 1e9), sx * popc_per_pair / popc_peak
                exec_pairs += sx
                steps.append({"hbm_ms": t_hbm * 1e3, "popc_ms": t_popc * 1e3, "measured_ms": sn * 1e3,
                              "bound": "hbm" if t_hbm > t_popc else "popc", "tests_executed": sx, "bucket_members_x_queries": sp})
                bound_s += max(t_hbm, t_popc)
                meas_s += sn
            roofline["which_bound_binds"] = (
                "`frac` is achieved / peak on the HBM axis, as the bench contract defines it; with %d queries per batch a probed bucket is read "
                "once for %.0f queries on average, so %d of the %d search steps are bound by the integer (POPC) pipe, not by HBM - north_star's "
                "roofline (the slower of the code bytes at HBM speed and the popcount work at integer-pipe peak, per step) is `combined.frac`" % (
                    Q, sum(x["bucket_members_x_queries"] for x in steps) / max(1.0, float(sum(ix.get_param("mih.step_codes.%d" % i) for i in range(n_steps)))),
                    sum(1 for x in steps if x["bound"] == "popc"), n_steps))
            roofline["combined"] = {"what": "sum over search steps of max(HBM time, POPC time) / measured verify-kernel time; "
                                            "%d POPC per test, POPC peak %.1f / clk / SM (tools/microbench.cu)" % (popc_per_pair, POPC_PER_CLK_PER_SM),
                                    "bound_ms": bound_s * 1e3, "measured_ms": meas_s * 1e3, "frac": bound_s / meas_s, "steps": steps}
            import glob
            tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))      # the newest round's capture
            if tpaths:
                tj = json.load(open(tpaths[-1]))
                if tj["config"] == {"n_codes": n_total, "batch": Q, "k": K_NN, "n_gpus": world}:
                    roofline["traffic"] = tj["traffic_bytes_per_search"]
                    roofline["traffic_source"] = "profiles/%s (ncu --set full, all verify launches of one search)" % os.path.basename(tpaths[-1])
        sm_clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        n_sms = ix.get_param("num_sms")
        pairs_s = (exec_pairs if batched else cands) / k_s
        integer_pipe = {"what": "code-query distance tests executed by the dominant kernel; it is POPC-bound when batched",
                        "pairs_per_s": pairs_s, "pairs_per_clk_per_sm": pairs_s / n_sms / sm_clk,
                        "popc_peak_per_clk_per_sm": POPC_PER_CLK_PER_SM, "popc_per_pair": popc_per_pair,
                        "frac_of_popc_peak": pairs_s / n_sms / sm_clk / POPC_PER_CLK_PER_SM * popc_per_pair,
                        "peak_source": "tools/microbench.cu on this pool's B200s: %.1f POPC/clk/SM (profiles/microbench_r02.log; nominal 16)" % POPC_PER_CLK_PER_SM}
    else:
        k_s = float(np.mean(kernel_ns)) * 1e-9
        achieved = n_shard * nbytes / k_s / 1e9
        roofline = {"bound": "hbm", "kernel": "scan_topk_kernel (batch 1)", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                    "frac": achieved / peak_gbs, "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": n_shard * nbytes, "kernel_ms": k_s * 1e3}

    # ---- e2e: host buffers through the C ABI (N = 1) or pinned -> device -> search -> merged -> host (N > 1) ----
    e2e_steps = max(3, min(args.steps, 5))
    out_host = torch.empty((Q, K_NN), dtype=torch.int64).pin_memory()

    def e2e_step(i):
        hb = host_batches[(args.warmup + i) % n_batches]
        if world == 1 and mode == "mih":
            ix.search_mih(hb, K_NN, max_radius=radius, with_stats=False)
        elif world == 1:
            ix.search_linear(hb, K_NN)
        else:
            dq = pinned[(args.warmup + i) % n_batches].to(dev, non_blocking=True)
            res = searcher.search(dq, K_NN, mode=mode, max_radius=radius)
            out_host.copy_(res, non_blocking=True)
            torch.cuda.synchronize()

    # the first host-buffer call allocates the library's pinned staging buffers (cudaHostAlloc: milliseconds): untimed
    barrier()
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e = {"value": Q / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": Q * nbytes,
           "d2h_bytes_per_step": Q * K_NN * 8 + (Q * 4 if world == 1 else 0)}

    # ---- the brute-force scan (config C4) on the same shard: HBM-bound at small batches ---------------------
    scan = {}
    shard_bytes = n_shard * nbytes
    for B in ((1, 2, 4, 1024) if args.config == "headline" else ()):
        if B > 8 and os.environ.get("VC_BENCH_SKIP_BIG_SCAN"):
            continue
        qd = dev_batches[0][:B].contiguous()
        reps = 5 if B <= 8 else 1
        for _ in range(2 if B <= 8 else 1):
            searcher.search(qd, K_NN, mode="linear")
        barrier()
        ks = []
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            searcher.search(qd, K_NN, mode="linear")
            ks.append(ix.get_param("last_kernel_ns"))
        s1.record()
        barrier()
        ms = max_over_ranks(s0.elapsed_time(s1) / reps)
        k_ms = float(np.mean(ks)) * 1e-6
        scan["B=%d" % B] = {"queries_per_s": B / (ms * 1e-3), "ms_per_batch": ms, "scan_kernel_ms": k_ms,
                            "hbm_GBps_per_gpu": shard_bytes / (k_ms * 1e-3) / 1e9,
                            "frac_of_peak": shard_bytes / (k_ms * 1e-3) / 1e9 / peak_gbs,
                            "frac_of_nominal_8TBps": shard_bytes / (k_ms * 1e-3) / 1e9 / 8000.0,
                            "pair_rate_T_per_s": B * n_shard / (k_ms * 1e-3) / 1e12}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own linear scan on a bounded sample -----
    cpu_baseline = None
    if world == 1 and rank == 0 and not os.environ.get("VC_BENCH_SKIP_CPU"):
        try:
            from oracle import reference as ref, restatement as R
            cores = os.cpu_count() or 1
            sample_n = int(os.environ.get("VC_BENCH_REF_N", 64_000_000 * 8 // nbytes))
            nq_cpu = max(cores, 8) * 8
            codes = R.synth_codes(DB_SEED, 0, sample_n, nbytes)
            mem = ref.RefMem(codes, 0)
            qh = make_queries(QUERY_SEED, nq_cpu, nbytes)
            t0 = time.perf_counter()
            mem.linear_search(qh, K_NN, n_procs=cores)
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": nq_cpu / dt * sample_n / n_total, "unit": "queries/s", "cores": cores, "kind": "reference",
                            "sample": "reference src/linear_search.cc:39-64 (unmodified, oracle/_ref) over an in-memory %d-code "
                                      "sample, %d queries on %d forked processes, %.1f s; scaled by sample/N (scan is linear in N)"
                                      % (sample_n, nq_cpu, cores, dt)}
        except Exception as ex:   # the checker library did not travel / build
            cpu_baseline = {"value": None, "unit": "queries/s", "cores": os.cpu_count(), "kind": "reference",
                            "sample": "unavailable: %r" % (ex,)}

    if rank == 0:
        if mode == "mih":
            workload = "%s k-NN (MIH m=%d, %s) k=%d over %d x %d-bit uniform codes, batch %d, id-sharded over %d GPU(s)" % (
                "fixed-radius" if fixed_radius else "exact", N_TABLES,
                "radius sweep %s, `value` = the largest radius" % cfg["radii"] if fixed_radius else "strict stop rule", K_NN, n_total, CODE_BITS, Q, world)
        else:
            workload = "exact k-NN (linear scan) k=%d over %d x %d-bit codes, batch sweep up to %d, id-sharded over %d GPU(s)" % (K_NN, n_total, CODE_BITS, Q, world)
        line = {
            "metric": metric_name(cfg), "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64 xor+popcount", "data": "synthetic",
            "config": {"workload": workload, "name": args.config, "reduced": cfg["reduced"],
                       "mode": mode, "batch": Q, "n_codes": n_total, "codes_per_gpu": n_shard,
                       "l2": "inputs larger than L2 (index %.1f GB per GPU, >= 300 MB touched per query)" % (ix.info()["device_bytes"] / 1e9),
                       "parallelism": "id-shard x%d (ids interleaved), top-k all-gather + merge kernel, per-step exchange of distance / id histograms: %s" % (world, searcher.exchange)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "integer_pipe": integer_pipe, "cpu_baseline": cpu_baseline,
            "oracle_check": headline["ocheck"], "breakdown": headline["breakdown"], "batch_4096": batch_4096, "approximate_mode": approximate_mode,
            "parity_selfcheck": None if headline["parity_ok"] is None else "mih == linear scan on 8 queries: %s" % headline["parity_ok"],
            "scan": scan, "build": {"generate_s": t_gen, "build_tables_s": t_build, "index_GB": ix.info()["device_bytes"] / 1e9},
        }
        if args.config != "headline":
            line["sweep"] = sweep_rows
        if batch_x_gpus is not None:
            line["batch_x_gpus"] = batch_x_gpus
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    searcher.close()
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    _private_stdout()
    main()
