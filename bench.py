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
scan   : the brute-force path (config C4): passes/s at small batches as HBM GB/s vs peak, queries/s at a large batch.
cpu_baseline: the reference's own linear scan (oracle/_ref, unmodified sources) on a bounded sample, host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
Environment overrides for quick runs: VC_BENCH_N (total codes), VC_BENCH_Q (batch), VC_BENCH_MODE (mih|linear).
"""
import argparse
import json
import os
import subprocess
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

# headline workload (BASELINE.json); the environment overrides exist to measure the other configs with the same harness
K_NN = int(os.environ.get("VC_BENCH_K", 100))
CODE_BITS = int(os.environ.get("VC_BENCH_BITS", 64))
N_TABLES = int(os.environ.get("VC_BENCH_TABLES", 4))
DB_SEED, QUERY_SEED = 12345, 67890


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


def reference_arm(args, n_total, world, rank):
    """--impl reference: the reference's own CPU linear scan (oracle/_ref) on a bounded sample."""
    if rank != 0:
        return
    from oracle import reference as ref, restatement as R
    cores = os.cpu_count() or 1
    sample_n = int(os.environ.get("VC_BENCH_REF_N", 64_000_000))
    nq = max(cores, 8) * 4
    codes = R.synth_codes(DB_SEED, 0, sample_n, CODE_BITS // 8)
    mem = ref.RefMem(codes, 0)
    times = []
    for step in range(args.warmup + args.steps):
        q = make_queries(QUERY_SEED + step, nq, CODE_BITS // 8)
        t0 = time.perf_counter()
        mem.linear_search(q, K_NN, n_procs=cores)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    qps_sample = nq / per_step
    qps = qps_sample * sample_n / n_total          # the scan is linear in N
    line = {
        "impl": "reference", "metric": "queries/sec (k=100, 1B 64-bit codes)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 xor+popcount", "data": "synthetic",
        "config": {"workload": "exact k-NN, k=100, %d x 64-bit uniform codes" % n_total},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "reference",
                         "sample": "reference src/linear_search.cc:39-64 (unmodified, oracle/_ref) over an in-memory "
                                   "%d-code sample, %d queries/step on %d forked processes; scaled by sample/N (scan is linear in N)"
                                   % (sample_n, nq, cores)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_total = int(os.environ.get("VC_BENCH_N", 1_000_000_000))
    Q = int(os.environ.get("VC_BENCH_Q", 4096))
    mode = os.environ.get("VC_BENCH_MODE", "mih")
    nbytes = CODE_BITS // 8

    if args.impl == "reference":
        reference_arm(args, n_total, world, rank)
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
    e = b + n_shard
    t0 = time.perf_counter()
    ix = capi.Index(CODE_BITS, N_TABLES, device=local_rank, first_id=b)
    ix.set_param("id_stride", id_stride)
    ix.add_synthetic(n_shard, DB_SEED)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    ix.build()
    t_build = time.perf_counter() - t0
    ix.set_param("profile", 1)
    searcher = ShardedSearcher(ix)

    n_batches = args.warmup + args.steps
    host_batches = [make_queries(QUERY_SEED + i, Q, nbytes) for i in range(n_batches)]
    pinned = [torch.from_numpy(hb).pin_memory() for hb in host_batches]
    dev_batches = [p.to(dev) for p in pinned]
    torch.cuda.synchronize()

    def step(i):
        return searcher.search(dev_batches[i], K_NN, mode=mode)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    # self-check outside the timed region: the MIH answer of a few queries equals the brute-force scan of the same
    # (sharded) database - bit-exact, through the same all-gather + merge
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
    kernel_ns = []
    barrier()
    ev0.record()
    for i in range(args.warmup, n_batches):
        step(i)
        kernel_ns.append(ix.get_param("last_kernel_ns"))
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = ix.get_param("launches") - launches0 + (0 if world == 1 else 0)
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    value = Q / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel: algorithmic bytes from the kernel's own statistics ------------
    stats_t = torch.zeros((Q, capi.STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    keys_t = torch.empty((Q, K_NN), dtype=torch.int64, device=dev)
    roofline = None
    integer_pipe = None
    exec_pairs = 0
    if mode == "mih":
        ix.search_mih_dev(dev_batches[-1].data_ptr(), Q, K_NN, keys_t.data_ptr(), d_stats=stats_t.data_ptr(),
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
                    "mean_radius": float(st["radius"].mean()),
                    "per_query_formula_GBps": per_query_bytes / k_s / 1e9,
                    "survey_formula_GBps": (probes * 8 + cands * (4 + nbytes)) / k_s / 1e9}
        if batched:
            # per search step the lower bound is the slower of its HBM time and its POPC time (north_star's roofline);
            # the sum over the steps against the measured verify-kernel time is the fraction of that combined roofline
            n_steps = ix.get_param("mih.last_levels")
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            popc_peak = 15.4 * ix.get_param("num_sms") * sm_hz          # tests/s at one POPC per test (prefiltered 64-bit codes)
            steps, bound_s, meas_s = [], 0.0, 0.0
            exec_pairs = 0
            for i in range(n_steps):
                sc, sp = ix.get_param("mih.step_codes.%d" % i), ix.get_param("mih.step_pairs.%d" % i)
                sx = ix.get_param("mih.step_exec.%d" % i)      # tests really executed (queries leave buckets at their k-th id)
                sn = ix.get_param("mih.step_ns.%d" % i) * 1e-9
                t_hbm, t_popc = sc * nbytes / (peak_gbs * 1e9), sx / popc_peak
                exec_pairs += sx
                steps.append({"hbm_ms": t_hbm * 1e3, "popc_ms": t_popc * 1e3, "measured_ms": sn * 1e3,
                              "bound": "hbm" if t_hbm > t_popc else "popc", "tests_executed": sx, "bucket_members_x_queries": sp})
                bound_s += max(t_hbm, t_popc)
                meas_s += sn
            roofline["combined"] = {"what": "sum over search steps of max(HBM time, POPC time) / measured verify-kernel time",
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
                        "popc_peak_per_clk_per_sm": 15.4, "popc_per_pair": 1 if nbytes <= 16 else nbytes // 4,
                        "frac_of_popc_peak": pairs_s / n_sms / sm_clk / 15.4 * (1 if nbytes <= 16 else nbytes // 4),
                        "peak_source": "tools/microbench.cu on this B200: 15.4 POPC/clk/SM"}

    # ---- e2e: host buffers through the C ABI (N = 1) or pinned -> device -> search -> merged -> host (N > 1) ----
    e2e_steps = max(3, min(args.steps, 5))
    out_host = torch.empty((Q, K_NN), dtype=torch.int64).pin_memory()

    def e2e_step(i):
        hb = host_batches[(args.warmup + i) % n_batches]
        if world == 1 and mode == "mih":
            ix.search_mih(hb, K_NN, with_stats=False)
        elif world == 1:
            ix.search_linear(hb, K_NN)
        else:
            dq = pinned[(args.warmup + i) % n_batches].to(dev, non_blocking=True)
            res = searcher.search(dq, K_NN, mode=mode)
            out_host.copy_(res, non_blocking=True)
            torch.cuda.synchronize()

    # the first host-buffer call allocates the library's pinned staging buffers (cudaHostAlloc: milliseconds): untimed
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = {"value": Q / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": Q * nbytes,
           "d2h_bytes_per_step": Q * K_NN * 8 + (Q * 4 if world == 1 else 0)}

    # ---- the brute-force scan (config C4) on the same shard: HBM-bound at small batches ---------------------
    scan = {}
    shard_bytes = (e - b) * nbytes
    for B in (1, 2, 4, 1024):
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
        ms = s0.elapsed_time(s1) / reps
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        k_ms = float(np.mean(ks)) * 1e-6
        scan["B=%d" % B] = {"queries_per_s": B / (ms * 1e-3), "ms_per_batch": ms, "scan_kernel_ms": k_ms,
                            "hbm_GBps_per_gpu": shard_bytes / (k_ms * 1e-3) / 1e9,
                            "frac_of_peak": shard_bytes / (k_ms * 1e-3) / 1e9 / peak_gbs,
                            "pair_rate_T_per_s": B * (e - b) / (k_ms * 1e-3) / 1e12}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's own linear scan on a bounded sample -----
    cpu_baseline = None
    if world == 1 and rank == 0 and not os.environ.get("VC_BENCH_SKIP_CPU"):
        try:
            from oracle import reference as ref, restatement as R
            cores = os.cpu_count() or 1
            sample_n = int(os.environ.get("VC_BENCH_REF_N", 64_000_000))
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
        line = {
            "metric": "queries/sec (k=%d, %s %d-bit codes)" % (K_NN, "1B" if n_total == 10**9 else str(n_total), CODE_BITS), "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64 xor+popcount", "data": "synthetic",
            "config": {"workload": "exact k-NN (MIH m=%d, strict stop rule) k=%d over %d x %d-bit uniform codes, batch %d, "
                                   "id-sharded over %d GPU(s)" % (N_TABLES, K_NN, n_total, CODE_BITS, Q, world)
                       if mode == "mih" else "exact k-NN (linear scan) k=%d over %d x %d-bit codes, batch %d" % (K_NN, n_total, CODE_BITS, Q),
                       "mode": mode, "batch": Q, "n_codes": n_total, "codes_per_gpu": e - b,
                       "l2": "inputs larger than L2 (tables %.1f GB per GPU, >= 300 MB touched per query)" % (ix.info()["device_bytes"] / 1e9),
                       "parallelism": "id-shard x%d (ids interleaved), NCCL all-gather top-k merge, per-step all-reduce of distance / id histograms" % world},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "integer_pipe": integer_pipe, "cpu_baseline": cpu_baseline,
            "parity_selfcheck": "mih == linear scan on 8 queries: %s" % parity_ok, "scan": scan, "build": {"generate_s": t_gen, "build_tables_s": t_build, "index_GB": ix.info()["device_bytes"] / 1e9},
        }
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    _private_stdout()
    main()
