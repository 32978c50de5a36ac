"""Collects the bench.py lines of this round's GPU runs (gpurun_out/r02_*.json - scratch) into the tracked summaries:
profiles/configs_r02.json (BASELINE configs C2 - C5 at the sizes / GPU counts they name, each with its oracle check),
profiles/scale_r02.json (headline at 1 / 2 / 8 GPUs, peer exchange vs NCCL) and profiles/bench_r02.json (the N = 1 line).
    python tools/collect_r02.py"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def load(name):
    path = os.path.join(OUT, name)
    if not os.path.exists(path) or os.path.getsize(path) == 0:
        return None
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


def brief(d, keep_sweep=True):
    if d is None:
        return None
    out = {k: d.get(k) for k in ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "config", "e2e", "gpu_launches", "clocks", "oracle_check",
                                 "breakdown", "batch_4096", "parity_selfcheck", "build")}
    r = d.get("roofline") or {}
    out["roofline"] = {k: r.get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "kernel_ms", "probes_per_query", "candidates_per_query",
                                             "mean_radius", "queries_redone_by_the_per_query_kernel")}
    if r.get("combined"):
        out["roofline"]["combined"] = r["combined"]
    out["integer_pipe"] = d.get("integer_pipe")
    if keep_sweep and d.get("sweep"):
        out["sweep"] = d["sweep"]
    if d.get("scan"):
        out["scan"] = d["scan"]
    return out


def main():
    configs = []
    for tag, name in (("C2: MIH 64-bit m=4, 100 M codes, k=100, one GPU", "r02_bench_c2.json"),
                      ("C3: MIH 128-bit m=8, 1 B codes over 8 GPUs, peer exchange + merge", "r02_bench8_c3.json"),
                      ("C3 on 2 GPUs (500 M codes per GPU)", "r02_bench2_c3.json"),
                      ("C4: brute-force scan 1 B x 64-bit, batch sweep, one GPU", "r02_bench_c4.json"),
                      ("C4 sharded over 8 GPUs", "r02_bench8_c4.json"),
                      ("C5: MIH 256-bit m=16, 500 M codes over 8 GPUs, fixed radius 0..3, k=1000", "r02_bench8_c5.json"),
                      ("C5 sparse tables (m=8, s=32) over 8 GPUs", "r02_bench8_c5m8.json"),
                      ("C5 reduced (125 M codes) on 2 GPUs", "r02_bench2_c5.json")):
        d = load(name)
        if d is not None:
            configs.append({"what": tag, "source": "gpurun_out/" + name, "line": brief(d)})
    extra = os.path.join(OUT, "r02_configs_misc.json")
    if os.path.exists(extra):
        try:
            configs.append({"what": "tools/bench_configs.py: C1 on the GPU, the reference's native shape, sparse C5", "lines": json.load(open(extra))})
        except Exception:
            pass
    json.dump(configs, open(os.path.join(ROOT, "profiles", "configs_r02.json"), "w"), indent=1)

    rows = []
    for n, peer, nccl in ((1, "r02_bench_final.json", None), (2, "r02_bench2_peer.json", "r02_bench2_nccl.json"),
                          (8, "r02_bench8_peer.json", "r02_bench8_nccl.json"), (8, "r02_bench8_hybrid.json", "r02_bench8_nccl2.json"),
                          (8, "r02_bench8_q4096.json", None)):
        for kind, name in ((("peer-memory exchange for everything (2 K bootstrap sample per shard)" if "peer" in (peer or "") else
                             "results over peer memory, histogram sums over NCCL (16 K bootstrap sample per shard)") if n > 1 else "single GPU", peer),
                           ("NCCL only (VC_XCHG=0)" + (" (2 K bootstrap sample per shard)" if nccl and "nccl2" not in nccl else ""), nccl)):
            d = load(name) if name else None
            if d is None:
                continue
            rows.append({"n_gpus": n, "exchange": kind, "batch": d["config"]["batch"], "queries_per_s": d["value"], "ms_per_batch": d["ms_per_step"],
                         "e2e_queries_per_s": d["e2e"]["value"], "oracle_check": d.get("oracle_check"), "breakdown": d.get("breakdown"),
                         "batch_4096": d.get("batch_4096"), "roofline_combined_frac": ((d.get("roofline") or {}).get("combined") or {}).get("frac"),
                         "gpu_launches": d.get("gpu_launches"), "clocks": d.get("clocks"), "source": "gpurun_out/" + name})
    json.dump({"what": "bench.py --gpus N, 1 B x 64-bit codes, k = 100, ids interleaved over the shards; strong scaling", "rows": rows},
              open(os.path.join(ROOT, "profiles", "scale_r02.json"), "w"), indent=1)
    d = load("r02_bench_final.json")
    if d is not None:
        json.dump(d, open(os.path.join(ROOT, "profiles", "bench_r02.json"), "w"), indent=1)
    d = load("r02_bench_ref.json")
    if d is not None:
        json.dump(d, open(os.path.join(ROOT, "profiles", "bench_ref_r02.json"), "w"), indent=1)
    print("configs:", len(configs), "scale rows:", len(rows))


if __name__ == "__main__":
    main()
