"""Summarises .ncu-rep captures (ncu --set full) into a small markdown file for profiles/.
    python tools/ncu_summary.py out.md rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.avg.per_second",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source_top(rep, launch, n=12):
    """Hottest SASS instructions of launch number `launch` of the report (stall samples per instruction address)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
    if not hi:
        return []
    hdr = rows[hi[0]]
    isrc, ismp, iadr = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Address")
    end = hi[1] - 1 if len(hi) > 1 else len(rows)          # the SASS view; a second (source-line) view may follow
    seen, data = set(), []
    for r in rows[hi[0] + 1:end]:
        if len(r) > ismp and r[ismp].isdigit() and r[iadr] not in seen:
            seen.add(r[iadr])
            data.append((int(r[ismp]), r[isrc].strip()))
    tot = sum(s for s, _ in data) or 1
    return [(s, 100.0 * s / tot, src) for s, src in sorted(data, reverse=True)[:n]]


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    md = ["# ncu summaries (`ncu --set full --clock-control none`, one launch each)\n"]
    for rep in reps:
        hdr, units, launches = raw(rep)
        for li, vals in enumerate(launches):
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            md.append("## %s\n" % rep.split("/")[-1])
            md.append("kernel: `%s`\n" % d.get("Kernel Name", "?"))
            md.append("| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in d:
                    md.append("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
            stalls = []
            for h in hdr:
                if "pcsamp_warps_issue_stalled_" in h and not h.endswith("_not_issued") and d[h]:
                    try:
                        stalls.append((float(d[h].replace(",", "")), h.split("issue_stalled_")[1]))
                    except ValueError:
                        pass
            tot_s = sum(v for v, _ in stalls) or 1.0
            stalls = sorted(stalls, reverse=True)[:7]
            md.append("\ntop warp stall reasons (% of the stall samples): " + ", ".join("%s %.1f" % (h, 100.0 * v / tot_s) for v, h in stalls))
            top = source_top(rep, li)
            if top:
                md.append("\nhottest SASS instructions (stall samples):\n")
                md.append("| samples | % | instruction |\n|---|---|---|")
                for s, pct, src in top:
                    md.append("| %d | %.1f | `%s` |" % (s, pct, src.replace("|", "\\|")))
            md.append("")
    open(out_path, "w").write("\n".join(md) + "\n")
    print("wrote", out_path)


if __name__ == "__main__":
    main()
