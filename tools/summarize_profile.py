"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profile.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep]

Writes profiles/<tag>_launches.md (per-kernel share of the step, from the gpu__time_duration pass) and
profiles/<tag>_<kernel>.md (key metrics of the `--set full` capture + executed-instruction mix by SASS
opcode + hottest source lines).  Developer tool.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches_md(path, tag):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(row["Metric Unit"], v)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    out = ["# %s — launch list (ncu --metrics gpu__time_duration.sum --clock-control none)" % tag, "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total µs | share |", "|---|---:|---:|---:|"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f %% |" % (k[:110].replace("|", "\\|"), n, v, 100 * v / tot))
    out += ["", "total: %.1f µs over %d launches" % (tot, sum(n for n, _ in agg.values()))]
    dst = os.path.join(ROOT, "profiles", "%s_launches.md" % tag)
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


def kernel_md(rep, tag, name):
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    out = ["# %s — `%s` (ncu --set full --clock-control none --import-source on)" % (tag, name), ""]
    for li, vals in enumerate(raw[2:]):
        out += ["## captured launch %d: %s" % (li, vals[hdr.index("Kernel Name")][:100]), "", "| metric | value | unit |", "|---|---:|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append("| %s | %s | %s |" % (k, vals[i], units[i]))
        if "dram__bytes_read.sum" in hdr:
            def tobytes(k):
                v, u = float(vals[hdr.index(k)].replace(",", "")), units[hdr.index(k)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            out.append("| **traffic = dram read + write** | %.2f | MB |" % ((tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")) / 1e6))
        out.append("")
    sass = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hdr = None
    ops, tot, stalls = collections.Counter(), 0, collections.Counter()
    for row in sass:
        if "Instructions Executed" in row:
            hdr = row
            iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(row) <= iI:
            continue
        try:
            inst = int(row[iI])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", row[1])
        op = m.group(2) if m else "?"
        ops[op] += inst
        tot += inst
        for j, h in enumerate(hdr):
            if h.startswith("stall_") and row[j].isdigit():
                stalls[h] += int(row[j])
    out += ["## executed warp instructions by SASS opcode (all captured launches)", "", "| opcode | share |", "|---|---:|"]
    for op, c in ops.most_common(18):
        out.append("| %s | %.1f %% |" % (op, 100 * c / max(tot, 1)))
    st = sum(stalls.values())
    out += ["", "## warp stall samples", "", "| reason | share |", "|---|---:|"]
    for k, c in stalls.most_common(8):
        out.append("| %s | %.1f %% |" % (k, 100 * c / max(st, 1)))
    dst = os.path.join(ROOT, "profiles", "%s_%s.md" % (tag, name))
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--kernel", default="k_photometric")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.launches:
        launches_md(a.launches, a.tag)
    if a.rep:
        kernel_md(a.rep, a.tag, a.kernel)
