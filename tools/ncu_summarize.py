#!/usr/bin/env python3
"""Turn ncu captures brought back in gpurun_out/ into the committed summaries under profiles/.

  ncu_summarize.py launches <launches.csv> <out_summary.txt>             per-kernel totals and shares of a launch list
  ncu_summarize.py full <prof.ncu-rep> <kernel regex> <out prefix>        raw-page CSV + a summary of the metrics that matter
                                                                          (+ <out prefix>.json with dram bytes, pipes, IPC and the
                                                                          kernel-source hash, for bench.py's roofline.traffic)"""
import csv
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_hash import kernel_sources_hash  # noqa: E402

KEEP = re.compile(r"dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|launch__(grid_size|block_size|registers_per_thread|occupancy_limit|shared_mem_per_block_static)|"
                  r"sm__cycles_elapsed\.avg$|sm__warps_active\.avg\.per_cycle_active|sm__inst_executed_pipe_(alu|fma|fp64|xu|lsu|fmaheavy)\S*pct_of_peak_sustained_active|"
                  r"sm__pipe_(alu|fma|fmaheavy|fp64)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|smsp__inst_executed\.(sum|avg)$|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active|sm__icc_request_hit_rate|smsp__issue_active\.avg\.per_cycle_active|sm__inst_executed\.avg\.per_cycle_elapsed|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared|smsp__inst_executed_op_shared|lts__t_sector_hit_rate")


def launches(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
            rows.append((r["Kernel Name"], r.get("Grid Size", ""), ns))
    pmt = [(k, g, ns) for k, g, ns in rows if "pmt::" in k or k.startswith("k_") or "k_level" in k or "k_tree" in k or "k_leaves" in k]
    total = sum(ns for _, _, ns in pmt) or 1.0
    agg = {}
    for k, g, ns in pmt:
        name = re.sub(r"\(.*", "", k).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns
    with open(out, "w") as f:
        f.write("# summary of %s (ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised launches; compare SHARES)\n" % os.path.basename(path))
        f.write("# libpmt kernels only; %d launches in the list, %d of them libpmt's\n" % (len(rows), len(pmt)))
        for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s %5d launches %10.3f ms %5.1f %%\n" % (name, cnt, ns / 1e6, 100 * ns / total))
        by_grid = {}
        for k, g, ns in pmt:
            if "coop" in k:
                by_grid.setdefault((re.sub(r"\(.*", "", k).replace("void ", ""), g), []).append(ns)
        f.write("# per-launch durations of the cooperative kernels by grid size (us, median):\n")
        for (name, g), v in sorted(by_grid.items()):
            v.sort()
            f.write("%-60s grid %-16s %4d launches  median %8.1f us\n" % (name, g, len(v), v[len(v) // 2] / 1e3))
    print(open(out).read())


def full(rep, kernel_re, prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    with open(prefix + "_ncu_raw.csv", "w") as f:
        f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    header, units = rows[0], rows[1]
    name_col = header.index("Kernel Name")
    picked = [r for r in rows[2:] if re.search(kernel_re, r[name_col])]
    out = ["# %s: kernels matching /%s/ (%d launches captured); ncu --set full --clock-control none --import-source on" % (os.path.basename(rep), kernel_re, len(picked))]
    summary = []
    for r in picked:
        d = {h: (v, u) for h, v, u in zip(header, r, units)}
        out.append("  Kernel Name  %s   grid %s block %s" % (r[name_col], d.get("Grid Size", ("", ""))[0], d.get("Block Size", ("", ""))[0]))
        vals = {}
        for h in header:
            if KEEP.search(h):
                out.append("  %-95s %s %s" % (h, d[h][0], d[h][1]))
                try:
                    vals[h] = float(d[h][0].replace(",", ""))
                except ValueError:
                    pass
        summary.append((r[name_col], d, vals))
        out.append("")
    with open(prefix + "_summary.txt", "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[:80]))
    return summary


def main():
    if sys.argv[1] == "launches":
        return launches(sys.argv[2], sys.argv[3])
    if sys.argv[1] == "full":
        summary = full(sys.argv[2], sys.argv[3], sys.argv[4])
        if summary:
            name, d, v = summary[0]

            def unit_scale(key):
                u = d[key][1].lower()
                return {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            rd = v.get("dram__bytes_read.sum", 0) * unit_scale("dram__bytes_read.sum") if "dram__bytes_read.sum" in d else None
            wr = v.get("dram__bytes_write.sum", 0) * unit_scale("dram__bytes_write.sum") if "dram__bytes_write.sum" in d else None
            js = {"kernel_sources_sha16": kernel_sources_hash(), "kernel": name, "source": os.path.basename(sys.argv[2]),
                  "dram_bytes_read": rd, "dram_bytes_write": wr, "grid": d.get("Grid Size", ("", ""))[0], "block": d.get("Block Size", ("", ""))[0],
                  "metrics": {k: val for k, val in v.items()}}
            with open(sys.argv[4] + ".json", "w") as f:
                json.dump(js, f, indent=1)
        return
    raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
