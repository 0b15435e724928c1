"""Per-kernel summary table of an ncu --set full report (means over the captured launches).
usage: python scripts/ncu_summary.py REPORT.ncu-rep OUT.md [TITLE]"""
import collections
import csv
import re
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size"]


def val(d, c):
    if c not in ix:
        return float("nan")
    try:
        f = float(d[ix[c]].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[ix[c]]
    if c.startswith("dram__bytes"):  # -> MB
        f *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
    if c == "gpu__time_duration.sum":  # -> ms
        f *= {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
    return f


agg = collections.OrderedDict()
for d in data:
    name = re.sub(r"^.*unnamed>::", "", re.sub(r"\(.*", "", d[ix["Kernel Name"]])).replace("void ", "")
    agg.setdefault(name, []).append([val(d, c) for c in COLS])
out = [f"# {title}", "",
       "ncu --set full --clock-control none; per kernel the MEAN over the captured launches.  Times under ncu are "
       "cold-cache and serialised: shares matter, not absolutes.", "",
       "| kernel | n | time ms | dram read MB | dram write MB | dram GB/s | issue active % | warps active % | "
       "thr/inst | regs | grid | block |", "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, ls in agg.items():
    n = len(ls)
    m = [sum(x[i] for x in ls) / n for i in range(len(COLS))]
    out.append(f"| {name} | {n} | {m[0]:.4f} | {m[1]:.1f} | {m[2]:.1f} | {(m[1] + m[2]) / m[0]:.0f} | {m[4]:.1f} | {m[5]:.1f} | "
               f"{m[6]:.1f} | {m[7]:.0f} | {m[8]:.0f} | {m[9]:.0f} |")
open(out_path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
