"""Hot CUDA source lines of one kernel from an ncu --set full --import-source on report.
usage: python scripts/ncu_hot_lines.py REPORT.ncu-rep KERNEL_REGEX [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, out = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
    elif hdr and r[0].isdigit():
        try:
            out.append((float(r[hdr["Instructions Executed"]]), cur_file, int(r[0]), r[1], r[hdr["Avg. Threads Executed"]],
                        r[hdr["# Samples"]]))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


tots = sum(num(o[5]) for o in out) or 1
print(f"total warp instructions {tot:.0f}")
for v, f, ln, src, thr, smp in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{100 * v / tot:5.1f}% inst {100 * num(smp) / tots:5.1f}% smp thr={thr:>4s} {f}:{ln:<4d} {src.strip()[:100]}")
