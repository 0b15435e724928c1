"""Wall-clock breakdown of the host-facing call sequence (diagnostic)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbindex_b200 as dbi
from dbindex_b200 import synth
from bench import CFG2

res, off = synth.config_proteome(2)
p = dbi.default_params(**CFG2)
lo = np.linspace(700, 5000, 10000); hi = lo + 0.01
for it in range(4):
    t = [time.perf_counter()]
    g = dbi.GpuIndex(p); t.append(time.perf_counter())
    g.add_proteins(res, off); t.append(time.perf_counter())
    g.upload(); t.append(time.perf_counter())
    g.build(); t.append(time.perf_counter())
    b, c = g.query(lo, hi); t.append(time.perf_counter())
    st = g.stats()
    g.close(); t.append(time.perf_counter())
    names = ["create", "add_proteins", "upload", "build", "query", "close"]
    print(it, {n: round(1e3 * (t[i + 1] - t[i]), 2) for i, n in enumerate(names)}, "entries", st["n_entries"])
p.profile = 1
g = dbi.GpuIndex(p); g.add_proteins(res, off); g.upload()
for it in range(3):
    g.reset_index(); t0 = time.perf_counter(); g.build(); t1 = time.perf_counter()
    st = g.stats()
    print("resident build wall ms", round(1e3 * (t1 - t0), 2), {k: round(v, 3) for k, v in st["stage_ms"].items() if v > 0},
          "dom", round(st["dom_ms"], 3), st["dom_launches"])
