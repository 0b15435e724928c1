"""Small profiling target: two builds of the bench workload (cfg2, full size) + one query batch.
The first build is the warm-up; run under ncu as scripts/gpu_round.sh / scripts/gpu_prof.sh / scripts/gpu_list.sh do."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbindex_b200 as dbi  # noqa: E402
from dbindex_b200 import synth  # noqa: E402
from bench import CFG2  # noqa: E402

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n_builds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
res, off = synth.config_proteome(2, n_prot)
p = dbi.default_params(**CFG2)
g = dbi.GpuIndex(p)
g.add_proteins(res, off)
g.upload()
for _ in range(n_builds):
    g.reset_index()
    g.build()
st = g.stats()
m = g.fetch(st["n_entries"] // 2, 4096, with_ids=False)["mass"]
_, _, lo, hi = synth.synth_queries(m, 10000, 1)
h = g.query_hits(lo, hi)  # bounds + materialisation of every hit (what the bench step does after the build)
print("entries", st["n_entries"], "hits", int(h["hit_off"][-1]), "launches", g.kernel_launches())
