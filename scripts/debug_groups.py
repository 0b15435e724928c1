import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbindex_b200 as dbi
from dbindex_b200 import synth
from oracle.oracle_py import Oracle
from tests.util import PARAM_SETS, bits, pack

p = dbi.default_params(**PARAM_SETS["cfg2_mods"])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
res, off = synth.synth_proteome(n, 1008, median_len=300, min_len=5)
for rep in range(2):
    g = dbi.GpuIndex(p); g.add_proteins(res, off); g.build()
    o = Oracle(p, threads=4); o.add_proteins(res, off); o.build()
    st = g.stats(); oc = o.counts()
    print("counts", st["n_emitted"], st["n_unique"], st["n_entries"], oc)
    got = g.fetch(0, st["n_entries"]); exp = o.entries()
    gm, em = bits(got["mass"]), bits(exp["mass"])
    m = min(len(gm), len(em))
    bad = np.nonzero(gm[:m] != em[:m])[0]
    print("mismatches", len(bad), "sorted?", bool(np.all(np.diff(got["mass"]) >= 0)))
    if len(bad):
        i = bad[0]
        print("first", i, got["mass"][i-2:i+4], exp["mass"][i-2:i+4])
        print("got pat", [hex(x) for x in got["modpat"][i-2:i+4]], "base", got["first_prot"][i-2:i+4], got["first_off"][i-2:i+4], got["len"][i-2:i+4])
        print("exp pat", [hex(x) for x in exp["modpat"][i-2:i+4]], "base", exp["first_prot"][i-2:i+4], exp["first_off"][i-2:i+4], exp["len"][i-2:i+4])
        # multiset comparison
        print("multiset equal:", np.array_equal(np.sort(gm), np.sort(em)))
        zeros = np.count_nonzero(gm == 0)
        print("zero masses in got:", zeros)
    g.close()
