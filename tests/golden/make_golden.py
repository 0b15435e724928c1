"""Regenerates tests/golden/small_index.json from the oracle.

The reference cannot run here (Java, no JDK, un-vendored dependencies) and ships no fixtures, so
this golden file pins the ORACLE's output on a fixed input; it is a drift guard, not a reference
output.  Usage: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import dbindex_b200 as dbi  # noqa: E402
from dbindex_b200 import synth  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402
from tests.util import PARAM_SETS, bits, pack  # noqa: E402

res, off = synth.synth_proteome(6, 4242, median_len=90, min_len=20)
proteins = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
proteins += ["MKWVTFISLLLLFSSAYSRGVFRR", proteins[1], "AAGGLLKGGAALLK"]
name = "cfg2_mods"
o = Oracle(dbi.default_params(**PARAM_SETS[name]))
o.add_proteins(*pack(proteins))
assert o.build() == 0
e = o.entries()
out = {
    "param_set": name,
    "proteins": proteins,
    "entries": [[int(x) for x in bits(e["mass"])], e["first_prot"].tolist(), e["first_off"].tolist(),
                e["len"].tolist(), e["modpat"].tolist(), e["prot_list_off"].tolist(), e["prot_ids"].tolist()],
}
with open(os.path.join(os.path.dirname(__file__), "small_index.json"), "w") as f:
    json.dump(out, f)
print(len(e["mass"]), "entries")
