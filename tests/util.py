"""Shared helpers of the parity tests."""
from __future__ import annotations

import numpy as np

PARAM_SETS = {
    # name: kwargs of dbindex_b200.default_params
    "cfg1_tryptic": dict(),
    "cfg2_mods": dict(static_mods={"C": 57.02146}, diff_mods=[("M", 15.9949), ("STY", 79.96633)],
                      max_mods_per_peptide=3),
    "cfg3_semi": dict(semi=1),
    "mc0": dict(max_missed=0),
    "mc3_nocut_P": dict(max_missed=3, nocut="P"),
    "semi_nocut_mods": dict(semi=1, nocut="P", max_missed=1, diff_mods=[("M", 15.9949)], max_mods_per_peptide=2),
    "avg_no_h2o": dict(add_h2o_proton=0, min_mass=500.0, max_mass=4000.0),
    "lysc_neg_mod": dict(enzyme="K", diff_mods=[("ST", -18.010565), ("K", 42.010565)], max_mods_per_peptide=2),
    # 4 distinct shifts, K = 3 -> 85 class sequences > 32: the per-variant (non-group) path
    "many_classes": dict(diff_mods=[("M", 15.9949), ("ST", 79.96633), ("K", 42.010565), ("N", 0.984016)],
                         max_mods_per_peptide=3, max_missed=1),
    # 3 and 5 distinct shifts with K = 2: 13 and 31 class sequences (group path, generic class loops)
    "three_classes_k2": dict(diff_mods=[("M", 15.9949), ("ST", 79.96633), ("K", 42.010565)], max_mods_per_peptide=2),
    "five_classes_k2": dict(diff_mods=[("M", 15.9949), ("ST", 79.96633), ("K", 42.010565), ("N", 0.984016),
                                       ("Q", 0.98)], max_mods_per_peptide=2, max_missed=1),
    # the cross-linker parameter set (DBIndexImpl.java:443-491) and the occurrence filter
    "mandatory_K_no_h2o": dict(add_h2o_proton=0, mandatory_internal="K", max_missed=3, min_mass=500.0),
    "mandatory_KC_semi": dict(semi=1, mandatory_internal="KC", max_missed=1),
    "filter_K2_mods": dict(peptide_filter=("K", 2), max_missed=3, diff_mods=[("M", 15.9949)], max_mods_per_peptide=2),
    "mandatory_filter": dict(mandatory_internal="K", peptide_filter=("L", 3), max_missed=4),
    "wide_mass_mod4": dict(min_mass=0.0, max_mass=8000.0, max_missed=1, diff_mods=[("W", 15.9949)],
                           max_mods_per_peptide=4),
}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def canonical_entries(e: dict):
    """Entries as a canonically ordered structured array + per-entry protein lists.

    Index order is by exact mass; ties have no contractual order, so ties are
    ordered by (first_prot, first_off, len, modpat) -- for one unique peptide the first occurrence
    is deterministic, so equal indexes give equal tuples."""
    n = len(e["mass"])
    order = np.lexsort((e["modpat"], e["len"], e["first_off"], e["first_prot"], bits(e["mass"])))
    rec = np.zeros(n, dtype=[("mass", "u8"), ("prot", "u4"), ("off", "u4"), ("len", "u2"), ("pat", "u4")])
    rec["mass"] = bits(e["mass"])[order]
    rec["prot"] = e["first_prot"][order]
    rec["off"] = e["first_off"][order]
    rec["len"] = e["len"][order]
    rec["pat"] = e["modpat"][order]
    plo = e["prot_list_off"].astype(np.int64)
    sizes = (plo[1:] - plo[:-1])[order]
    # protein lists concatenated in canonical order
    starts = plo[:-1][order]
    total = int(sizes.sum())
    idx = np.repeat(starts - np.concatenate(([0], np.cumsum(sizes)[:-1])), sizes) + np.arange(total)
    ids = e["prot_ids"][idx] if total else np.zeros(0, np.uint32)
    return rec, sizes, ids


def assert_entries_equal(got: dict, exp: dict):
    assert len(got["mass"]) == len(exp["mass"]), (len(got["mass"]), len(exp["mass"]))
    # index order: masses must be bit-identical position by position (both sorted by exact mass)
    assert np.array_equal(bits(got["mass"]), bits(exp["mass"])), "mass arrays differ"
    assert np.all(np.diff(got["mass"]) >= 0), "index not sorted by mass"
    g_rec, g_sz, g_ids = canonical_entries(got)
    e_rec, e_sz, e_ids = canonical_entries(exp)
    if not np.array_equal(g_rec, e_rec):
        bad = np.nonzero(g_rec != e_rec)[0]
        raise AssertionError(f"{len(bad)} entries differ, first at {bad[0]}: got {g_rec[bad[0]]} expected {e_rec[bad[0]]}")
    assert np.array_equal(g_sz, e_sz), "protein list sizes differ"
    assert np.array_equal(g_ids, e_ids), "protein id lists differ"


def assert_emitted_equal(got: dict, exp: dict):
    assert len(got["mass"]) == len(exp["mass"]), (len(got["mass"]), len(exp["mass"]))
    for k in ("prot", "off", "len"):
        if not np.array_equal(got[k], exp[k]):
            bad = np.nonzero(got[k] != exp[k])[0]
            raise AssertionError(f"emitted {k} differs at {bad[0]} ({len(bad)} total): {got[k][bad[0]]} vs {exp[k][bad[0]]}")
    assert np.array_equal(bits(got["mass"]), bits(exp["mass"])), "emitted masses not bit-identical"


def pack(seqs):
    """list of str -> (residues, offsets)"""
    residues = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
    return residues, offsets
