"""CPU tests of the oracle (oracle/dbindex_oracle.cpp): known-answer tests derived by hand from
the reference's code, the store-level literals of DBIndexStoreSQLiteMult.main, and a cross-check
against an independently written set formulation (tests/pyref.py).

PARITY UNPINNED: the reference holds no golden vectors for this path (src/test/.gitignore), so
these KATs are hand-derived from the cited lines, not reference outputs."""
import json
import os
import struct

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import dbindex_b200 as dbi
from dbindex_b200 import synth
from dbindex_b200.indexer import get_residues, merge_intervals, tolerance_in_dalton, MassRange
from oracle import oracle_py
from oracle.oracle_py import Oracle

from . import pyref
from .util import PARAM_SETS, bits, pack

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def run_oracle(params, seqs, threads=1):
    o = Oracle(params, threads=threads)
    o.add_proteins(*pack(seqs))
    assert o.build() == 0
    return o


def test_kat_albumin_prefix():
    """MKWVTFISLLLLFSSAYSRGVFRR, trypsin, 2 missed cleavages (SURVEY.md 8c KAT).
    Fragments MK | WVTFISLLLLFSSAYSR | GVFR | R.  From start 0: MK..SR (1 MC), MK..GVFR (2 MC); the
    next end has 3 MC -> break (DBIndexer.java:322).  From start 2: 0, 1 and 2 MC.  GVFR / GVFRR / R
    are shorter than MIN_PEP_LENGTH = 6 (Constants.java:10)."""
    p = dbi.default_params()
    seq = "MKWVTFISLLLLFSSAYSRGVFRR"
    o = run_oracle(p, [seq])
    e = o.emitted()
    got = [(int(a), int(b)) for a, b in zip(e["off"], e["len"])]
    assert got == [(0, 19), (0, 23), (2, 17), (2, 21), (2, 22)]
    for (off, ln), m in zip(got, e["mass"]):
        assert m == pyref.seq_mass(p, seq[off:off + ln])  # bit-identical: same summation order
        assert m == o.calculate_mass(seq[off:off + ln].encode())  # IndexUtil.calculateMass == cutSeq mass
    # monoisotopic mass of WVTFISLLLLFSSAYSR + H2O + proton, independent arithmetic
    assert abs(e["mass"][2] - 2003.1000) < 2e-3  # hand sum of the 17 residue masses + 19.0178


def test_kat_cterm_free_missed_cleavage():
    """Q2: the limit is tested only when checkCleavage is true and counts enzyme residues - 1, so a
    protein-C-terminal peptide that does not end in K/R gets one extra missed cleavage."""
    p = dbi.default_params(max_missed=0, min_mass=100.0)
    seq = "AAAAAAKBBBBBBKCCCCCC"  # B, C: mass table has B, 'C' = Cys
    e = run_oracle(p, [seq]).emitted()
    got = sorted((int(a), int(b)) for a, b in zip(e["off"], e["len"]))
    # 0 MC fragments: AAAAAAK (0,7), BBBBBBK (7,7), CCCCCC (14,6); plus BBBBBBKCCCCCC (7,13): one
    # enzyme residue inside, count-1 = 0 <= 0 because the C-terminal end has no K/R
    assert got == [(0, 7), (7, 7), (7, 13), (14, 6)]


def test_kat_nocut_and_semi():
    p = dbi.default_params(nocut="P", max_missed=0, min_mass=100.0)
    seq = "AAAAAAKPAAAAAKAAAAAA"
    e = run_oracle(p, [seq]).emitted()
    got = sorted((int(a), int(b)) for a, b in zip(e["off"], e["len"]))
    # K at 6 is followed by P: not a cleavage site, but still counted as a missed cleavage (Q2):
    # AAAAAAKPAAAAAK has 2 enzyme residues -> 1 MC > 0 -> break; start 0 yields nothing.
    # start 14 (after K13): AAAAAA (14,6).
    assert got == [(14, 6)]
    p2 = dbi.default_params(semi=1, max_missed=0, min_mass=100.0, min_len=6)
    seq2 = "GGGGGGGKGG"
    e2 = run_oracle(p2, [seq2]).emitted()
    exp = pyref.digest_set(p2, [seq2])
    assert [(int(a), int(b)) for a, b in zip(e2["off"], e2["len"])] == [(s, l) for _, s, l, _ in exp]
    assert (0, 6) in [(s, l) for _, s, l, _ in exp] and (1, 7) in [(s, l) for _, s, l, _ in exp]


def test_store_level_kat_mult_main():
    """The literals of DBIndexStoreSQLiteMult.main (DBIndexStoreSQLiteMult.java:497-571)."""
    p = dbi.default_params(min_mass=0.0, max_mass=8000.0)
    f32 = lambda x: struct.unpack("f", struct.pack("f", x))[0]
    o = Oracle(p)
    o.add_proteins(*pack(["ABCDEFGHIJKL", "GHIJKLMNOPR"]))
    mass = [1.0, 2.0, 3.0, 4.0, f32(6000.42323), f32(6999.42323), 3.0, 3.0, 5.0, 3.0, 3.0]
    prot = [0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1]
    off = [0, 0, 0, 0, 0, 0, 6, 1, 2, 0, 0]
    ln = [1, 2, 3, 4, 5, 6, 3, 3, 5, 3, 3]
    assert o.build_from_records(mass, prot, off, ln) == 0
    c = o.counts()
    assert c["n_emitted"] == 11 and c["n_unique"] == 9
    e = o.entries()
    seqs = ["ABCDEFGHIJKL", "GHIJKLMNOPR"]
    pep = [seqs[pr][of:of + l] for pr, of, l in zip(e["first_prot"], e["first_off"], e["len"])]
    assert sorted(pep) == sorted(["A", "AB", "ABC", "GHI", "HIJ", "ABCD", "IJKLM", "ABCDE", "ABCDEF"])
    i = pep.index("GHI")  # first occurrence protein 0 offset 6; every occurrence listed, duplicates kept
    assert (e["first_prot"][i], e["first_off"][i]) == (0, 6)
    assert e["prot_ids"][int(e["prot_list_off"][i]):int(e["prot_list_off"][i + 1])].tolist() == [0, 1, 1]
    # getSequences(10, 8.9f): [1.1000004, 18.8999996], inclusive
    tol = f32(8.9)
    b, cnt, contig = o.query([max(0.0, 10 - tol)], [10 + tol])
    assert contig and cnt[0] == 6
    assert sorted(pep[int(b[0]):int(b[0] + cnt[0])]) == sorted(["AB", "ABC", "GHI", "HIJ", "ABCD", "IJKLM"])
    # 4 ranges {6+-1, 2+-1, 6+-1, 6+-1.2f} -> [1,3] U [4.8,7.2]
    lo, hi = oracle_py.merge_intervals([6.0, 2.0, 6.0, 6.0], [1.0, 1.0, 1.0, f32(1.2)])
    assert lo.tolist() == [1.0, 6.0 - f32(1.2)] and hi.tolist() == [3.0, 6.0 + f32(1.2)]
    b, cnt, _ = o.query(lo, hi)
    got = [x for bb, cc in zip(b, cnt) for x in pep[int(bb):int(bb + cc)]]
    assert sorted(got) == sorted(["A", "AB", "ABC", "GHI", "HIJ", "IJKLM"])


def test_isomers_stay_separate():
    """H5: equal mass != equal peptide.  AGK... permutations have bit-identical masses only when
    the additions round the same way; either way they must remain separate entries."""
    p = dbi.default_params(min_mass=300.0, max_missed=0)
    o = run_oracle(p, ["AAGGLLK", "GGAALLK", "AAGGLLK"])
    e = o.entries()
    assert o.counts()["n_unique"] == 2
    sizes = np.diff(e["prot_list_off"].astype(np.int64)).tolist()
    assert sorted(sizes) == [1, 2]


@pytest.mark.parametrize("name", list(PARAM_SETS))
def test_oracle_matches_set_formulation(name):
    p = dbi.default_params(**PARAM_SETS[name])
    res, off = synth.synth_proteome(12, 777, median_len=120, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs += ["", "K", "KKKKKKKKKKKK", "MSTYMSTYMSTYK", seqs[0]]  # empty, tiny, poly-K, mod-rich, duplicate
    o = run_oracle(p, seqs)
    exp = pyref.digest_set(p, seqs)
    e = o.emitted()
    assert len(exp) == len(e["mass"])
    assert [(int(a), int(b), int(c)) for a, b, c in zip(e["prot"], e["off"], e["len"])] == [(a, b, c) for a, b, c, _ in exp]
    assert np.array_equal(bits(e["mass"]), bits([m for *_, m in exp]))
    # unique peptides + variants
    uniq = {}
    for pid, s, l, m in exp:
        uniq.setdefault(seqs[pid][s:s + l], (m, pid, s, []))[3].append(pid)
    ent = o.entries()
    assert o.counts()["n_unique"] == len(uniq)
    exp_entries = []
    for pepseq, (m, pid, s, plist) in uniq.items():
        for vm, pos in pyref.expand_set(p, pepseq, m):
            pat = sum((q + 1) << (8 * k) for k, q in enumerate(pos))
            exp_entries.append((np.float64(vm).view(np.uint64).item(), pid, s, len(pepseq), pat, tuple(plist)))
    got_entries = []
    plo = ent["prot_list_off"].astype(np.int64)
    for i in range(len(ent["mass"])):
        got_entries.append((int(bits(ent["mass"])[i]), int(ent["first_prot"][i]), int(ent["first_off"][i]),
                            int(ent["len"][i]), int(ent["modpat"][i]), tuple(ent["prot_ids"][plo[i]:plo[i + 1]].tolist())))
    assert sorted(got_entries) == sorted(exp_entries)
    assert np.all(np.diff(ent["mass"]) >= 0)


def test_mod_groups_match_reference_multisets():
    """SURVEY.md 8c cross-check: the multiset of shift sums produced by the expansion equals what the
    reference's modGroupList (multisets of distinct shifts of size 1..max, SearchParamReader.java:671-687)
    enumerates for feasible groups."""
    from itertools import combinations_with_replacement
    p = dbi.default_params(diff_mods=[("M", 15.9949), ("STY", 79.96633)], max_mods_per_peptide=3, min_mass=100.0)
    pep = "MSTYMK"
    o = run_oracle(p, [pep])
    ent = o.entries()
    base = ent["mass"][ent["modpat"] == 0][0]
    shifts = {round(m - base, 5) for m in ent["mass"] if m != base}
    groups = set()
    for k in (1, 2, 3):
        for g in combinations_with_replacement([15.9949, 79.96633], k):
            if g.count(15.9949) <= 2 and g.count(79.96633) <= 3:  # feasible: 2 M, 3 STY sites
                groups.add(round(sum(g), 5))
    assert shifts == groups


def test_query_semantics_inclusive_and_clamped():
    p = dbi.default_params()
    res, off = synth.synth_proteome(40, 5)
    o = Oracle(p)
    o.add_proteins(res, off)
    o.build()
    m = o.entries()["mass"]
    x = m[len(m) // 2]
    b, c, contig = o.query([x, x, np.nextafter(x, np.inf), 0.0], [x, np.nextafter(x, -np.inf), 9000.0, 1e9])
    assert contig
    assert c[0] == np.count_nonzero(m == x) and c[0] >= 1  # both ends inclusive (Merge:415-419)
    assert c[1] == 0
    assert c[2] == np.count_nonzero(m > x)
    assert c[3] == len(m)


def test_ppm_and_intervals_and_flanks():
    assert tolerance_in_dalton(1000.0, 10.0) == oracle_py.tolerance_in_dalton(1000.0, 10.0)
    assert abs(tolerance_in_dalton(1000.0, 10.0) - 0.0099999) < 1e-6
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(1, 8))
        mass = rng.uniform(0, 20, n)
        tol = rng.uniform(0, 3, n)
        lo, hi = oracle_py.merge_intervals(mass, tol)
        got = merge_intervals([MassRange(a, b) for a, b in zip(mass, tol)])
        assert [g[0] for g in got] == lo.tolist() and [g[1] for g in got] == hi.tolist()
    prot = "ABCDEFGHIJ"
    # Q8: right flank is one residue short near the C-terminus (Util.java:140-146)
    assert get_residues(0, 3, prot) == ("---", "DEF")
    assert get_residues(2, 3, prot) == ("-AB", "FGH")
    assert get_residues(4, 3, prot) == ("BCD", "HI-")   # 3 residues left, only 2 shown
    assert get_residues(6, 3, prot) == ("DEF", "---")   # 1 residue left, none shown
    assert get_residues(7, 3, prot) == ("EFG", "---")
    for off, ln in [(0, 3), (2, 3), (4, 3), (6, 3), (7, 3), (5, 5), (0, 10)]:
        assert get_residues(off, ln, prot) == oracle_py.get_residues(prot.encode(), off, ln)


def test_oracle_threads_equal_single_thread():
    p = dbi.default_params(**PARAM_SETS["cfg2_mods"])
    res, off = synth.synth_proteome(150, 11)
    a = Oracle(p, threads=1); a.add_proteins(res, off); a.build()
    b = Oracle(p, threads=4); b.add_proteins(res, off); b.build()
    ea, eb = a.entries(), b.entries()
    for k in ea:
        assert np.array_equal(ea[k], eb[k]), k


def test_golden_fixture():
    """Regression fixture produced by tests/golden/make_golden.py from the oracle (committed with the
    script).  Guards the oracle itself against drift."""
    with open(os.path.join(GOLDEN, "small_index.json")) as f:
        g = json.load(f)
    p = dbi.default_params(**PARAM_SETS[g["param_set"]])
    o = run_oracle(p, g["proteins"])
    e = o.entries()
    got = [[int(x) for x in bits(e["mass"])], e["first_prot"].tolist(), e["first_off"].tolist(), e["len"].tolist(),
           e["modpat"].tolist(), e["prot_list_off"].tolist(), e["prot_ids"].tolist()]
    assert got == g["entries"]


aa = st.sampled_from(list("ACDEFGHIKLMNPQRSTVWY"))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.text(aa, min_size=0, max_size=60), min_size=1, max_size=4), st.integers(0, 3), st.booleans(),
       st.sampled_from(["", "P", "PG"]))
def test_property_digest(seqs, mc, semi, nocut):
    p = dbi.default_params(max_missed=mc, semi=int(semi), nocut=nocut, min_mass=400.0, max_mass=3000.0)
    o = run_oracle(p, seqs)
    exp = pyref.digest_set(p, seqs)
    e = o.emitted()
    assert [(int(a), int(b), int(c)) for a, b, c in zip(e["prot"], e["off"], e["len"])] == [(a, b, c) for a, b, c, _ in exp]
    assert np.array_equal(bits(e["mass"]), bits([m for *_, m in exp]))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.text(aa, min_size=0, max_size=60), min_size=1, max_size=4), st.integers(0, 4), st.booleans(),
       st.sampled_from([None, "", "K", "KR", "C"]), st.sampled_from([None, ("K", 0), ("K", 1), ("L", 2)]))
def test_property_digest_filters(seqs, mc, semi, mandatory, pep_filter):
    """mandatoryInternalAAs (break / skip semantics) and PeptideFilterByMaxOccurrencies; an
    empty (non-null) mandatory array emits nothing, like the reference."""
    kw = dict(max_missed=mc, semi=int(semi), min_mass=400.0, max_mass=3000.0)
    if mandatory is not None:
        kw["mandatory_internal"] = mandatory
    if pep_filter is not None:
        kw["peptide_filter"] = pep_filter
    p = dbi.default_params(**kw)
    o = run_oracle(p, seqs)
    exp = pyref.digest_set(p, seqs)
    e = o.emitted()
    assert [(int(a), int(b), int(c)) for a, b, c in zip(e["prot"], e["off"], e["len"])] == [(a, b, c) for a, b, c, _ in exp]
    assert np.array_equal(bits(e["mass"]), bits([m for *_, m in exp]))
    if mandatory == "":
        assert len(exp) == 0


def test_kat_mandatory_internal_and_filter():
    """Hand-derived from DBIndexer.java:310-313,334-344 and DBIndexStoreSQLiteMult.java:245-263.
    Protein AAAAAAKAAAAAAR, trypsin, 2 missed cleavages, min mass 300: windows from start 0 are
    AAAAAAK (K only as the LAST residue: found, but the store SKIPs it) and AAAAAAKAAAAAAR (K internal:
    kept); start 7 gives AAAAAAR -- no K at all, so the start ends there."""
    p = dbi.default_params(mandatory_internal="K", min_mass=300.0)
    o = run_oracle(p, ["AAAAAAKAAAAAAR"])
    e = o.emitted()
    assert [(int(a), int(b)) for a, b in zip(e["off"], e["len"])] == [(0, 14)]
    # without the mandatory set: all three windows
    o2 = run_oracle(dbi.default_params(min_mass=300.0), ["AAAAAAKAAAAAAR"])
    e2 = o2.emitted()
    assert [(int(a), int(b)) for a, b in zip(e2["off"], e2["len"])] == [(0, 7), (0, 14), (7, 7)]
    # occurrence filter A <= 6: the 14-residue window holds 12 A, the walk from start 0 ends at the 7th A
    o3 = run_oracle(dbi.default_params(peptide_filter=("A", 6), min_mass=300.0), ["AAAAAAKAAAAAAR"])
    e3 = o3.emitted()
    assert [(int(a), int(b)) for a, b in zip(e3["off"], e3["len"])] == [(0, 7), (7, 7)]


def test_ppm_probe_loop_restatement():
    """DBIndexer.getSequencesUsingPPMTolerance (DBIndexer.java:787-844) restated in the oracle: the result
    starts with the Dalton answer and only ever grows by entries at the probed exact masses."""
    p = dbi.default_params(min_mass=400.0, max_mass=3000.0)
    res, off = synth.synth_proteome(30, 4242, median_len=150, min_len=20)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    o = run_oracle(p, seqs)
    ent = o.entries()
    for m in ent["mass"][:: max(1, len(ent["mass"]) // 25)]:
        for ppm in (5.0, 50.0):
            idx, probes = o.query_ppm(float(m) * (1 - 2e-6), ppm)
            tol = tolerance_in_dalton(float(m) * (1 - 2e-6), ppm)
            b, c, _ = o.query(np.array([max(0.0, float(m) * (1 - 2e-6) - tol)]), np.array([float(m) * (1 - 2e-6) + tol]))
            assert idx[:int(c[0])].tolist() == list(range(int(b[0]), int(b[0] + c[0])))
            assert len(set(idx.tolist())) == len(idx) and probes >= 1


def test_oracle_matches_a_sqlite_restatement_of_the_reference_store():
    """The oracle's rows / merge / query against tests/sqlite_ref.py: the reference's table, records
    and SQL on a real SQLite B-tree, fed with the same addSequence calls."""
    import struct

    from .sqlite_ref import SqliteStore
    p = dbi.default_params()
    res, off = synth.synth_proteome(60, 4242, median_len=220, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs += [seqs[0], seqs[7], "AAGGLLKGGAALLKAAGGLLK", "GALKAGLKLAGKGLAK", seqs[0]]  # duplicates and isomers
    o = Oracle(p)
    o.add_proteins(*pack(seqs))
    assert o.build() == 0
    em = o.emitted()
    store = SqliteStore(seqs, p.mass_group_factor)
    for m, pr, of, ln in zip(em["mass"], em["prot"], em["off"], em["len"]):
        store.add_sequence(float(m), int(of), int(ln), int(pr))
    store.stop_add_seq()

    def canon(entries):
        return sorted((struct.pack("<d", m), s, ids) for m, s, ids in entries)

    e = o.entries()
    plo = e["prot_list_off"].astype(np.int64)
    mine = [(float(e["mass"][i]), seqs[int(e["first_prot"][i])][int(e["first_off"][i]):int(e["first_off"][i]) + int(e["len"][i])],
             tuple(int(x) for x in e["prot_ids"][plo[i]:plo[i + 1]])) for i in range(len(e["mass"]))]
    assert canon(store.all_entries()) == canon(mine)
    # queries: 10 ppm windows around indexed masses, wide windows, decoys, zero tolerance
    rng = np.random.default_rng(3)
    qm = np.concatenate([e["mass"][rng.integers(0, len(e["mass"]), 40)], rng.uniform(600, 6000, 20)])
    tol = np.concatenate([qm[:20] * 1e-5, np.zeros(20), np.full(10, 3.0), np.full(10, 0.02)])
    lo, hi = np.maximum(qm - tol, 0.0), qm + tol
    b, c, contig = o.query(lo, hi)
    assert contig
    for k in range(len(qm)):
        got = canon(store.get_sequences(float(qm[k]), float(tol[k])))
        exp = canon(mine[int(b[k]):int(b[k] + c[k])])
        assert got == exp, (k, qm[k], tol[k], len(got), len(exp))


@pytest.mark.parametrize("name", ["cfg1_tryptic", "cfg2_mods", "semi_nocut_mods"])
def test_oracle_query_hits_match_entries_and_python_flanks(name):
    """orc_query_hits (the checker of dbi_query_hits) against an independent formulation: the hits of [lo, hi] are
    the entries with lo <= mass <= hi (both ends inclusive), each with the peptide cut from its first
    protein (ProteinCache.getPeptideSequence), the flanks of the Python restatement of Util.getResidues
    (dbindex_b200/indexer.get_residues, Util.java:130-162 incl. the right-side off-by-one), its mod pattern and
    its protein list."""
    from dbindex_b200.indexer import get_residues
    p = dbi.default_params(**PARAM_SETS[name])
    res, off = synth.synth_proteome(25, 2024, median_len=160, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs += [seqs[1], "MSTYMSTYK", "K", ""]
    o = run_oracle(p, seqs)
    e = o.entries()
    n = len(e["mass"])
    plo = e["prot_list_off"].astype(np.int64)
    rng = np.random.default_rng(12)
    qm = np.concatenate([e["mass"][rng.integers(0, n, 25)], rng.uniform(600, 6000, 10), [e["mass"][0], e["mass"][-1]]])
    tol = np.concatenate([qm[:10] * 1e-5, np.zeros(15), np.full(10, 3.0), [0.0, 0.0]])
    lo, hi = np.maximum(qm - tol, 0.0), qm + tol
    h = o.query_hits(lo, hi)
    ho, so, po = (h[k].astype(np.int64) for k in ("hit_off", "seq_off", "prot_list_off"))
    fl = h["flanks"].reshape(-1, 6)
    assert ho[0] == 0 and ho[-1] == len(h["mass"])
    for q in range(len(qm)):
        want = []
        for i in np.nonzero((e["mass"] >= lo[q]) & (e["mass"] <= hi[q]))[0]:
            prot, o_, l_ = int(e["first_prot"][i]), int(e["first_off"][i]), int(e["len"][i])
            left, right = get_residues(o_, l_, seqs[prot])
            want.append((int(bits(e["mass"][i:i + 1])[0]), prot, o_, l_, int(e["modpat"][i]), seqs[prot][o_:o_ + l_],
                         left + right, tuple(int(x) for x in e["prot_ids"][plo[i]:plo[i + 1]])))
        got = [(int(bits(h["mass"][i:i + 1])[0]), int(h["first_prot"][i]), int(h["first_off"][i]), int(h["len"][i]),
                int(h["modpat"][i]), h["seq"][so[i]:so[i + 1]].tobytes().decode(), fl[i].tobytes().decode(),
                tuple(int(x) for x in h["prot_ids"][po[i]:po[i + 1]])) for i in range(ho[q], ho[q + 1])]
        assert sorted(got) == sorted(want), (q, qm[q], tol[q])
