"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bar: bit-exact (integer/index work and IEEE-double masses alike)."""
import os
import struct

import numpy as np
import pytest

import dbindex_b200 as dbi
from dbindex_b200 import synth
from dbindex_b200.indexer import DBIndexImpl, DBIndexer, MassRange, tolerance_in_dalton
from oracle.oracle_py import Oracle

from .util import PARAM_SETS, assert_emitted_equal, assert_entries_equal, bits, pack

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1


def both(params, residues, offsets, keep=True):
    params = params.copy()
    params.keep_emitted = 1 if keep else 0
    g = dbi.GpuIndex(params)
    g.add_proteins(residues, offsets)
    g.build()
    o = Oracle(params, threads=NCPU)
    o.add_proteins(residues, offsets)
    assert o.build() == 0
    return g, o


def check_all(g, o, nq=500, seed=1):
    st, oc = g.stats(), o.counts()
    assert (st["n_emitted"], st["n_unique"], st["n_entries"]) == (oc["n_emitted"], oc["n_unique"], oc["n_entries"])
    assert_emitted_equal(g.debug_emitted(), o.emitted())
    exp = o.entries()
    got = g.fetch(0, st["n_entries"])
    assert_entries_equal(got, exp)
    _, _, lo, hi = synth.synth_queries(exp["mass"], nq, seed, da_fraction=0.3)
    b, c = g.query(lo, hi)
    ob, oc_, contig = o.query(lo, hi)
    assert contig
    assert np.array_equal(b, ob), "hit_begin differs"
    assert np.array_equal(c, oc_), "hit_count differs"
    check_hits(g, o, lo[:200], hi[:200])


def hits_canonical(h):
    """Per query the multiset of materialised hits (ties inside a query have no contractual order)."""
    ho, so, po = (h[k].astype(np.int64) for k in ("hit_off", "seq_off", "prot_list_off"))
    fl = h["flanks"].reshape(-1, 6)
    out = []
    for q in range(len(ho) - 1):
        out.append(sorted(
            (int(bits(h["mass"][i:i + 1])[0]), int(h["first_prot"][i]), int(h["first_off"][i]), int(h["len"][i]),
             int(h["modpat"][i]), h["seq"][so[i]:so[i + 1]].tobytes(), fl[i].tobytes(),
             tuple(h["prot_ids"][po[i]:po[i + 1]].tolist())) for i in range(ho[q], ho[q + 1])))
    return out


def check_hits(g, o, lo, hi):
    """dbi_query_hits against the oracle's parseAddPeptideInfo restatement: mass bits, first occurrence,
    peptide string, flanks (with the reference's off-by-one), mod pattern and protein lists."""
    raw = g.query_hits(lo, hi, per_hit=False)
    got = dbi.capi.expand_hits(raw)
    exp = o.query_hits(lo, hi)
    assert np.array_equal(got["hit_off"], exp["hit_off"])
    assert hits_canonical(got) == hits_canonical(exp)
    # the run structure of the raw answer: every run has hits, runs never straddle a query, the two
    # CSRs over queries agree, and neighbouring runs of a query differ in peptide or mass (runs are maximal)
    c = raw["counts"]
    ho, po, pho = (raw[k].astype(np.int64) for k in ("hit_off", "pep_off", "pep_hit_off"))
    assert len(pho) == c.n_peps + 1 and pho[0] == 0 and pho[-1] == c.n_hits and np.all(np.diff(pho) > 0)
    assert po[0] == 0 and po[-1] == c.n_peps and np.array_equal(pho[po], ho)
    if c.n_peps > 1:
        key = np.stack([raw["first_prot"].astype(np.uint64), raw["first_off"].astype(np.uint64),
                        raw["len"].astype(np.uint64), bits(raw["mass"])])
        same = np.all(key[:, 1:] == key[:, :-1], axis=0)
        inner = np.ones(c.n_peps - 1, bool)
        inner[po[1:-1][(po[1:-1] > 0) & (po[1:-1] < c.n_peps)] - 1] = False  # a query boundary sits between them
        assert not np.any(same & inner)


def test_radix_sort_hook():
    """K7 alone: stability, tile boundaries, skewed digits."""
    rng = np.random.default_rng(0)
    with dbi.GpuIndex(dbi.default_params()) as g:
        for n in (1, 2, 31, 255, 4095, 4096, 4097, 8192, 100_003, 1_000_000):
            for kind in ("random", "equal", "few", "sorted", "reverse"):
                if kind == "random":
                    keys = rng.integers(0, 1 << 54, size=n, dtype=np.uint64)
                elif kind == "equal":
                    keys = np.full(n, 0x2A5A5A5A5A5A5A, dtype=np.uint64)
                elif kind == "few":
                    keys = rng.integers(0, 5, size=n, dtype=np.uint64) << np.uint64(20)
                elif kind == "sorted":
                    keys = np.arange(n, dtype=np.uint64) * np.uint64(977)
                else:
                    keys = (np.arange(n, dtype=np.uint64)[::-1] * np.uint64(977)).copy()
                vals = np.arange(n, dtype=np.uint64)
                k2, v2 = keys.copy(), vals.copy()
                g.debug_radix_sort(k2, v2, 0, 54)
                order = np.argsort(keys, kind="stable")
                assert np.array_equal(k2, keys[order]), (n, kind, "keys")
                assert np.array_equal(v2, vals[order]), (n, kind, "stability / values")
        # partial bit ranges: only bits [8, 20) take part
        keys = rng.integers(0, 1 << 30, size=50_000, dtype=np.uint64)
        vals = np.arange(50_000, dtype=np.uint64)
        k2, v2 = keys.copy(), vals.copy()
        g.debug_radix_sort(k2, v2, 8, 20)
        order = np.argsort((keys >> np.uint64(8)) & np.uint64(0xFFF), kind="stable")
        assert np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order])


def test_kat_albumin_prefix_gpu():
    p = dbi.default_params()
    seq = "MKWVTFISLLLLFSSAYSRGVFRR"
    p.keep_emitted = 1
    with dbi.GpuIndex(p) as g:
        g.add_proteins(*pack([seq]))
        g.build()
        e = g.debug_emitted()
        assert [(int(a), int(b)) for a, b in zip(e["off"], e["len"])] == [(0, 19), (0, 23), (2, 17), (2, 21), (2, 22)]
        for off, ln, m in zip(e["off"], e["len"], e["mass"]):
            assert m == g.calculate_mass(seq[off:off + ln].encode())  # zero-tolerance lookups work


@pytest.mark.parametrize("name", list(PARAM_SETS))
def test_parity_param_sets(name):
    p = dbi.default_params(**PARAM_SETS[name])
    res, off = synth.synth_proteome(120, 1000 + len(name), median_len=300, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    # edge cases the domain has: empty and tiny proteins, poly-K, duplicates (protein lists with
    # repeats), isomers, mod-rich peptides, a protein without any cleavage site
    seqs += ["", "K", "A", "KKKKKKKKKKKKKKKK", seqs[0], seqs[0], "AAGGLLKGGAALLKAAGGLLK", "MSTYMSTYMSTYMSTYK",
             "GGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG", "", seqs[3][:40], "PKPKPKPRPRPAAAAAAAAK",
             # modifiable sites beyond residue 64 of a peptide: the long-peptide expansion path
             "G" * 70 + "MSTYMWSTK" + "G" * 12 + "MR", "A" * 66 + "SSTTYYMMWK"]
    g, o = both(p, *pack(seqs))
    try:
        check_all(g, o)
    finally:
        g.close()


@pytest.mark.parametrize("name", ["cfg1_tryptic", "cfg2_mods"])
def test_query_hits_batch_shapes(name):
    """dbi_query_hits on the shapes a batch can take: the empty batch, only misses, windows of thousands of
    hits (several 256-hit segments, runs that straddle a segment boundary) next to windows of none, a
    window over the whole index, duplicated and degenerate ranges."""
    p = dbi.default_params(**PARAM_SETS[name])
    res, off = synth.synth_proteome(150, 4242, median_len=300, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs += [seqs[0], seqs[1], "MSTYMSTYMSTYMSTYK", "MSTYMSTYMSTYMSTYK"]
    g, o = both(p, *pack(seqs))
    try:
        m = o.entries()["mass"]
        empty = g.query_hits(np.zeros(0), np.zeros(0), per_hit=False)
        assert empty["counts"].n_hits == 0 and empty["counts"].n_peps == 0 and empty["hit_off"].tolist() == [0]
        assert empty["pep_off"].tolist() == [0] and empty["pep_hit_off"].tolist() == [0]
        lo = np.array([10.0, 7000.0, 599.0])
        hi = np.array([20.0, 7999.0, 599.5])
        check_hits(g, o, lo, hi)  # misses only
        mid = float(np.median(m))
        lo = np.array([0.0, mid - 3.0, 5.0, mid, m[0], m[-1], mid + 1.0, mid - 50.0, mid - 3.0, 6500.0])
        hi = np.array([8000.0, mid + 3.0, 6.0, mid, m[0], m[-1], mid - 1.0, mid + 50.0, mid + 3.0, 6400.0])
        check_hits(g, o, lo, hi)
    finally:
        g.close()


def test_wide_sequence_hash_path(monkeypatch):
    """Large builds sort a sequence hash of more than 32 bits (the number of equal-mass string pairs
    grows with N^2); force that path on a small input with duplicates and isomers."""
    monkeypatch.setenv("DBI_HASH_BITS", "44")
    p = dbi.default_params(**PARAM_SETS["cfg2_mods"])
    res, off = synth.synth_proteome(150, 77, median_len=250, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs += [seqs[0], seqs[5], "AAGGLLKGGAALLKAAGGLLK", "GALKAGLKLAGKGLAK", seqs[0]]
    g, o = both(p, *pack(seqs))
    try:
        check_all(g, o)
        assert g.stats()["sort_bits_base"] >= 44
    finally:
        g.close()


def test_empty_and_degenerate_indexes():
    p = dbi.default_params()
    for seqs in ([], [""], ["AAA"], ["K" * 5], ["G" * 200]):
        with dbi.GpuIndex(p) as g:
            g.add_proteins(*pack(seqs))
            g.build()
            st = g.stats()
            o = Oracle(p); o.add_proteins(*pack(seqs)); o.build()
            assert st["n_entries"] == o.counts()["n_entries"]
            b, c = g.query(np.array([0.0, 700.0]), np.array([9000.0, 701.0]))
            ob, oc_, _ = o.query(np.array([0.0, 700.0]), np.array([9000.0, 701.0]))
            assert np.array_equal(c, oc_) and np.array_equal(b, ob)


def test_long_protein_and_many_small():
    """A titin-sized protein (work unit is a start residue, not a protein) next to many short ones,
    added in several dbi_add_proteins calls."""
    p = dbi.default_params(**PARAM_SETS["cfg2_mods"])
    p.keep_emitted = 1
    big, boff = synth.synth_proteome(1, 31, median_len=36000, sigma=0.0, max_len=36000)
    small, soff = synth.synth_proteome(400, 32, median_len=60, min_len=1)
    g = dbi.GpuIndex(p)
    o = Oracle(p, threads=NCPU)
    for r, f in ((small[:int(soff[200])], soff[:201]), (big, boff), (small[int(soff[200]):], soff[200:] - soff[200])):
        g.add_proteins(r, f)
        o.add_proteins(r, f)
    g.build()
    assert o.build() == 0
    try:
        check_all(g, o)
    finally:
        g.close()


def test_store_level_kat_mult_main_gpu():
    """DBIndexStoreSQLiteMult.main literals (DBIndexStoreSQLiteMult.java:497-571) through the GPU store."""
    p = dbi.default_params(min_mass=0.0, max_mass=8000.0)
    f32 = lambda x: struct.unpack("f", struct.pack("f", x))[0]
    seqs = ["ABCDEFGHIJKL", "GHIJKLMNOPR"]
    mass = [1.0, 2.0, 3.0, 4.0, f32(6000.42323), f32(6999.42323), 3.0, 3.0, 5.0, 3.0, 3.0]
    prot = [0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1]
    off = [0, 0, 0, 0, 0, 0, 6, 1, 2, 0, 0]
    ln = [1, 2, 3, 4, 5, 6, 3, 3, 5, 3, 3]
    with dbi.GpuIndex(p) as g:
        g.add_proteins(*pack(seqs))
        g.build_from_records(mass, prot, off, ln)
        o = Oracle(p); o.add_proteins(*pack(seqs)); o.build_from_records(mass, prot, off, ln)
        e = g.fetch(0, g.stats()["n_entries"])
        assert_entries_equal(e, o.entries())
        pep = [seqs[a][b:b + c] for a, b, c in zip(e["first_prot"], e["first_off"], e["len"])]
        i = pep.index("GHI")
        assert e["prot_ids"][int(e["prot_list_off"][i]):int(e["prot_list_off"][i + 1])].tolist() == [0, 1, 1]
        tol = f32(8.9)
        b, c = g.query(np.array([max(0.0, 10 - tol)]), np.array([10 + tol]))
        assert sorted(pep[int(b[0]):int(b[0] + c[0])]) == sorted(["AB", "ABC", "GHI", "HIJ", "ABCD", "IJKLM"])


def test_export_reference_sqlite_index(tmp_path):
    """the GPU-built index written as the reference's own bucket directory; its rows, read back
    through the row walk of parseAddPeptideInfo, are the oracle's unmodified entries (cfg2: variants are skipped)."""
    from dbindex_b200.sqlite_export import read_rows
    p = dbi.default_params(**PARAM_SETS["cfg2_mods"])
    res, off = synth.synth_proteome(80, 31, median_len=250, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs += [seqs[2], seqs[2]]
    ix = DBIndexer(p)
    ix.init()
    ix.run(proteins=(synth.deflines(len(seqs)), seqs))
    try:
        info = ix.export_reference_index(str(tmp_path / "proteome.fasta_0123"))
        o = Oracle(p); o.add_proteins(*pack(seqs)); assert o.build() == 0
        e = o.entries()
        plo = e["prot_list_off"].astype(np.int64)
        want = sorted((int(bits(e["mass"][i:i + 1])[0]), int(e["first_off"][i]), int(e["len"][i]),
                       tuple(int(x) for x in e["prot_ids"][plo[i]:plo[i + 1]]))
                      for i in range(len(e["mass"])) if e["modpat"][i] == 0)
        got = sorted((int(bits(np.array([m]))[0]), o_, l_, ids) for _, peps in read_rows(info["dir"]) for m, o_, l_, ids in peps)
        assert info["peptides"] == len(want) == o.counts()["n_unique"]
        assert got == want
    finally:
        ix.close()


def test_reference_api_mirror():
    """The DBIndexImpl / DBIndexer surface (DBIndexImpl.java:180-237,501-513; DBIndexer.java:762-947)."""
    p = dbi.default_params()
    res, off = synth.synth_proteome(60, 77)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs.append(seqs[5])  # a duplicated protein: every peptide of it maps to two proteins
    defs = synth.deflines(len(seqs))
    impl = DBIndexImpl(p, proteins=(defs, seqs))
    try:
        o = Oracle(p); o.add_proteins(*pack(seqs)); o.build()
        exp = o.entries()
        m = float(exp["mass"][len(exp["mass"]) // 3])
        tol = tolerance_in_dalton(m, 10.0)
        hits = impl.getSequences(m, tol)
        ob, oc, _ = o.query([m - tol], [m + tol])
        assert len(hits) == int(oc[0]) >= 1
        for h in hits:
            assert abs(h.mass - m) <= tol
            pid = h.proteinIds[0]
            assert seqs[pid][h.sequenceOffset:h.sequenceOffset + h.sequenceLen] == h.sequence
            assert impl.indexer.index.calculate_mass(h.sequence.encode()) == h.mass
            assert len(h.resLeft) == 3 and len(h.resRight) == 3
        # getProteins(String): zero-tolerance lookup by recomputed mass (bit-exact masses required)
        target = hits[0]
        prots = impl.getProteins(target.sequence)
        assert {q.id for q in prots} == set(target.proteinIds)
        assert all(q.accession == defs[q.id] for q in prots)
        assert impl.getProteinSequenceById(3) == seqs[3]
        assert impl.getIndexedProteinById(3).accession == defs[3]
        # multi-range: union over merged intervals, no duplicates (MergeIntervals.java:31)
        rs = [MassRange(m, 0.5), MassRange(m + 0.2, 0.5), MassRange(m, 0.5), MassRange(m + 300, 0.01)]
        multi = impl.getSequences(rs)
        keys = [s.key() for s in multi]
        assert len(keys) == len(set(keys))
        lo = min(m - 0.5, m - 0.3)
        n_exp = np.count_nonzero((exp["mass"] >= lo) & (exp["mass"] <= m + 0.7)) + \
            np.count_nonzero((exp["mass"] >= m + 300 - 0.01) & (exp["mass"] <= m + 300 + 0.01))
        assert len(multi) == n_exp
        # queries reaching MAX_MASS return nothing (DBIndexStoreSQLiteMult.java:333-338)
        assert impl.getSequences(7990.0, 20.0) == []
        # ppm entry point of DBIndexer
        ppm_hits = impl.indexer.getSequencesUsingPPMTolerance(m, 10.0)
        assert {s.key() for s in ppm_hits} >= {s.key() for s in hits}
        # ... and equal to the oracle's restatement of the probe loop (DBIndexer.java:810-839), in list order
        ent_key = lambda i: (int(bits(exp["mass"][i:i + 1])[0]), int(exp["first_prot"][i]), int(exp["first_off"][i]),
                             int(exp["len"][i]))  # noqa: E731
        for mm, ppm in ((m, 10.0), (m * (1 + 4e-6), 5.0), (float(exp["mass"][7]) - 1e-7, 50.0), (1234.5678, 20.0)):
            got_ppm = impl.indexer.getSequencesUsingPPMTolerance(mm, ppm)
            idx, probes = o.query_ppm(mm, ppm)
            assert len(got_ppm) == len(idx)
            got_k = [(s.key()[0], s.proteinIds[0], s.sequenceOffset, s.sequenceLen) for s in got_ppm]
            # list order: the Dalton answer first, then what every probe adds; ties inside one query are free
            assert sorted(got_k) == sorted(ent_key(int(i)) for i in idx)
        assert impl.indexer.getNumberSequences() == len(exp["mass"])
        pm = impl.indexer.getParentMasses()
        assert pm == sorted(set(int(x * 10000) / 10000.0 for x in exp["mass"]))
    finally:
        impl.close()
    d = DBIndexer(p)
    with pytest.raises(Exception):
        d.getSequencesUsingDaltonTolerance(1000.0, 1.0)  # "Indexer is not initialized"


@pytest.mark.parametrize("name", ["cfg1_tryptic", "cfg2_mods", "many_classes"])
def test_save_load_roundtrip(name, tmp_path):
    """dbi_save / dbi_load -- a fresh handle that loads the file answers exactly like the one
    that built the index (entries, protein lists, queries, materialised hits, proteins), and other search
    parameters are refused."""
    p = dbi.default_params(**PARAM_SETS[name])
    res, off = synth.synth_proteome(300, 99, median_len=200, min_len=5)
    g, o = both(p, res, off, keep=False)
    path = str(tmp_path / "index.gpuidx")
    try:
        g.save(path)
        g2 = dbi.GpuIndex(p)
        try:
            g2.load(path)
            st, st2 = g.stats(), g2.stats()
            for k in ("n_proteins", "n_residues", "n_emitted", "n_unique", "n_entries"):
                assert st[k] == st2[k], k
            assert_entries_equal(g2.fetch(0, st2["n_entries"]), o.entries())
            _, _, lo, hi = synth.synth_queries(o.entries()["mass"], 200, 4, da_fraction=0.3)
            check_hits(g2, o, lo, hi)
            assert g2.get_protein(7) == res[int(off[7]):int(off[8])].tobytes()
            with pytest.raises(dbi.DbiError):
                g2.load(path)  # not a fresh handle any more
        finally:
            g2.close()
        q = dbi.default_params(**dict(PARAM_SETS[name], min_mass=p.min_mass + 1.0))
        g3 = dbi.GpuIndex(q)
        try:
            with pytest.raises(dbi.DbiError) as ei:
                g3.load(path)
            assert "other search parameters" in str(ei.value)
        finally:
            g3.close()
    finally:
        g.close()


def test_resume_from_disk_through_the_mirror(tmp_path):
    """DBIndexImpl with inMemoryIndex = false: the first instance indexes and writes
    <fasta>_<md5(params)>.gpuidx, the second finds it and skips indexing (DBIndexer.java:522-531)."""
    from dbindex_b200.indexer import INDEX_FILE_SUFFIX, createFullIndexFileName, getDefaultDBIndexParams
    res, off = synth.synth_proteome(80, 5, median_len=150, min_len=5)
    fasta = str(tmp_path / "small.fasta")
    synth.write_fasta(fasta, res, off)
    sp = getDefaultDBIndexParams(fasta, inMemoryIndex=False, use_mono=True, max_missed=2)
    name = createFullIndexFileName(sp)
    assert name.startswith(fasta + "_") and len(name) == len(fasta) + 33
    a = DBIndexImpl(sp)
    try:
        assert os.path.exists(name + INDEX_FILE_SUFFIX) and not a.indexer.loaded_from_disk
        hits_a = [(s.mass, s.sequence, tuple(s.proteinIds)) for s in a.getSequences(1500.0, 2.0)]
    finally:
        a.close()
    b = DBIndexImpl(sp)
    try:
        assert b.indexer.loaded_from_disk
        assert [(s.mass, s.sequence, tuple(s.proteinIds)) for s in b.getSequences(1500.0, 2.0)] == hits_a
        assert b.getIndexedProteinById(3).accession.startswith("sp|S0000003")
    finally:
        b.close()


def test_lifecycle_errors_and_rebuild():
    p = dbi.default_params()
    res, off = synth.synth_proteome(50, 9)
    with dbi.GpuIndex(p) as g:
        with pytest.raises(dbi.DbiError) as ei:
            g.query(np.array([1.0]), np.array([2.0]))
        assert ei.value.code == -1  # DBI_ENOTINIT
        g.add_proteins(res, off)
        g.build()
        n1 = g.stats()["n_entries"]
        first = g.fetch(0, n1)
        with pytest.raises(dbi.DbiError) as ei:
            g.build()
        assert ei.value.code == -2  # DBI_EALREADY
        with pytest.raises(dbi.DbiError):
            g.add_proteins(res, off)
        g.reset_index()
        g.build()  # idempotent: the rebuilt index is identical
        again = g.fetch(0, g.stats()["n_entries"])
        for k in first:
            assert np.array_equal(first[k], again[k]), k
        bad = res.copy(); bad[10] = 0
    with dbi.GpuIndex(p) as g:
        g.add_proteins(bad, off)
        with pytest.raises(dbi.DbiError) as ei:
            g.build()
        assert ei.value.code == -3  # residue byte 0 rejected


def test_cfg1_full_size_parity():
    """BASELINE.json configs[0]: ~20k proteins / ~11M residues, trypsin, 2 MC, 600-6000 Da, no mods,
    index build + 10k queries, complete comparison with the oracle."""
    p = dbi.default_params()
    res, off = synth.config_proteome(1)
    g, o = both(p, res, off)
    try:
        check_all(g, o, nq=10_000, seed=20240601)
        # entry keys = distinct (int)(mass*10000)
        keys = g.entry_keys()
        exp_keys = np.unique((o.entries()["mass"] * 10000).astype(np.int32))
        assert np.array_equal(keys, exp_keys)
    finally:
        g.close()


def test_cfg2_mods_midsize_and_full_size_properties():
    """BASELINE.json configs[1] (static C, variable M + STY, <= 3 per peptide): complete comparison
    on 2000 proteins; at the full 20k proteins size-independent properties plus an order-independent
    checksum of every entry against the oracle."""
    p = dbi.default_params(**PARAM_SETS["cfg2_mods"])
    res, off = synth.config_proteome(2, 2000)
    g, o = both(p, res, off)
    try:
        check_all(g, o, nq=2000)
    finally:
        g.close()
    res, off = synth.config_proteome(2)
    g, o = both(p, res, off, keep=False)
    try:
        st, oc = g.stats(), o.counts()
        assert (st["n_emitted"], st["n_unique"], st["n_entries"]) == (oc["n_emitted"], oc["n_unique"], oc["n_entries"])
        n = st["n_entries"]
        chunk = 8_000_000
        def checksum(e):
            h = bits(e["mass"]) * np.uint64(0x9E3779B97F4A7C15)
            h ^= (e["first_prot"].astype(np.uint64) << np.uint64(32) | e["first_off"].astype(np.uint64)) * np.uint64(0xC2B2AE3D27D4EB4F)
            h ^= (e["len"].astype(np.uint64) << np.uint64(32) | e["modpat"].astype(np.uint64)) * np.uint64(0x165667B19E3779F9)
            h *= np.uint64(0xFF51AFD7ED558CCD)
            return int(np.bitwise_xor.reduce(h)), int(h.sum(dtype=np.uint64))
        last = 0.0
        for s in range(0, n, chunk):
            c = min(chunk, n - s)
            got = g.fetch(s, c, with_ids=False)
            exp = o.entries(s, c)
            assert np.array_equal(bits(got["mass"]), bits(exp["mass"]))   # same masses at the same index
            assert got["mass"][0] >= last and np.all(np.diff(got["mass"]) >= 0)  # sortedness
            last = got["mass"][-1]
        # order-independent checksum over all entries (ties may be ordered differently)
        gx = gs = ox = os_ = 0
        for s in range(0, n, chunk):
            c = min(chunk, n - s)
            a, b = checksum(g.fetch(s, c, with_ids=False))
            gx ^= a; gs = (gs + b) % (1 << 64)
            a, b = checksum(o.entries(s, c))
            ox ^= a; os_ = (os_ + b) % (1 << 64)
        assert (gx, gs) == (ox, os_)
        m_all = o.entries(0, n)["mass"] if n < 100_000_000 else None
        _, _, lo, hi = synth.synth_queries(m_all, 10_000, 5, da_fraction=0.5)
        b, c = g.query(lo, hi)
        ob, oc_, contig = o.query(lo, hi)
        assert contig and np.array_equal(b, ob) and np.array_equal(c, oc_)
    finally:
        g.close()


def test_cfg3_semi_sample():
    """BASELINE.json configs[2] (semi-tryptic window enumeration) on a 4000-protein sample."""
    p = dbi.default_params(semi=1)
    res, off = synth.config_proteome(3, 4000)
    g, o = both(p, res, off)
    try:
        check_all(g, o, nq=2000)
        assert g.stats()["n_emitted"] > 4 * len(res)  # the window explosion is real
    finally:
        g.close()


def _visible_gpus():
    import torch
    return torch.cuda.device_count()


def _dup_proteome(n_prot, seed=4711):
    """A proteome with duplicated proteins far apart (different shards): the merge must stay global."""
    res, off = synth.synth_proteome(n_prot, seed, median_len=150 if n_prot <= 100 else 400, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    seqs = seqs + [seqs[0], seqs[1]] + seqs[20:23] + [np.frombuffer(b"G" * 70 + b"MSTYMWSTK" + b"G" * 12 + b"MR", np.uint8)]
    res = np.concatenate(seqs)
    off = np.zeros(len(seqs) + 1, np.uint64)
    np.cumsum([len(x) for x in seqs], out=off[1:])
    return res, off


def check_sharded(handles, o, split_mass):
    """The rank slices, in rank order, ARE the oracle's index: masses position by position, first
    occurrences and protein lists global, slices cut at the splitter masses, routed queries and
    materialised hits summing to the global answer."""
    from dbindex_b200.multigpu import owned_mask, route_queries
    exp = o.entries()
    world = len(handles)
    counts = [g.stats()["n_entries"] for g in handles]
    assert sum(counts) == len(exp["mass"]), (counts, len(exp["mass"]))
    plo = exp["prot_list_off"].astype(np.int64)
    for r, g in enumerate(handles):
        n = counts[r]
        got = g.fetch(0, n)
        assert not np.any(got["first_prot"] == 0xFFFFFFFF), "a base peptide of another rank was not resolved"
        # the oracle's entries whose mass falls into one of this rank's slices, in mass order
        idx = np.nonzero(owned_mask(exp["mass"], split_mass, r, world))[0]
        assert len(idx) == n, (r, len(idx), n)
        sl = {k: exp[k][idx] for k in ("mass", "first_prot", "first_off", "len", "modpat")}
        sizes = (plo[idx + 1] - plo[idx]) if n else np.zeros(0, np.int64)
        sl["prot_list_off"] = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
        sl["prot_ids"] = (np.concatenate([exp["prot_ids"][plo[i]:plo[i + 1]] for i in idx]) if n
                          else np.zeros(0, np.uint32))
        assert_entries_equal(got, sl)
    _, _, lo, hi = synth.synth_queries(exp["mass"], 300, 3, da_fraction=0.4)
    ob, oc, _ = o.query(lo, hi)
    total = np.zeros(len(lo), np.int64)
    exp_hits = hits_canonical(o.query_hits(lo, hi))
    got_hits = [[] for _ in lo]
    for r, g in enumerate(handles):
        sel = route_queries(lo, hi, split_mass, r, world)
        b, c = g.query(lo, hi)
        part = np.zeros(len(lo), np.int64)
        part[sel] = c[sel]
        assert np.array_equal(c.astype(np.int64), part), "routing dropped hits"
        total += part
        if len(sel):
            hc = hits_canonical(g.query_hits(lo[sel], hi[sel]))
            for k, q in enumerate(sel):
                got_hits[q].extend(hc[k])
    assert np.array_equal(total, oc.astype(np.int64)), "routed hit counts differ from the global answer"
    assert [sorted(x) for x in got_hits] == exp_hits, "materialised hits differ from the oracle's"


@pytest.mark.parametrize("name,world", [("cfg1_tryptic", 2), ("cfg2_mods", 2), ("cfg2_mods", 3), ("cfg3_semi", 4),
                                        ("many_classes", 2), ("three_classes_k2", 3), ("mandatory_K_no_h2o", 2)])
def test_sharded_build_in_one_process_matches_oracle(name, world):
    """SURVEY.md 8e / invariant 11 through dbi_mg_build_local: `world` ranks = `world` handles of THIS
    process (all on cuda:0 here, so the 1-GPU test box exercises every kernel of the sharded path:
    shard packing + peer pulls, the two multisplit-scatter exchanges into mapped windows, group
    expansion from travelling site masks, hits resolved through the owners' windows)."""
    from dbindex_b200.multigpu import build_local, shard_proteins, split_masses
    p = dbi.default_params(**PARAM_SETS[name])
    res, off = _dup_proteome(600)
    handles = []
    try:
        for r in range(world):
            g = dbi.GpuIndex(p)
            sres, soff, _ = shard_proteins(res, off, r, world)
            g.add_proteins(sres, soff)
            handles.append(g)
        build_local(handles)
        o = Oracle(p, threads=NCPU)
        o.add_proteins(res, off)
        assert o.build() == 0
        check_sharded(handles, o, split_masses(handles[0], world))
        # idempotent: reset every rank and build again
        for g in handles:
            g.reset_index()
        build_local(handles)
        assert sum(g.stats()["n_entries"] for g in handles) == o.counts()["n_entries"]
    finally:
        for g in handles:
            g.close()


@pytest.mark.parametrize("name", ["cfg1_tryptic", "cfg2_mods", "cfg3_semi", "many_classes"])
def test_multi_gpu_sharded_build_matches_oracle(name, tmp_path):
    """The same through one process per GPU (torchrun, NCCL for the small collectives, CUDA IPC mappings
    of the windows).  Runs with every visible GPU (needs >= 2)."""
    import subprocess
    import sys
    n = _visible_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "r.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(29800 + len(name)), os.path.join(root, "tests", "dist_worker.py"),
           name, str(out), "gpu", "1500"]
    p = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=900, env=dict(os.environ))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
