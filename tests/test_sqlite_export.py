"""the exporter writes the reference's on-disk index (bucket directory, table, merged rows).
Checked against tests/sqlite_ref.py -- the storage-faithful restatement of DBIndexStoreSQLiteByte /
...IndexMerge fed with the same addSequence calls -- row by row, and read back through the reference's own
SQL + row walk.  Entries come from the oracle here (same shape as GpuIndex.fetch); the GPU parity suite shows
the GPU's entries equal them."""
import os
import sqlite3
import struct

import numpy as np

import dbindex_b200 as dbi
from dbindex_b200 import synth
from dbindex_b200.sqlite_export import SqliteIndexWriter, export_sqlite, index_dir_for, read_rows
from oracle.oracle_py import Oracle

from .sqlite_ref import SqliteStore, TABLE
from .util import PARAM_SETS, pack


def _build(params, n=50, seed=99):
    res, off = synth.synth_proteome(n, seed, median_len=220, min_len=5)
    seqs = [res[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(len(off) - 1)]
    seqs += [seqs[0], seqs[3], "AAGGLLKGGAALLKAAGGLLK", "GALKAGLKLAGKGLAK", seqs[0]]  # duplicates and isomers
    o = Oracle(params)
    o.add_proteins(*pack(seqs))
    assert o.build() == 0
    return o, seqs


def test_exported_rows_equal_the_reference_store(tmp_path):
    p = dbi.default_params()
    o, seqs = _build(p)
    n = o.counts()["n_entries"]
    info = export_sqlite(lambda b, c: o.entries(b, c), n, str(tmp_path / "proteome.fasta_abc"), index_factor=8,
                         mass_group_factor=p.mass_group_factor, chunk=97)  # chunk boundaries inside rows
    assert info["dir"] == index_dir_for(str(tmp_path / "proteome.fasta_abc")) and info["dir"].endswith(".idx")
    assert sorted(os.listdir(info["dir"])) == sorted(f"{i}.idx" for i in range(8))
    assert info["peptides"] == n and info["dropped_over_max_mass"] == 0
    # the same addSequence calls through the restatement of the reference's store
    em = o.emitted()
    store = SqliteStore(seqs, p.mass_group_factor)
    for m, pr, of, ln in zip(em["mass"], em["prot"], em["off"], em["len"]):
        store.add_sequence(float(m), int(of), int(ln), int(pr))
    store.stop_add_seq()
    ref_rows = {}
    for key, data in store.con.execute(f"SELECT precursor_mass_key, data FROM {TABLE};"):
        out = []
        store._parse_add(bytes(data), out, 0.0, float("inf"))
        ref_rows[key] = sorted((struct.pack("<d", m), s, ids) for m, s, ids in out)
    got_rows = {}
    for key, peps in read_rows(info["dir"]):
        assert key not in got_rows, "a row key appears twice"
        masses = [m for m, _, _, _ in peps]
        assert masses == sorted(masses), "rows are mass-sorted (IndexedSeqMerged.compareTo)"
        assert all(int(m * p.mass_group_factor) == key for m in masses)
        got_rows[key] = sorted((struct.pack("<d", m), seqs[ids[0]][off:off + ln], ids) for m, off, ln, ids in peps)
    assert got_rows == ref_rows
    # every row sits in the bucket Mult.getBucketForMass names, and the descending index exists
    for i in range(8):
        con = sqlite3.connect(os.path.join(info["dir"], f"{i}.idx"))
        keys = [k for (k,) in con.execute(f"SELECT precursor_mass_key FROM {TABLE};")]
        assert all(int(k / p.mass_group_factor) // 1000 == i for k in keys)
        assert ("precursor_mass_key_index_dsc",) in con.execute("SELECT name FROM sqlite_master WHERE type = 'index';").fetchall()
        con.close()


def test_exporter_answers_queries_like_the_oracle(tmp_path):
    """getSequences over the exported files (the reference's SQL + parseAddPeptideInfo row walk) == the oracle."""
    p = dbi.default_params(**PARAM_SETS["semi_nocut_mods"])
    p.n_mods, p.max_mods_per_peptide = 0, 0  # the reference's store holds unmodified peptides
    o, seqs = _build(p, n=30, seed=7)
    e = o.entries()
    info = export_sqlite(lambda b, c: o.entries(b, c), len(e["mass"]), str(tmp_path / "x"), index_factor=4,
                         mass_group_factor=p.mass_group_factor)
    reader = SqliteStore(seqs, p.mass_group_factor)   # only its query side is used, over the exported bucket files
    rng = np.random.default_rng(1)
    qm = np.concatenate([e["mass"][rng.integers(0, len(e["mass"]), 30)], rng.uniform(600, 6000, 10)])
    tol = np.concatenate([qm[:15] * 1e-5, np.zeros(15), np.full(10, 2.5)])
    lo, hi = np.maximum(qm - tol, 0.0), qm + tol
    b, c, contig = o.query(lo, hi)
    assert contig
    plo = e["prot_list_off"].astype(np.int64)
    for k in range(len(qm)):
        got = []
        for i in range(4):  # Mult.getSequences walks the buckets the range touches; walking all is a superset
            reader.con = sqlite3.connect(os.path.join(info["dir"], f"{i}.idx"))
            got += reader.get_sequences(float(qm[k]), float(tol[k]))
            reader.con.close()
        exp = [(float(e["mass"][i]), seqs[int(e["first_prot"][i])][int(e["first_off"][i]):int(e["first_off"][i]) + int(e["len"][i])],
                tuple(int(x) for x in e["prot_ids"][plo[i]:plo[i + 1]])) for i in range(int(b[k]), int(b[k] + c[k]))]
        canon = lambda xs: sorted((struct.pack("<d", m), s, ids) for m, s, ids in xs)  # noqa: E731
        assert canon(got) == canon(exp), (k, qm[k], tol[k])


def test_variants_are_skipped_and_heavy_peptides_dropped(tmp_path):
    w = SqliteIndexWriter(str(tmp_path / "v"), index_factor=8)
    w.add({"mass": np.array([700.0, 700.00001, 715.9949, 7999.5]), "first_off": np.array([0, 5, 0, 9]),
           "len": np.array([6, 6, 6, 60]), "modpat": np.array([0, 0, 3, 0]), "prot_list_off": np.array([0, 1, 3, 4, 5]),
           "prot_ids": np.array([0, 1, 1, 0, 2])})
    info = w.close()
    assert info["peptides"] == 3 and info["rows"] == 2 and info["dropped_over_max_mass"] == 0
    rows = dict(read_rows(info["dir"]))
    assert rows[7000000] == [(700.0, 0, 6, (0,)), (700.00001, 5, 6, (1, 1))]
    assert rows[79995000] == [(7999.5, 9, 60, (2,))]
