"""Export a built index into the reference's on-disk format: the directory
`<name>.idx/` of DBIndexStoreSQLiteMult with one SQLite file `<bucket>.idx` per mass bucket
(DBIndexStoreSQLiteMult.java:101-142), each holding the table of DBIndexStoreSQLiteByte
(`blazmass_sequences(precursor_mass_key INTEGER PRIMARY KEY, data BINARY)`, :586-587; index
`precursor_mass_key_index_dsc`, :607) with the MERGED rows DBIndexStoreSQLiteByteIndexMerge writes
(getMergedData, Merge:696-716): per row key `(int)(mass * massGroupFactor)` (Byte:187) the peptides of
the row in mass order, each as little-endian `double mass, int offset, int length, int proteinId...`
closed by the separator 0x7FFFFFFF (Merge:28).  An unmodified reference opens such a directory as an
existing index (Mult.indexExists, :72-88) and answers getSequences from it.

Host-side format code only: it reads entries in the shape GpuIndex.fetch returns (entries in mass
order: mass, first_off, len, modpat, prot_list_off, prot_ids) and never touches the device.  Variants
with a mod pattern are skipped: the reference's store holds unmodified peptides only (it parses
differential mods but never expands them, SearchParamReader.java:631-687).
"""
from __future__ import annotations

import os
import sqlite3
import struct
from typing import Callable, Dict, Iterable, Optional

import numpy as np

SEQ_SEPARATOR_INT = 2 ** 31 - 1   # DBIndexStoreSQLiteByteIndexMerge.SEQ_SEPARATOR_INT (Merge:28)
MAX_MASS = 8000                   # Constants.MAX_PRECURSOR_MASS as the store uses it (Mult:55-56)
IDX_SUFFIX = ".idx"               # DBIndexStoreSQLiteMult.IDX_SUFFIX (:32)
TABLE = "blazmass_sequences"


def _java_int(x: float) -> int:
    """(int) of a non-negative double: truncation, saturating like Java."""
    return int(min(max(x, -2147483648.0), 2147483647.0))


def index_dir_for(database_id: str) -> str:
    """Directory DBIndexStoreSQLiteMult.init derives from its database id (Mult:101-108)."""
    base = os.path.abspath(database_id)
    name = os.path.basename(base)
    if not name.endswith(IDX_SUFFIX):
        name += IDX_SUFFIX
    return os.path.join(os.path.dirname(base), name)


class SqliteIndexWriter:
    """Streams entries (in mass order, chunk by chunk) into the bucket files."""

    def __init__(self, database_id: str, index_factor: int = 8, mass_group_factor: int = 10000):
        self.dir = index_dir_for(database_id)
        os.makedirs(self.dir, exist_ok=True)
        self.n_buckets = int(index_factor)
        self.bucket_range = MAX_MASS // self.n_buckets          # Constants.BUCKET_MASS_RANGE (Mult:56)
        self.factor = mass_group_factor
        self.cons = []
        for i in range(self.n_buckets):
            path = os.path.join(self.dir, f"{i}{IDX_SUFFIX}")
            if os.path.exists(path):
                os.remove(path)
            con = sqlite3.connect(path)
            con.execute(f"CREATE TABLE IF NOT EXISTS {TABLE} (precursor_mass_key INTEGER PRIMARY KEY, data BINARY);")
            self.cons.append(con)
        self._row_key: Optional[int] = None
        self._row = bytearray()
        self.n_peptides = 0
        self.n_rows = 0
        self.n_dropped = 0

    def _flush(self):
        if self._row_key is None or not self._row:
            return
        mass_of_row = self._row_key / self.factor
        bucket = int(mass_of_row) // self.bucket_range          # getBucketForMass (Mult:215-217)
        self.cons[bucket].execute(f"INSERT INTO {TABLE} (precursor_mass_key, data) VALUES (?, ?);",
                                  (self._row_key, bytes(self._row)))
        self.n_rows += 1
        self._row = bytearray()

    def add(self, entries: Dict[str, np.ndarray]):
        """Entries in ascending mass order (continuing the previous call)."""
        mass = np.asarray(entries["mass"], dtype=np.float64)
        off = np.asarray(entries["first_off"], dtype=np.int64)
        ln = np.asarray(entries["len"], dtype=np.int64)
        pat = np.asarray(entries["modpat"]) if "modpat" in entries and entries["modpat"] is not None else None
        plo = np.asarray(entries["prot_list_off"], dtype=np.int64)
        ids = np.asarray(entries["prot_ids"], dtype=np.int64)
        for i in range(len(mass)):
            if pat is not None and pat[i] != 0:
                continue  # a differential-mod variant: no counterpart in the reference's store
            m = float(mass[i])
            if int(m) // self.bucket_range > self.n_buckets - 1:   # "unsupported precursor mass" (Mult:282-287)
                self.n_dropped += 1
                continue
            key = _java_int(m * self.factor)
            if key != self._row_key:
                self._flush()
                self._row_key = key
            self._row += struct.pack("<dii", m, int(off[i]), int(ln[i]))
            self._row += struct.pack(f"<{plo[i + 1] - plo[i]}i", *ids[plo[i]:plo[i + 1]].tolist())
            self._row += struct.pack("<i", SEQ_SEPARATOR_INT)
            self.n_peptides += 1

    def close(self) -> dict:
        self._flush()
        for con in self.cons:
            con.execute(f"CREATE INDEX IF NOT EXISTS precursor_mass_key_index_dsc ON {TABLE} (precursor_mass_key DESC);")
            con.commit()
            con.close()
        return {"dir": self.dir, "buckets": self.n_buckets, "rows": self.n_rows, "peptides": self.n_peptides,
                "dropped_over_max_mass": self.n_dropped}


def export_sqlite(fetch: Callable[[int, int], Dict[str, np.ndarray]], n_entries: int, database_id: str,
                  index_factor: int = 8, mass_group_factor: int = 10000, chunk: int = 1 << 20) -> dict:
    """Write entries [0, n_entries) -- `fetch(begin, count)` returns them in the shape of GpuIndex.fetch -- as a
    reference index directory.  Returns {"dir", "buckets", "rows", "peptides", ...}."""
    w = SqliteIndexWriter(database_id, index_factor, mass_group_factor)
    for b in range(0, int(n_entries), chunk):
        w.add(fetch(b, min(chunk, int(n_entries) - b)))
    return w.close()


def read_rows(index_dir: str) -> Iterable[tuple]:
    """(key, [(mass, offset, length, (protein ids...)), ...]) of every row, buckets in order -- the reader side of
    parseAddPeptideInfo (Merge:386-481), for checks."""
    names = sorted((f for f in os.listdir(index_dir) if f.endswith(IDX_SUFFIX)), key=lambda f: int(f[:-len(IDX_SUFFIX)]))
    for name in names:
        con = sqlite3.connect(os.path.join(index_dir, name))
        for key, data in con.execute(f"SELECT precursor_mass_key, data FROM {TABLE} ORDER BY precursor_mass_key;"):
            data = bytes(data)
            i, peps = 0, []
            while i < len(data):
                m, off, ln = struct.unpack_from("<dii", data, i)
                i += 16
                prots = []
                while True:
                    (p,) = struct.unpack_from("<i", data, i)
                    i += 4
                    if p == SEQ_SEPARATOR_INT:
                        break
                    prots.append(p)
                peps.append((m, off, ln, tuple(prots)))
            yield key, peps
        con.close()
