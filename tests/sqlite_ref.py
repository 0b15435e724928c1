"""TEST INFRASTRUCTURE: a storage-faithful restatement of the reference's default store on stdlib
sqlite3 -- the same table, the same 20-byte little-endian records, the same SQL -- used to check
the oracle's row / merge / query logic against an independent formulation that goes through a real
SQLite B-tree (SURVEY.md 8d "fidelity variant").  It is not the reference (Java, un-buildable here)
and does not pin parity; it narrows what an error in the oracle could hide behind.

Follows, line by line:
  DBIndexStoreSQLiteByte.createTables        (DBIndexStoreSQLiteByte.java:586-587)   schema
  DBIndexStoreSQLiteByte.updateCachedData    (:185-226)   key = (int)(mass * massGroupFactor), record
  DBIndexStoreSQLiteByte.insertSequence      (:235-264)   INSERT / UPDATE data = (data || ?)
  DBIndexStoreSQLiteByteIndexMerge.mergePeptides / getMergedData   (Merge:85-127, 620-719)
  DBIndexStoreSQLiteByteIndexMerge.getSequences / parseAddPeptideInfo  (Merge:146-217, 386-481)
"""
from __future__ import annotations

import sqlite3
import struct
from typing import List, Tuple

SEQ_SEPARATOR_INT = 2 ** 31 - 1          # Merge:28
CHUNK_SIZE = 1000 * 20                    # Constants.COMMIT_SEQUENCES * BYTE_PER_SEQUENCE (Byte:36)
TABLE = "blazmass_sequences"


class SqliteStore:
    def __init__(self, proteins: List[str], mass_group_factor: int = 10000):
        self.proteins = proteins               # ProteinCache: id = insertion index
        self.factor = mass_group_factor
        self.con = sqlite3.connect(":memory:")
        self.con.execute(f"CREATE TABLE IF NOT EXISTS {TABLE} (precursor_mass_key INTEGER PRIMARY KEY, data BINARY);")
        self.cache = {}                        # dataMap: rowId -> DynByteBuffer

    # -- build ------------------------------------------------------------------------------
    def add_sequence(self, mass: float, offset: int, length: int, protein_id: int):
        row = int(mass * self.factor)          # (int) truncates toward zero (Byte:187)
        buf = self.cache.setdefault(row, bytearray())
        buf += struct.pack("<diii", mass, offset, length, protein_id)   # Byte:202-212, little-endian
        if len(buf) > CHUNK_SIZE:              # Byte:216-224
            self._insert(row, bytes(buf))
            buf.clear()

    def _insert(self, key: int, data: bytes):  # Byte:235-264
        cur = self.con.execute(f"SELECT 1 FROM {TABLE} WHERE precursor_mass_key = ?;", (key,))
        if cur.fetchone():
            self.con.execute(f"UPDATE {TABLE} SET data = (data || ?) WHERE precursor_mass_key = ?;", (data, key))
        else:
            self.con.execute(f"INSERT INTO {TABLE} (precursor_mass_key, data) VALUES (?, ?);", (key, data))

    def stop_add_seq(self):                    # commitCachedData + createIndex (Abstract:292-299, Merge:64-83)
        for key, buf in self.cache.items():
            if buf:
                self._insert(key, bytes(buf))
        self.cache.clear()
        rows = self.con.execute(f"SELECT precursor_mass_key, data FROM {TABLE};").fetchall()
        for key, data in rows:                 # mergePeptides (Merge:85-127)
            self.con.execute(f"UPDATE {TABLE} SET data = ? WHERE precursor_mass_key = ?;",
                             (self._merged(bytes(data)), key))
        self.con.execute(f"CREATE INDEX IF NOT EXISTS precursor_mass_key_index_dsc ON {TABLE} (precursor_mass_key DESC);")
        self.con.commit()

    def _merged(self, data: bytes) -> bytes:   # getMergedData (Merge:620-719)
        groups = {}                            # peptide string -> [(mass, offset, length, protein)]
        for i in range(0, len(data), 20):
            mass, off, ln, prot = struct.unpack_from("<diii", data, i)
            pep = self.proteins[prot][off:off + ln]           # ProteinCache.getPeptideSequence
            groups.setdefault(pep, []).append((mass, off, ln, prot))
        merged = [(seqs[0][0], seqs[0][1], seqs[0][2], [s[3] for s in seqs]) for seqs in groups.values()]
        merged.sort(key=lambda m: m[0])        # IndexedSeqMerged.compareTo: by mass only
        out = bytearray()
        for mass, off, ln, prots in merged:
            out += struct.pack("<dii", mass, off, ln)
            for p in prots:
                out += struct.pack("<i", p)
            out += struct.pack("<i", SEQ_SEPARATOR_INT)
        return bytes(out)

    # -- query ------------------------------------------------------------------------------
    def get_sequences(self, prec_mass: float, tolerance: float) -> List[Tuple[float, str, Tuple[int, ...]]]:
        min_f = max(prec_mass - tolerance, 0.0)               # Merge:155-159
        max_f = prec_mass + tolerance
        min_key = max(int(min_f * self.factor), 0)            # Merge:170-174
        max_key = int(max_f * self.factor)                    # Merge:178
        ret = []
        cur = self.con.execute(f"SELECT precursor_mass_key, data FROM {TABLE} WHERE precursor_mass_key BETWEEN ? AND ?;",
                               (min_key, max_key))
        for _, data in cur:
            self._parse_add(bytes(data), ret, min_f, max_f)
        return ret

    def _parse_add(self, data: bytes, ret, min_mass: float, max_mass: float):   # Merge:386-481
        n = len(data)
        if n % 4 != 0:
            return
        i = 0
        while i < n:
            mass, off, ln, prot = struct.unpack_from("<diii", data, i)
            i += 20
            if mass > max_mass:                # the row is mass-sorted
                break
            if mass < min_mass:                # skip to the separator
                while True:
                    (p,) = struct.unpack_from("<i", data, i)
                    i += 4
                    if p == SEQ_SEPARATOR_INT:
                        break
                continue
            prots = [prot]
            while True:
                (p,) = struct.unpack_from("<i", data, i)
                i += 4
                if p == SEQ_SEPARATOR_INT:
                    break
                prots.append(p)
            ret.append((mass, self.proteins[prot][off:off + ln], tuple(prots)))

    def all_entries(self):
        out = []
        for _, data in self.con.execute(f"SELECT precursor_mass_key, data FROM {TABLE} ORDER BY precursor_mass_key;"):
            self._parse_add(bytes(data), out, 0.0, float("inf"))
        return out
