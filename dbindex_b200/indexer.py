"""Host-side mirror of the reference's indexer / search API over the C ABI.

Same names, argument meaning and error behaviour as the Java classes, so that a
user of ``DBIndexImpl`` / ``DBIndexer`` finds the calls they know:

* ``DBIndexer``   -- DBIndexer.java (init / run / getSequencesUsing*Tolerance /
  getSequences(ranges) / getProteins)
* ``DBIndexImpl`` -- DBIndexImpl.java, the ``DBIndexInterface`` facade
  (getSequences(mass, tol), getSequences(ranges), getProteins, getIndexedProteinById,
  getProteinSequenceById)

Only string assembly happens here (peptide substrings, flanking residues); all
digestion, sorting, merging and searching runs in libdbindex_gpu.so.  The Java shim
under ``java/`` is the same logic in the reference's own language.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Set, Tuple

import numpy as np

from .capi import DbiError, DbiParams, GpuIndex, parse_fasta

MAX_INDEX_RESIDUE_LEN = 3      # Constants.java:44
MAX_PRECURSOR_MASS = 8000      # Constants.java:20
PRECISION = 1e-6               # Constants.java:50
ONE_MILLION = 1000000.0


class DBIndexStoreException(Exception):
    """edu.scripps.yates.utilities.fasta.dbindex.DBIndexStoreException"""


class DBIndexerException(Exception):
    """DBIndexerException.java:9-18"""


@dataclass
class MassRange:
    precMass: float
    tolerance: float


@dataclass(frozen=True)
class IndexedProtein:
    accession: str
    id: int


@dataclass
class IndexedSequence:
    """What parseAddPeptideInfo builds per hit (DBIndexStoreSQLiteByteIndexMerge.java:452-462):
    IndexedSequence(0, mass, sequence, "", "") + setProteinIds + setResidues.  ``sequenceOffset``
    and ``modPositions`` are the superset north_star asks for."""

    mass: float
    sequence: str
    proteinIds: List[int]
    resLeft: str = ""
    resRight: str = ""
    sequenceOffset: int = -1
    sequenceLen: int = 0
    modPositions: Tuple[int, ...] = ()
    id: int = 0

    def key(self):
        return (np.float64(self.mass).view(np.uint64).item(), self.sequence, self.modPositions)


def get_residues(seq_offset: int, seq_len: int, protein_sequence: str) -> Tuple[str, str]:
    """Util.getResidues (Util.java:130-162), right flank off-by-one kept."""
    prot_len = len(protein_sequence)
    left_i = seq_offset - MAX_INDEX_RESIDUE_LEN if seq_offset >= MAX_INDEX_RESIDUE_LEN else 0
    left_len = min(MAX_INDEX_RESIDUE_LEN, seq_offset)
    left = protein_sequence[left_i:left_i + left_len]
    end = seq_offset + seq_len
    right_len = min(MAX_INDEX_RESIDUE_LEN, prot_len - end - 1)
    right = protein_sequence[end:end + right_len] if (end < prot_len and right_len > 0) else ""
    return left.rjust(MAX_INDEX_RESIDUE_LEN, "-"), right.ljust(MAX_INDEX_RESIDUE_LEN, "-")


def tolerance_in_dalton(actual_mass: float, ppm: float) -> float:
    """IndexUtil.getToleranceInDalton (util/IndexUtil.java:238-240)."""
    return actual_mass * (1 - 1 / (ppm / ONE_MILLION + 1))


def merge_intervals(ranges: Sequence[MassRange]) -> List[Tuple[float, float]]:
    """Interval.massRangeToInterval (Interval.java:27-38) + MergeIntervals.mergeIntervals
    (MergeIntervals.java:16-46)."""
    iv = []
    for r in ranges:
        lo = r.precMass - r.tolerance
        if lo < 0.0:
            lo = 0.0
        iv.append((lo, r.precMass + r.tolerance))
    if len(iv) < 2:
        return iv
    iv.sort(key=lambda t: t[0])  # stable, by start
    out = []
    start, end = iv[0]
    for s, e in iv[1:]:
        if end >= s:
            end = max(end, e)
        else:
            out.append((start, end))
            start, end = s, e
    out.append((start, end))
    return out


def read_fasta(path: str) -> Tuple[List[str], np.ndarray, np.ndarray]:
    """FASTA -> (deflines, residues, offsets) through the native multi-threaded parser
    (capi.parse_fasta; the reference uses the external FastaReader, DBIndexer.java:560-571)."""
    return parse_fasta(path)


# resources/dbindex.properties:3-26, restated (the file itself is configuration of the reference)
DBINDEX_PROPERTIES = {
    "default_index_type": "INDEX_NORMAL", "default_in_memory_index": True, "default_index_factor": 8,
    "default_max_internal_cleavages": 6, "default_max_precursor_mass": 6000.0, "default_min_precursor_mass": 500.0,
    "default_use_index": True, "default_enzyme_nocut_residues": "", "default_enzyme_residues": "KR",
    "default_enzyme_offset": 0, "default_mass_type_parent": "1", "mass_group_factor": 10000,
    "add_h2o_plus_proton": True, "mandatory_internal_AAs": "K", "default_semicleavage": False,
}


@dataclass
class DBIndexSearchParams:
    """io/DBIndexSearchParamsImpl.java:46-75 -- the programmatic parameter object: the kernel-facing
    part is `params` (dbi_params), the rest is what the facade needs."""
    params: DbiParams
    dataBaseName: str
    indexFactor: int = 8
    inMemoryIndex: bool = True
    useIndex: bool = True
    indexType: str = "INDEX_NORMAL"
    enzymeOffset: int = 0          # never reaches the Enzyme (SearchParams.java:303); only named the index file
    useMonoParent: bool = False
    mandatoryInternalAAs: Optional[str] = None
    discardDecoyRegexp: Optional[str] = None
    staticParams: str = ""         # SearchParams.getStaticParams(): text of the static mods, part of the index name

    def key(self) -> str:
        """What getByParam keys its registry on (DBIndexImpl.java:44-49: the full index file name, i.e.
        the database plus every parameter that shapes the index, IndexUtil.java:276-323)."""
        p = self.params
        enz = "".join(chr(i) for i in range(256) if p.is_enzyme[i])
        nocut = "".join(chr(i) for i in range(256) if p.is_nocut[i])
        mods = ",".join(f"{chr(p.mods[i].residue)}{p.mods[i].delta!r}" for i in range(p.n_mods))
        statics = ",".join(f"{i}:{p.residue_mass[i]!r}" for i in range(256) if p.residue_mass[i])
        return "|".join(str(x) for x in (
            self.dataBaseName, self.indexFactor, p.max_missed, p.min_mass, p.max_mass, enz, nocut, self.enzymeOffset,
            self.useMonoParent, p.add_h2o_proton, p.mass_group_factor, p.semi, p.min_len, self.mandatoryInternalAAs,
            self.discardDecoyRegexp, p.max_mods_per_peptide, mods, statics))


def getDefaultDBIndexParams(fastaFilePath, inMemoryIndex: Optional[bool] = None, use_mono: Optional[bool] = None,
                            **overrides) -> DBIndexSearchParams:
    """DBIndexImpl.getDefaultDBIndexParams(File | String [, boolean inMemoryIndex]) (DBIndexImpl.java:243-332):
    the defaults of dbindex.properties -- KR, no no-cut residues, 6 missed cleavages, 500-6000 Da, full
    specificity, H2O + proton added, factor 10000, index factor 8.

    Mass type: the reference reads `default_mass_type_parent=1` with `Boolean.valueOf("1")`, which is
    false, so this factory builds an AVERAGE-mass index (DBIndexImpl.java:281-282; only the proteoform
    factory compares with "1", :407-409).  use_mono=None reproduces that; pass True for monoisotopic."""
    from .capi import default_params
    pr = DBINDEX_PROPERTIES
    mono = bool(use_mono) if use_mono is not None else False
    kw = dict(enzyme=pr["default_enzyme_residues"], nocut=pr["default_enzyme_nocut_residues"],
              max_missed=pr["default_max_internal_cleavages"], min_mass=pr["default_min_precursor_mass"],
              max_mass=pr["default_max_precursor_mass"], mass_group_factor=pr["mass_group_factor"],
              add_h2o_proton=1 if pr["add_h2o_plus_proton"] else 0, semi=1 if pr["default_semicleavage"] else 0)
    kw.update(overrides)
    return DBIndexSearchParams(
        params=default_params(mono=mono, **kw), dataBaseName=os.fspath(fastaFilePath),
        indexFactor=pr["default_index_factor"],
        inMemoryIndex=pr["default_in_memory_index"] if inMemoryIndex is None else bool(inMemoryIndex),
        useIndex=pr["default_use_index"], indexType=pr["default_index_type"],
        enzymeOffset=pr["default_enzyme_offset"], useMonoParent=mono)


def getDefaultDBIndexParamsForCrosslinkerAnalysis(fastaFilePath, inMemoryIndex: Optional[bool] = None,
                                                  use_mono: Optional[bool] = None, **overrides) -> DBIndexSearchParams:
    """DBIndexImpl.getDefaultDBIndexParamsForCrosslinkerAnalysis (DBIndexImpl.java:342-372,443-491): the
    dbindex.properties defaults, but H2O + proton is NOT added to the peptide masses (:473) and every indexed
    peptide needs one of `mandatory_internal_AAs` (= "K") as an internal residue (:477-478; semantics
    DBIndexer.java:334-344 and DBIndexStoreSQLiteMult.java:245-263, applied by the digestion kernels)."""
    kw = dict(add_h2o_proton=0, mandatory_internal=DBINDEX_PROPERTIES["mandatory_internal_AAs"])
    kw.update(overrides)
    sp = getDefaultDBIndexParams(fastaFilePath, inMemoryIndex, use_mono, **kw)
    sp.mandatoryInternalAAs = kw["mandatory_internal"]
    return sp


def getDefaultDBIndexParamsForProteoformAnalysis(*args, **kwargs):
    """DBIndexImpl.java:392-432: needs UniProt annotations and proteoform FASTA expansion -- out of scope."""
    raise DBIndexerException("proteoform analysis needs the UniProt annotation service; out of scope of the GPU index")


def createFullIndexFileName(sparam: "DBIndexSearchParams") -> str:
    """IndexUtil.createFullIndexFileName (util/IndexUtil.java:242-324): <database>_<md5 of the parameters that
    shape the index>, the same fields in the same order and spelling.  Two things cannot be reproduced
    byte for byte and are documented deviations: the reference appends `char[].toString()` of
    mandatoryInternalAAs (a JVM identity hash, different on every run) -- the characters themselves are used
    here -- and it knows no differential mods: they are appended only when present, so indexes without them
    keep the reference's key."""
    import hashlib
    p = sparam.params

    def jbool(b):
        return "true" if b else "false"

    def jdouble(x):  # Double.toString for the magnitudes that occur here
        return repr(float(x))

    enz = "".join(chr(i) for i in range(256) if p.is_enzyme[i])
    nocut = "".join(chr(i) for i in range(256) if p.is_nocut[i])
    u = "enzymeOffset:%d" % sparam.enzymeOffset
    u += ", enzymeResidues:" + enz
    u += ", enzymeNoCutResidues:" + nocut
    u += ", maxCleavages:%d" % p.max_missed
    u += ", minPrecursorMass:" + jdouble(p.min_mass)
    u += ", maxPrecursorMass:" + jdouble(p.max_mass)
    u += ", static:" + sparam.staticParams
    u += ", semiCleave:" + jbool(p.semi)
    if p.filter_aa > 0:
        u += ", pepFilter:%s%d" % (chr(p.filter_aa), p.filter_max)  # PeptideFilterByMaxOccurrencies.toString
    u += ", isH2OPlusProtonAdded:" + jbool(p.add_h2o_proton)
    u += ", massGroupFactor:%d" % p.mass_group_factor
    u += ", massType:" + jbool(sparam.useMonoParent)
    if p.has_mandatory:
        u += ", mandatoryInternalAAs:" + "".join(chr(i) for i in range(256) if p.is_mandatory[i])
    u += ", proteoForms:false, maxVariationsPerPeptide:null, useUniprot:false, uniprotVersion:null"
    u += ", usePhosphoSite:false, phosphoSiteSpecies:null, sufix:null"
    u += ", discardDecoys:" + ("null" if sparam.discardDecoyRegexp is None else sparam.discardDecoyRegexp)
    if p.n_mods > 0 and p.max_mods_per_peptide > 0:
        u += ", diffMods:" + ",".join("%s%r" % (chr(p.mods[i].residue), p.mods[i].delta) for i in range(p.n_mods))
        u += ", maxDiffMods:%d" % p.max_mods_per_peptide
    return sparam.dataBaseName + "_" + hashlib.md5(u.encode("utf-8")).hexdigest()


INDEX_FILE_SUFFIX = ".gpuidx"  # the reference's directory is <name>.idx (DBIndexer.java:63,429-431)


class DBIndexer:
    """DBIndexer.java: the indexer / orchestrator, GPU store behind it."""

    def __init__(self, params: DbiParams, index_factor: int = 8, index_path: Optional[str] = None):
        """index_path: where the finished index lives on disk (None = in-memory index only).  run() skips
        indexing when the file exists ("Found existing index, skipping indexing", DBIndexer.java:522-531)
        and writes it after a fresh build."""
        self.sparam = params
        self.index_factor = index_factor  # dbindex.properties:6; only shapes the ">= MAX_MASS" query rule
        self.index_path = index_path
        self.loaded_from_disk = False
        self.inited = False
        self.index: Optional[GpuIndex] = None
        self.deflines: List[str] = []   # ProteinCache.defs
        self._seq_cache: dict = {}

    # -- lifecycle: init() (DBIndexer.java:412) -> run() (:508) --------------------------
    def init(self):
        if self.inited:
            raise DBIndexerException("Already inited")  # DBIndexer.java:413-415
        self.index = GpuIndex(self.sparam)
        self.inited = True

    def export_reference_index(self, database_id: str) -> dict:
        """Write the built index as the reference's own on-disk index -- directory `<database_id>.idx/` with one
        SQLite file per mass bucket, merged rows (dbindex_b200/sqlite_export.py; DBIndexStoreSQLiteMult.java:101-142,
        DBIndexStoreSQLiteByteIndexMerge.java:696-716) -- so that an unmodified reference finds an existing index
        there.  Differential-mod variants have no counterpart in that format and are left out."""
        from .sqlite_export import export_sqlite
        self._require()
        n = self.index.stats()["n_entries"]
        return export_sqlite(lambda b, c: self.index.fetch(b, c), n, database_id, index_factor=self.index_factor,
                             mass_group_factor=self.sparam.mass_group_factor)

    def _require(self):
        if not self.inited or self.index is None:
            raise DBIndexStoreException("Indexer is not initialized")  # DBIndexStoreSQLiteMult.java:153

    def add_proteins(self, deflines: Sequence[str], residues: np.ndarray, offsets: np.ndarray):
        """protCache.addProtein for a packed batch (DBIndexer.java:605)."""
        self._require()
        self.deflines.extend(d.replace("\t", " ") for d in deflines)  # ProteinCache.java:87-89
        self.index.add_proteins(residues, offsets)

    def run(self, fasta: Optional[str] = None, proteins: Optional[Tuple[Sequence[str], Sequence[str]]] = None):
        """DBIndexer.run(): stream the FASTA, cut every protein, close the store."""
        self._require()
        if self.index_path is not None and os.path.exists(self.index_path):
            # indexStore.indexExists() -> "Found existing index, skipping indexing" (DBIndexer.java:522-531)
            try:
                self.index.load(self.index_path)
            except DbiError as e:
                raise DBIndexerException(str(e)) from e
            self.loaded_from_disk = True
            # the deflines (ProteinCache.defs) are not part of the index file: re-read them when a FASTA is at hand
            if fasta is not None:
                self.deflines = [d.replace("\t", " ") for d in read_fasta(fasta)[0]]
            elif proteins is not None:
                self.deflines = [d.replace("\t", " ") for d in proteins[0]]
            return
        if fasta is not None:
            deflines, residues, offsets = read_fasta(fasta)
            self.add_proteins(deflines, residues, offsets)
        elif proteins is not None:
            deflines, seqs = proteins
            residues = np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8)
            offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
            np.cumsum([len(s) for s in seqs], out=offsets[1:])
            self.add_proteins(deflines, residues, offsets)
        try:
            self.index.build()  # cutSeq per protein + stopAddSeq (DBIndexer.java:616,666)
            if self.index_path is not None:
                self.index.save(self.index_path)
        except DbiError as e:
            raise DBIndexerException(str(e)) from e

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None
        self.inited = False

    # -- helpers ---------------------------------------------------------------------------
    def getProteinSequence(self, pid: int) -> str:
        s = self._seq_cache.get(pid)
        if s is None:
            s = self.index.get_protein(pid).decode("latin-1")
            if len(self._seq_cache) < 100000:
                self._seq_cache[pid] = s
        return s

    def _out_of_buckets(self, lo: float, hi: float) -> bool:
        """DBIndexStoreSQLiteMult.getBucketsForMassRange / the 'unsupported precursor mass'
        early return (Mult:55-56,215-217,333-338): empty answer when a bound reaches MAX_MASS."""
        bucket_range = MAX_PRECURSOR_MASS // self.index_factor
        nb = self.index_factor
        return int(lo) // bucket_range > nb - 1 or int(hi) // bucket_range > nb - 1

    def _materialise(self, lo: np.ndarray, hi: np.ndarray, keep: Optional[np.ndarray] = None) -> List[List[IndexedSequence]]:
        """getSequences for a batch of ranges in ONE device pass (dbi_query_hits): the objects
        parseAddPeptideInfo builds (DBIndexStoreSQLiteByteIndexMerge.java:452-462) -- sequence, flanks,
        mass and protein ids all come back from the GPU; only the Python objects are made here."""
        try:
            h = self.index.query_hits(lo, hi)
        except DbiError as e:
            raise DBIndexStoreException(str(e)) from e
        ho, so, po = (h[k].astype(np.int64) for k in ("hit_off", "seq_off", "prot_list_off"))
        seq, flanks = h["seq"].tobytes().decode("latin-1"), h["flanks"].tobytes().decode("latin-1")
        out = []
        for q in range(len(lo)):
            if keep is not None and not keep[q]:
                out.append([])
                continue
            lst = []
            for i in range(ho[q], ho[q + 1]):
                pat = int(h["modpat"][i])
                pos = tuple(((pat >> (8 * k)) & 0xFF) - 1 for k in range(4) if (pat >> (8 * k)) & 0xFF)
                lst.append(IndexedSequence(float(h["mass"][i]), seq[so[i]:so[i + 1]], h["prot_ids"][po[i]:po[i + 1]].tolist(),
                                           flanks[6 * i:6 * i + 3], flanks[6 * i + 3:6 * i + 6], int(h["first_off"][i]),
                                           int(h["len"][i]), pos))
            out.append(lst)
        return out

    # -- queries --------------------------------------------------------------------------
    def getSequencesBatch(self, masses: Sequence[float], tolerances: Sequence[float]) -> List[List[IndexedSequence]]:
        """Many getSequences(precMass, tol) in one device call."""
        self._require()
        m = np.asarray(masses, dtype=np.float64)
        t = np.asarray(tolerances, dtype=np.float64)
        lo = np.maximum(m - t, 0.0)  # Mult:324-329
        hi = m + t
        keep = np.array([not self._out_of_buckets(lo[i], hi[i]) for i in range(len(m))], dtype=bool)
        return self._materialise(lo, hi, keep)

    def getSequencesUsingDaltonTolerance(self, precursorMass: float, massToleranceInDa: float) -> List[IndexedSequence]:
        """DBIndexer.java:762-772 -> DBIndexStoreSQLiteMult.getSequences (Mult:315-350)."""
        return self.getSequencesBatch([precursorMass], [massToleranceInDa])[0]

    def getSequencesUsingPPMTolerance(self, precursorMass: float, massToleranceInPPM: float) -> List[IndexedSequence]:
        """DBIndexer.java:787-844: one Dalton query, then exact-mass probes above the upper
        bound until one comes back empty."""
        tol = tolerance_in_dalton(precursorMass, massToleranceInPPM)
        sequences = self.getSequencesUsingDaltonTolerance(precursorMass, tol)
        have = {s.key() for s in sequences}
        upper = precursorMass + tol
        while True:
            tol2 = tolerance_in_dalton(upper, massToleranceInPPM)
            if upper - tol2 < precursorMass:
                seq2 = self.getSequencesUsingDaltonTolerance(upper, 0.0)
                if not seq2:
                    break
                for s in seq2:
                    if s.key() not in have:
                        have.add(s.key())
                        sequences.append(s)
            else:
                break
            new_upper = upper + PRECISION
            if new_upper == upper:
                break
            upper = new_upper
        return sequences

    def getSequences(self, ranges: Sequence[MassRange]) -> List[IndexedSequence]:
        """DBIndexer.java:855-871 -> Mult.getSequences(List<MassRange>) (Mult:353-430)."""
        self._require()
        if len(ranges) == 1:
            return self.getSequencesUsingDaltonTolerance(ranges[0].precMass, ranges[0].tolerance)
        merged = merge_intervals(ranges)
        for lo, hi in merged:  # "Cannot query, unsupported precursor mass" -> empty list (Mult:383-387)
            if self._out_of_buckets(lo, hi):
                return []
        if not merged:
            return []
        lo = np.array([m[0] for m in merged])
        hi = np.array([m[1] for m in merged])
        out: List[IndexedSequence] = []
        for lst in self._materialise(lo, hi):
            out.extend(lst)
        return out

    def getProteins(self, seq) -> Set[IndexedProtein] | List[IndexedProtein]:
        """getProteins(String) (DBIndexer.java:925-947) / getProteins(IndexedSequence) (:882)."""
        self._require()
        if isinstance(seq, IndexedSequence):
            return [IndexedProtein(self.deflines[i] if i < len(self.deflines) else "", i) for i in seq.proteinIds]
        mass = self.index.calculate_mass(seq.encode("latin-1"))  # IndexUtil.calculateMass
        ret: Set[IndexedProtein] = set()
        for s in self.getSequencesUsingDaltonTolerance(mass, 0.0):
            if s.sequence == seq and not s.modPositions:
                ret.update(self.getProteins(s))
        return ret

    def getNumberSequences(self) -> int:
        self._require()
        return int(self.index.stats()["n_entries"])

    def getParentMasses(self) -> List[float]:
        """DBIndexer.java:984-992: entry keys divided back by massGroupFactor."""
        self._require()
        f = float(self.sparam.mass_group_factor)
        return [k / f for k in self.index.entry_keys().tolist()]


class DBIndexImpl:
    """DBIndexImpl.java: the DBIndexInterface facade a search engine holds."""

    _by_param_key: dict = {}  # DBIndexImpl.java:35 dbIndexByParamKey

    def __init__(self, params, fasta: Optional[str] = None,
                 proteins: Optional[Tuple[Sequence[str], Sequence[str]]] = None, index_factor: int = 8):
        """params: dbi_params, or a DBIndexSearchParams (then the FASTA is its dataBaseName, as in
        DBIndexImpl(DBIndexSearchParams), DBIndexImpl.java:117-145)."""
        self.sparam = params if isinstance(params, DBIndexSearchParams) else None
        if self.sparam is not None:
            if fasta is None and proteins is None:
                fasta = self.sparam.dataBaseName
            index_factor = self.sparam.indexFactor
            params = self.sparam.params
        index_path = None
        if self.sparam is not None and not self.sparam.inMemoryIndex:
            # an on-disk index, named like the reference's (DBIndexer.java:429 createFullIndexFileName)
            index_path = createFullIndexFileName(self.sparam) + INDEX_FILE_SUFFIX
        self.indexer = DBIndexer(params, index_factor, index_path=index_path)
        self.indexer.init()
        self.indexer.run(fasta=fasta, proteins=proteins)
        self._proteins_by_seq: dict = {}  # DBIndexImpl.java:33
        if self.sparam is not None:
            DBIndexImpl._by_param_key[self.sparam.key()] = self  # DBIndexImpl.java:137-139

    @staticmethod
    def getByParam(sParam: "DBIndexSearchParams") -> "DBIndexImpl":
        """DBIndexImpl.getByParam (DBIndexImpl.java:44-49): one index per parameter key."""
        hit = DBIndexImpl._by_param_key.get(sParam.key())
        return hit if hit is not None else DBIndexImpl(sParam)

    def getSequences(self, *args) -> List[IndexedSequence]:
        """getSequences(double precursorMass, double massTolerance) (DBIndexImpl.java:180) or
        getSequences(List<MassRange>) (:194)."""
        if len(args) == 2:
            return self.indexer.getSequencesUsingDaltonTolerance(float(args[0]), float(args[1]))
        return self.indexer.getSequences(list(args[0]))

    def getProteins(self, seq):
        if isinstance(seq, str):  # DBIndexImpl.java:222-237, memoised
            if seq not in self._proteins_by_seq:
                self._proteins_by_seq[seq] = self.indexer.getProteins(seq)
            return self._proteins_by_seq[seq]
        return self.indexer.getProteins(seq)  # DBIndexImpl.java:208

    def getIndexedProteinById(self, pid: int) -> IndexedProtein:
        return IndexedProtein(self.indexer.deflines[pid], pid)  # DBIndexImpl.java:501-504

    def getProteinSequenceById(self, pid: int) -> str:
        return self.indexer.getProteinSequence(pid)  # DBIndexImpl.java:511-513

    def close(self):
        if self.sparam is not None:
            DBIndexImpl._by_param_key.pop(self.sparam.key(), None)
        self.indexer.close()
