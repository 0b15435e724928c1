"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package (dbindex_b200/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [vp]
    lib.orc_destroy.argtypes = [vp]
    lib.orc_add_proteins.argtypes = [vp, vp, vp, C.c_uint32]
    lib.orc_set_threads.argtypes = [C.c_int]
    lib.orc_max_threads.restype = C.c_int
    lib.orc_build.argtypes = [vp]
    lib.orc_build.restype = C.c_int
    lib.orc_build_from_records.argtypes = [vp, vp, vp, vp, vp, C.c_uint64]
    lib.orc_build_from_records.restype = C.c_int
    lib.orc_counts.argtypes = [vp, u64p, u64p, u64p, u64p]
    lib.orc_emitted.argtypes = [vp, vp, vp, vp, vp]
    lib.orc_entries.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, vp, vp]
    lib.orc_query.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, C.POINTER(C.c_int)]
    lib.orc_query_ppm.argtypes = [vp, C.c_double, C.c_double, vp, C.c_uint64, u64p]
    lib.orc_query_ppm.restype = C.c_uint64
    lib.orc_query_hits.argtypes = [vp, vp, vp, C.c_uint64] + [vp] * 11 + [u64p, u64p, u64p]
    lib.orc_query_hits.restype = C.c_int
    lib.orc_calculate_mass.argtypes = [vp, C.c_char_p, C.c_uint64]
    lib.orc_calculate_mass.restype = C.c_double
    lib.orc_tolerance_in_dalton.argtypes = [C.c_double, C.c_double]
    lib.orc_tolerance_in_dalton.restype = C.c_double
    lib.orc_merge_intervals.argtypes = [vp, vp, C.c_uint64, vp, vp]
    lib.orc_merge_intervals.restype = C.c_uint64
    lib.orc_get_residues.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_char_p, C.c_char_p]
    _lib = lib
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(mono: bool = True, **overrides):
    """dbi_params for the oracle WITHOUT touching the product library: liboracle.so carries its own
    copy of the host-only parameter helpers (csrc/params.cpp); only the struct definition (pure Python)
    comes from dbindex_b200.capi.  bench.py --impl reference uses this."""
    from dbindex_b200.capi import DbiParams, default_params as _dp
    lib = load()
    pp = C.POINTER(DbiParams)
    lib.dbi_default_params.restype = None
    lib.dbi_default_params.argtypes = [pp, C.c_int]
    lib.dbi_params_add_static_mod.restype = None
    lib.dbi_params_add_static_mod.argtypes = [pp, C.c_uint8, C.c_double]
    lib.dbi_params_set_enzyme.restype = None
    lib.dbi_params_set_enzyme.argtypes = [pp, C.c_char_p, C.c_char_p]
    lib.dbi_params_add_diff_mod.restype = C.c_int
    lib.dbi_params_add_diff_mod.argtypes = [pp, C.c_char_p, C.c_double]
    return _dp(mono, _lib=lib, **overrides)


class Oracle:
    """CPU restatement of the reference path (see dbindex_oracle.cpp)."""

    def __init__(self, params, threads: int = 1):
        self.lib = load()
        self.lib.orc_set_threads(threads)
        self.threads = threads
        self._params = params  # keep alive
        self.h = C.c_void_p(self.lib.orc_create(C.byref(params)))

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_proteins(self, residues, offsets):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.lib.orc_add_proteins(self.h, _p(residues), _p(offsets), len(offsets) - 1)

    def build(self) -> int:
        self.lib.orc_set_threads(self.threads)
        return self.lib.orc_build(self.h)

    def build_from_records(self, mass, prot, off, length) -> int:
        mass = np.ascontiguousarray(mass, np.float64)
        prot = np.ascontiguousarray(prot, np.uint32)
        off = np.ascontiguousarray(off, np.uint32)
        length = np.ascontiguousarray(length, np.uint16)
        return self.lib.orc_build_from_records(self.h, _p(mass), _p(prot), _p(off), _p(length), len(mass))

    def counts(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.orc_counts(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return {"n_emitted": a.value, "n_unique": b.value, "n_entries": c.value, "n_prot_ids": d.value}

    def emitted(self):
        n = self.counts()["n_emitted"]
        out = {"mass": np.empty(n, np.float64), "prot": np.empty(n, np.uint32), "off": np.empty(n, np.uint32),
               "len": np.empty(n, np.uint16)}
        self.lib.orc_emitted(self.h, _p(out["mass"]), _p(out["prot"]), _p(out["off"]), _p(out["len"]))
        return out

    def entries(self, begin=0, count=None):
        c = self.counts()
        if count is None:
            count = c["n_entries"] - begin
        out = {
            "mass": np.empty(count, np.float64), "first_prot": np.empty(count, np.uint32),
            "first_off": np.empty(count, np.uint32), "len": np.empty(count, np.uint16),
            "modpat": np.empty(count, np.uint32), "prot_list_off": np.zeros(count + 1, np.uint64),
        }
        # size the id list
        self.lib.orc_entries(self.h, begin, count, None, None, None, None, None, _p(out["prot_list_off"]), None)
        ids = np.empty(int(out["prot_list_off"][count]), np.uint32)
        self.lib.orc_entries(self.h, begin, count, _p(out["mass"]), _p(out["first_prot"]), _p(out["first_off"]),
                             _p(out["len"]), _p(out["modpat"]), _p(out["prot_list_off"]), _p(ids))
        out["prot_ids"] = ids
        return out

    def query(self, lo, hi):
        lo = np.ascontiguousarray(lo, np.float64)
        hi = np.ascontiguousarray(hi, np.float64)
        nq = len(lo)
        b = np.empty(nq, np.uint64)
        c = np.empty(nq, np.uint64)
        contig = C.c_int(1)
        self.lib.orc_set_threads(self.threads)
        self.lib.orc_query(self.h, _p(lo), _p(hi), nq, _p(b), _p(c), C.byref(contig))
        return b, c, bool(contig.value)

    def query_hits(self, lo, hi, flanks: bool = True) -> dict:
        """Every hit of every [lo, hi] materialised like parseAddPeptideInfo (same shape as dbi_query_hits)."""
        lo = np.ascontiguousarray(lo, np.float64)
        hi = np.ascontiguousarray(hi, np.float64)
        nq = len(lo)
        hit_off = np.zeros(nq + 1, np.uint64)
        nh, ns, ni = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.orc_set_threads(self.threads)
        rc = self.lib.orc_query_hits(self.h, _p(lo), _p(hi), nq, _p(hit_off), *([None] * 10), C.byref(nh), C.byref(ns),
                                     C.byref(ni))
        assert rc == 0
        H = nh.value
        out = {"hit_off": hit_off, "mass": np.empty(H, np.float64), "first_prot": np.empty(H, np.uint32),
               "first_off": np.empty(H, np.uint32), "len": np.empty(H, np.uint16), "modpat": np.empty(H, np.uint32),
               "seq_off": np.zeros(H + 1, np.uint64), "seq": np.empty(ns.value, np.uint8),
               "flanks": np.empty(6 * H, np.uint8) if flanks else None,
               "prot_list_off": np.zeros(H + 1, np.uint64), "prot_ids": np.empty(ni.value, np.uint32)}
        rc = self.lib.orc_query_hits(self.h, _p(lo), _p(hi), nq, _p(hit_off), _p(out["mass"]), _p(out["first_prot"]),
                                     _p(out["first_off"]), _p(out["len"]), _p(out["modpat"]), _p(out["seq_off"]),
                                     _p(out["seq"]), _p(out["flanks"]), _p(out["prot_list_off"]), _p(out["prot_ids"]),
                                     C.byref(nh), C.byref(ns), C.byref(ni))
        assert rc == 0
        return out

    def query_ppm(self, precursor_mass: float, ppm: float):
        """getSequencesUsingPPMTolerance restated: entry indices in result-list order, number of probes."""
        probes = C.c_uint64()
        n = self.lib.orc_query_ppm(self.h, precursor_mass, ppm, None, 0, C.byref(probes))
        idx = np.empty(n, np.uint64)
        self.lib.orc_query_ppm(self.h, precursor_mass, ppm, _p(idx), n, C.byref(probes))
        return idx, probes.value

    def calculate_mass(self, seq: bytes) -> float:
        return self.lib.orc_calculate_mass(self.h, seq, len(seq))


def tolerance_in_dalton(mass: float, ppm: float) -> float:
    return load().orc_tolerance_in_dalton(mass, ppm)


def merge_intervals(mass, tol):
    mass = np.ascontiguousarray(mass, np.float64)
    tol = np.ascontiguousarray(tol, np.float64)
    lo = np.empty(len(mass), np.float64)
    hi = np.empty(len(mass), np.float64)
    n = load().orc_merge_intervals(_p(mass), _p(tol), len(mass), _p(lo), _p(hi))
    return lo[:n].copy(), hi[:n].copy()


def get_residues(prot_seq: bytes, off: int, length: int):
    l = C.create_string_buffer(3)
    r = C.create_string_buffer(3)
    load().orc_get_residues(prot_seq, len(prot_seq), off, length, l, r)
    return l.raw.decode(), r.raw.decode()
