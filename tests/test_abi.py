"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dbindex_gpu.h declares, its structs match the ctypes mirror, and without a CUDA device the
product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

import dbindex_b200 as dbi
from dbindex_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dbindex_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dbi_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = dbi.load_library()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libdbindex_gpu.so does not export {s}"
    assert sorted(capi.ABI_SYMBOLS) == syms, "capi.ABI_SYMBOLS out of sync with the header"


def test_struct_layout_matches():
    lib = dbi.load_library()
    sp, ss = C.c_uint64(), C.c_uint64()
    lib.dbi_abi_sizes(C.byref(sp), C.byref(ss))
    assert sp.value == C.sizeof(capi.DbiParams)
    assert ss.value == C.sizeof(capi.DbiStats)


def test_default_params_contract():
    p = dbi.default_params()
    assert p.abi_version == 3 and p.min_len == 6 and p.mass_group_factor == 10000
    assert p.max_missed == 2 and p.semi == 0 and (p.min_mass, p.max_mass) == (600.0, 6000.0)
    assert [chr(i) for i in range(256) if p.is_enzyme[i]] == ["K", "R"]
    assert not any(p.is_nocut)
    assert abs(p.residue_mass[ord("G")] - 57.02146372) < 1e-9
    assert p.residue_mass[ord("L")] == p.residue_mass[ord("I")] == p.residue_mass[ord("X")]
    assert abs(p.h2o_proton - 19.01784115) < 1e-7
    assert p.residue_mass[ord("*")] == 0.0
    # static mods only when delta > 0 (AssignMassToStaticParam.java:9-10)
    q = dbi.default_params(static_mods={"C": 57.02146, "M": -1.0})
    assert abs(q.residue_mass[ord("C")] - (p.residue_mass[ord("C")] + 57.02146)) < 1e-12
    assert q.residue_mass[ord("M")] == p.residue_mass[ord("M")]
    # diff mods: one table entry per residue, zero shifts ignored (SearchParamReader.java:646)
    r = dbi.default_params(diff_mods=[("M", 15.9949), ("STY", 79.96633), ("X", 0.0)], max_mods_per_peptide=3)
    assert r.n_mods == 4 and [chr(r.mods[i].residue) for i in range(4)] == ["M", "S", "T", "Y"]
    avg = dbi.default_params(mono=False)
    assert abs(avg.residue_mass[ord("G")] - 57.0513) < 1e-9


def test_create_rejects_bad_params():
    lib = dbi.load_library()
    for kw in (dict(min_mass=-1.0), dict(max_mass=9000.0), dict(min_mass=700.0, max_mass=650.0), dict(min_len=0),
               dict(max_mods_per_peptide=5), dict(abi_version=99)):
        p = dbi.default_params(**kw)
        h = C.c_void_p()
        assert lib.dbi_create(C.byref(p), C.byref(h)) == -3  # DBI_EINVAL
        assert lib.dbi_last_error()


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly (DBI_ECUDA), never compute."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(dbi.DbiError) as ei:
        dbi.GpuIndex(dbi.default_params())
    assert ei.value.code == -5
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under dbindex_b200/ or include/ may reference it."""
    for d in ("dbindex_b200", "include"):
        for root, _, files in os.walk(os.path.join(ROOT, d)):
            if "build" in root.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(root, f), errors="replace").read()
                    assert "liboracle" not in txt and "oracle_py" not in txt and "orc_" not in txt, os.path.join(root, f)


def test_default_dbindex_params_factory_mirrors_the_reference():
    """DBIndexImpl.getDefaultDBIndexParams (DBIndexImpl.java:243-332) with the values of
    resources/dbindex.properties, including the reference's mass-type quirk (Boolean.valueOf("1"))."""
    from dbindex_b200.indexer import (DBIndexerException, getDefaultDBIndexParams,
                                      getDefaultDBIndexParamsForCrosslinkerAnalysis)
    sp = getDefaultDBIndexParams("/data/uniprot.fasta")
    p = sp.params
    assert (sp.dataBaseName, sp.indexFactor, sp.inMemoryIndex, sp.useIndex, sp.enzymeOffset) == \
        ("/data/uniprot.fasta", 8, True, True, 0)
    assert (p.max_missed, p.min_mass, p.max_mass, p.semi, p.min_len, p.mass_group_factor, p.add_h2o_proton) == \
        (6, 500.0, 6000.0, 0, 6, 10000, 1)
    assert [chr(i) for i in range(256) if p.is_enzyme[i]] == ["K", "R"] and not any(p.is_nocut)
    # average masses by default, as the reference computes it; monoisotopic on request
    assert not sp.useMonoParent and abs(p.residue_mass[ord("G")] - 57.0513) < 1e-9
    mono = getDefaultDBIndexParams("/data/uniprot.fasta", use_mono=True)
    assert mono.useMonoParent and abs(mono.params.residue_mass[ord("G")] - 57.02146372) < 1e-9
    assert getDefaultDBIndexParams("x.fasta", inMemoryIndex=False).inMemoryIndex is False
    # the registry key separates everything that shapes the index
    assert sp.key() != mono.key() and sp.key() == getDefaultDBIndexParams("/data/uniprot.fasta").key()
    assert sp.key() != getDefaultDBIndexParams("/data/uniprot.fasta", max_missed=2).key()
    # cross-linker set (DBIndexImpl.java:443-491): no H2O + proton, mandatory internal K
    xl = getDefaultDBIndexParamsForCrosslinkerAnalysis("x.fasta")
    assert xl.params.add_h2o_proton == 0 and xl.params.has_mandatory == 1 and xl.mandatoryInternalAAs == "K"
    assert [chr(i) for i in range(256) if xl.params.is_mandatory[i]] == ["K"]
    assert xl.key() != getDefaultDBIndexParams("x.fasta").key()
    assert DBIndexerException is not None
