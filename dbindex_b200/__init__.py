"""dbindex_b200 -- B200-native (CUDA sm_100a) implementation of the dbIndex hot path.

FASTA residues -> in-silico digestion -> differential-mod expansion -> mass-sorted,
de-duplicated peptide index -> batched precursor-mass lookup, behind the reference's
indexer / store API.  The compute lives in ``libdbindex_gpu.so`` (C ABI declared in
``include/dbindex_gpu.h``); this package is the thin host-side binding.  There is no
CPU fallback: without the built library or without a CUDA device every entry point
raises.
"""
from .capi import (  # noqa: F401
    DbiParams,
    DbiError,
    GpuIndex,
    default_params,
    load_library,
    STAGE_NAMES,
)
from .indexer import (  # noqa: F401
    DBIndexer,
    DBIndexImpl,
    DBIndexSearchParams,
    IndexedSequence,
    IndexedProtein,
    MassRange,
    getDefaultDBIndexParams,
)
