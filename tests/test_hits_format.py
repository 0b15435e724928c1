"""Host side of the run-grouped query answer (dbi_hit_buffers): expand_hits turns runs back into one record
per hit -- the shape parseAddPeptideInfo produces (DBIndexStoreSQLiteByteIndexMerge.java:386-481)."""
import numpy as np

from dbindex_b200.capi import HIT_FIELDS, DbiHitBuffers, DbiHitCounts, expand_hits


def test_expand_hits_by_hand():
    # 2 queries; query 0: run A (3 hits), run B (1 hit); query 1: run C (2 hits, empty protein list is impossible
    # in a real index but the CSR arithmetic must hold anyway)
    raw = {
        "hit_off": np.array([0, 4, 6], np.uint64), "pep_off": np.array([0, 2, 3], np.uint64),
        "modpat": np.array([0, 1, 2, 0, 5, 6], np.uint32), "pep_hit_off": np.array([0, 3, 4, 6], np.uint64),
        "mass": np.array([1000.5, 1000.5000001, 2000.25]), "first_prot": np.array([7, 8, 9], np.uint32),
        "first_off": np.array([10, 20, 30], np.uint32), "len": np.array([3, 2, 4], np.uint16),
        "flanks": np.frombuffer(b"--KAAAR--GGGKKKLLL", np.uint8).copy(),
        "seq_off": np.array([0, 3, 5, 9], np.uint64), "seq": np.frombuffer(b"PEPTKIDER", np.uint8).copy(),
        "prot_list_off": np.array([0, 2, 3, 3], np.uint64), "prot_ids": np.array([7, 70, 8], np.uint32),
    }
    h = expand_hits(raw)
    assert h["hit_pep"].tolist() == [0, 0, 0, 1, 2, 2]
    assert h["mass"].tolist() == [1000.5] * 3 + [1000.5000001] + [2000.25] * 2
    assert h["first_prot"].tolist() == [7, 7, 7, 8, 9, 9] and h["first_off"].tolist() == [10, 10, 10, 20, 30, 30]
    assert h["len"].tolist() == [3, 3, 3, 2, 4, 4] and h["modpat"].tolist() == [0, 1, 2, 0, 5, 6]
    so = h["seq_off"].astype(np.int64)
    assert [h["seq"][so[i]:so[i + 1]].tobytes() for i in range(6)] == [b"PEP"] * 3 + [b"TK"] + [b"IDER"] * 2
    po = h["prot_list_off"].astype(np.int64)
    assert [h["prot_ids"][po[i]:po[i + 1]].tolist() for i in range(6)] == [[7, 70]] * 3 + [[8]] + [[]] * 2
    assert h["flanks"].reshape(-1, 6)[3].tobytes() == b"R--GGG" and h["flanks"].reshape(-1, 6)[5].tobytes() == b"KKKLLL"


def test_expand_hits_empty_and_partial():
    raw = {"hit_off": np.zeros(3, np.uint64), "pep_hit_off": np.zeros(1, np.uint64), "modpat": np.zeros(0, np.uint32),
           "mass": np.zeros(0), "seq_off": np.zeros(1, np.uint64), "seq": np.zeros(0, np.uint8)}
    h = expand_hits(raw)
    assert len(h["mass"]) == 0 and h["seq_off"].tolist() == [0] and "prot_ids" not in h


def test_hit_buffers_struct_matches_field_table():
    assert [n for n, _ in DbiHitBuffers._fields_] == list(HIT_FIELDS)
    assert [n for n, _ in DbiHitCounts._fields_] == ["nq", "n_hits", "n_peps", "n_seq_bytes", "n_prot_ids"]
