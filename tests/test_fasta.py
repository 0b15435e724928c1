"""The native FASTA ingest (csrc/fasta.cpp, SURVEY.md 8f rank 2: the step before the hot path)
against a pure-Python restatement of the record grammar.  Host-only code: no GPU needed."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import dbindex_b200 as dbi
from dbindex_b200 import synth
from dbindex_b200.capi import DbiError, parse_fasta

from .pyref import read_fasta_py


def check(path, threads=0):
    deflines, residues, offsets = parse_fasta(str(path), threads)
    exp_def, exp_seq = read_fasta_py(str(path))
    assert deflines == exp_def
    assert len(offsets) == len(exp_seq) + 1 and offsets[0] == 0
    got = [residues[int(offsets[i]):int(offsets[i + 1])].tobytes().decode("latin-1") for i in range(len(exp_seq))]
    assert got == exp_seq
    assert int(offsets[-1]) == len(residues)
    return deflines, residues, offsets


EDGE_CASES = {
    "plain": ">sp|P1|A first\nMKWVTFISLL\nLLFSSAYSRG\n>sp|P2|B second\nACDEFGHIK\n",
    "no_trailing_newline": ">a\nMKK\n>b\nPEPTIDEK",
    "crlf": ">a desc\r\nMKWV\r\nTFIS\r\n>b\r\nAAAK\r\n",
    "blank_and_space_lines": "\n\n>a\n\nMK WV\n   \n\tTFIS  \n\n>b\n\n",
    "junk_before_first_record": "this is not fasta\nMKKKK\n>a\nPEPK\n",
    "lowercase_and_stop": ">a\nmkwvtf*\n>b\nAcDeFgHiK\n",
    "empty_records": ">a\n>b\n>c\nK\n>d\n",
    "gt_inside_line": ">a x>y\nMK>K\nAAA\n>b\nKK\n",
    "only_text": "MKWVTFISLL\nACDEFG\n",
    "empty_file": "",
    "single_gt": ">",
    "defline_only_spaces": ">   \nAAAK\n",
}


@pytest.mark.parametrize("name", list(EDGE_CASES))
@pytest.mark.parametrize("threads", [1, 3, 0])
def test_fasta_edge_cases(tmp_path, name, threads):
    p = tmp_path / f"{name}.fasta"
    p.write_bytes(EDGE_CASES[name].encode("latin-1"))
    check(p, threads)


def test_fasta_round_trip_of_the_synthetic_proteome(tmp_path):
    """synth.write_fasta -> native parser gives back the packed proteome the generator made, for
    every thread count (records straddle the chunk boundaries of the parallel scan)."""
    res, off = synth.synth_proteome(3000, 99, median_len=200, min_len=1)
    p = tmp_path / "syn.fasta"
    synth.write_fasta(str(p), res, off, width=60)
    for threads in (1, 2, 7, 16, 0):
        deflines, r2, o2 = parse_fasta(str(p), threads)
        assert np.array_equal(r2, res) and np.array_equal(o2, off)
        assert deflines[:2] == ["sp|S0000000|SYN_0", "sp|S0000001|SYN_1"] and len(deflines) == 3000
    check(p)


def test_fasta_missing_file_is_an_error(tmp_path):
    with pytest.raises(DbiError):
        parse_fasta(str(tmp_path / "nope.fasta"))


line = st.text(alphabet=st.sampled_from(list("ACDEFGHIKLMNPQRSTVWYacdk >*\t\r")), max_size=30)


@settings(max_examples=150, deadline=None)
@given(st.lists(line, max_size=25), st.sampled_from([1, 2, 5]))
def test_fasta_random_texts(tmp_path_factory, lines, threads):
    p = tmp_path_factory.mktemp("fa") / "r.fasta"
    p.write_bytes("\n".join(lines).encode("latin-1"))
    check(p, threads)
