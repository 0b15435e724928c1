"""A second, independently written restatement of the digestion as a SET formulation
(no loop-with-breaks): used to cross-check the oracle, in the spirit of the reference's
MassRangeFilteringIndex (MassRangeFilteringIndex.java:90-108), the brute-force store that
re-digests instead of indexing."""
from __future__ import annotations

from itertools import combinations


def seq_mass(p, s: str) -> float:
    m = 0.0
    if p.add_h2o_proton:
        m += p.h2o_proton
    m += p.cterm
    m += p.nterm
    for ch in s:
        m += p.residue_mass[ord(ch)]
    return m


def digest_set(p, seqs):
    """All (prot, start, len, mass) the reference emits, in (prot, start, end) order.  Because
    residue masses are >= 0 both the mass and the missed-cleavage count are monotone in `end`, so
    the reference's break conditions are equivalent to plain per-window predicates."""
    out = []
    filt = chr(p.filter_aa) if p.filter_aa > 0 else None
    for pid, s in enumerate(seqs):
        L = len(s)
        enz = [bool(p.is_enzyme[ord(c)]) for c in s]
        nocut = [bool(p.is_nocut[ord(c)]) for c in s]
        mand = [bool(p.has_mandatory and p.is_mandatory[ord(c)]) for c in s]
        for start in range(L):
            n_ok = start == 0 or (enz[start - 1] and not nocut[start])
            dead = False  # a qualifying window without any mandatory residue ended this start (DBIndexer.java:334-344)
            for end in range(start, L):
                # PeptideFilterByMaxOccurrencies: the occurrence count is monotone in `end`, so the
                # reference's break is a per-window predicate
                if filt is not None and s[start:end + 1].count(filt) > p.filter_max:
                    continue
                c_ok = end == L - 1 or (enz[end] and not nocut[end + 1])
                if not ((n_ok or c_ok) if p.semi else (n_ok and c_ok)):
                    continue
                if sum(enz[start:end + 1]) - 1 > p.max_missed:
                    continue
                ln = end - start + 1
                if ln < p.min_len:
                    continue
                m = seq_mass(p, s[start:end + 1])
                if m < p.min_mass or m > p.max_mass:
                    continue
                if p.has_mandatory:
                    if dead:
                        continue
                    if not any(mand[start:end + 1]):
                        dead = True
                        continue
                    if not any(mand[start:end]):  # only as the last residue: SKIP (DBIndexStoreSQLiteMult.java:245-263)
                        continue
                out.append((pid, start, ln, m))
    return out


def expand_set(p, pep: str, base_mass: float):
    """Mod variants of one peptide per the SPEC: [(mass, positions tuple)], k ascending,
    lexicographic."""
    diff = {}
    for i in range(p.n_mods):
        diff[chr(p.mods[i].residue)] = p.mods[i].delta
    K = p.max_mods_per_peptide if p.n_mods > 0 else 0
    sites = [i for i, c in enumerate(pep) if c in diff]
    out = [(base_mass, ())]
    for k in range(1, min(K, len(sites)) + 1):
        for combo in combinations(sites, k):
            m = base_mass
            for pos in combo:
                m = m + diff[pep[pos]]
            if p.min_mass <= m <= p.max_mass:
                out.append((m, combo))
    return out


def read_fasta_py(path: str):
    """Pure-Python restatement of the FASTA record grammar (what csrc/fasta.cpp must reproduce):
    '>' at the start of a line opens a record, the defline is the rest of that line; the sequence is
    the following non-empty lines, stripped at both ends, concatenated, upper-cased; anything before
    the first '>' is ignored.  Lines end with '\\n' (an optional '\\r' before it is dropped)."""
    deflines, seqs, cur = [], [], []
    with open(path, "rb") as f:
        data = f.read().decode("latin-1")
    for line in data.split("\n"):
        if line.startswith(">"):
            if deflines:
                seqs.append("".join(cur))
            deflines.append(line[1:].rstrip("\r\n"))
            cur = []
        else:
            t = line.strip(" \t\r\n\v\f")
            if t:
                cur.append("".join(chr(ord(c) - 32) if "a" <= c <= "z" else c for c in t))
    if deflines:
        seqs.append("".join(cur))
    return deflines, seqs
