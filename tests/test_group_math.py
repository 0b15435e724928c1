"""Executable specification of the arithmetic behind the group path (csrc/mods_grp.cu), checked
against brute force on the CPU:

  * class sequences in heap numbering (children of v are v*C + c + 1) and their masses;
  * K5g/K6g: the number of variants of a group = occurrences of its class sequence as a subsequence
    of the peptide's site-class string, by one DP pass over the sites (deeper sequences first);
  * K6x: inside a group the entries are ordered by product-set BLOCKS (k = 2: first site; k = 3:
    middle site, entries = sites below x sites above, last site fastest; k = 4: second site), so
    that rank -> occurrence is one table lookup, one division and bit selects -- and that mapping
    must be a bijection onto the occurrences.

The kernels are exercised on the GPU by tests/test_gpu_parity.py; this file pins the maths they
implement so that a change of the formulas fails here first, without a GPU."""
from itertools import combinations, product

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def n_seq(C, K):
    return sum(C ** k for k in range(K + 1))


def seq_of(v, C):
    """class sequence of heap node v, first chosen site first (pack_seq of mods_common.cuh)"""
    out = []
    while v > 0:
        out.append((v - 1) % C)
        v = (v - 1) // C
    return out[::-1]


def dp_counts(classes, C, K):
    """SeqCounts::site of mods_grp.cu: cnt[v] for every class sequence after all sites"""
    S = n_seq(C, K)
    n_par = (S - 1) // C
    cnt = [0] * S
    cnt[0] = 1
    for c in classes:
        for p in range(n_par - 1, -1, -1):
            if cnt[p]:
                cnt[p * C + c + 1] += cnt[p]
    return cnt


def brute_counts(classes, C, K):
    S = n_seq(C, K)
    index = {tuple(seq_of(v, C)): v for v in range(S)}
    cnt = [0] * S
    for k in range(K + 1):
        for sites in combinations(range(len(classes)), k):
            cnt[index[tuple(classes[i] for i in sites)]] += 1
    return cnt


@pytest.mark.parametrize("C,K", [(1, 4), (2, 3), (2, 4), (3, 2), (4, 2), (5, 2), (16, 1)])
def test_heap_numbering_covers_every_sequence_once(C, K):
    S = n_seq(C, K)
    assert S <= 32 or (C, K) == (16, 1) and S == 17
    seqs = [tuple(seq_of(v, C)) for v in range(S)]
    assert len(set(seqs)) == S
    assert set(seqs) == {t for k in range(K + 1) for t in product(range(C), repeat=k)}
    for v in range(1, S):  # parent = sequence without its last class
        assert seqs[(v - 1) // C] == seqs[v][:-1] and seqs[v][-1] == (v - 1) % C


@settings(max_examples=200, deadline=None)
@given(st.sampled_from([(1, 4), (2, 3), (3, 2), (5, 2)]), st.data())
def test_dp_counts_equal_brute_force(ck, data):
    C, K = ck
    classes = data.draw(st.lists(st.integers(0, C - 1), max_size=9))
    assert dp_counts(classes, C, K) == brute_counts(classes, C, K)


def test_total_variants_is_the_sum_of_binomials():
    from math import comb
    for n in range(0, 12):
        classes = [i % 2 for i in range(n)]
        assert sum(dp_counts(classes, 2, 3)) == sum(comb(n, k) for k in range(4))


# ---- block ordering of the expansion (block_sites / block_size / block_entry of mods_grp.cu) ----
def sites_of(mask):
    return [i for i in range(64) if mask >> i & 1]


def occurrences(masks):
    """all site tuples i0 < i1 < ... with i_j in masks[j] (lexicographic)"""
    out = [()]
    for m in masks:
        out = [t + (i,) for t in out for i in sites_of(m) if not t or i > t[-1]]
    return out


def block_table(masks):
    """[(site p, exclusive prefix of the block sizes)] -- what K6x stages per group"""
    k = len(masks)
    tab, acc = [], 0
    for p in sites_of(masks[0] if k == 2 else masks[1]):
        tab.append((p, acc))
        if k == 2:
            size = sum(1 for i in sites_of(masks[1]) if i > p)
        else:
            nl = sum(1 for i in sites_of(masks[0]) if i < p)
            if k == 3:
                size = nl * sum(1 for i in sites_of(masks[2]) if i > p)
            else:
                size = nl * sum(1 for i2 in sites_of(masks[2]) if i2 > p for i3 in sites_of(masks[3]) if i3 > i2)
        acc += size
    return tab, acc


def unrank(masks, r):
    """entry r of the group -> site tuple, as the kernel computes it"""
    k = len(masks)
    if k == 1:
        return (sites_of(masks[0])[r],)
    tab, total = block_table(masks)
    assert r < total
    lo = max(t for t in range(len(tab)) if tab[t][1] <= r)  # last block whose first entry is <= r
    p, q = tab[lo][0], r - tab[lo][1]
    if k == 2:
        return (p, [i for i in sites_of(masks[1]) if i > p][q])
    L = [i for i in sites_of(masks[0]) if i < p]
    if k == 3:
        R = [i for i in sites_of(masks[2]) if i > p]
        return (L[q // len(R)], p, R[q % len(R)])
    pairs = [(i2, i3) for i2 in sites_of(masks[2]) if i2 > p for i3 in sites_of(masks[3]) if i3 > i2]
    return (L[q // len(pairs)], p) + pairs[q % len(pairs)]


# up to 12 sites, spread over both halves of the 64-bit mask
mask64 = st.integers(0, (1 << 12) - 1).map(lambda x: (x & 0x3f) << 3 | (x >> 6) << 52)


@settings(max_examples=150, deadline=None)
@given(st.integers(1, 4), st.data())
def test_block_order_is_a_bijection_onto_the_occurrences(k, data):
    # disjoint or equal class masks, as real peptides have (a residue has one class)
    base = [data.draw(mask64) for _ in range(2)]
    base[1] &= ~base[0]
    masks = [base[data.draw(st.integers(0, 1))] for _ in range(k)]
    occ = occurrences(masks)
    if k == 1:
        assert [unrank(masks, r) for r in range(len(occ))] == occ
        return
    _, total = block_table(masks)
    assert total == len(occ)
    got = [unrank(masks, r) for r in range(total)]
    assert sorted(got) == occ and len(set(got)) == total
    if k == 2:  # first-site blocks keep the lexicographic order
        assert got == occ
