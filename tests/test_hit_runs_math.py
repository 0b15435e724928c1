"""Executable specification of the run numbering of dbi_query_hits (csrc/query.cu: hits_count_runs_kernel,
hits_runs_kernel): the hits of a query are cut into segments of 256, one warp each; a hit is the head of a run
when it is the first of its query or differs from the previous ENTRY in peptide or mass; a segment's runs are
numbered from the exclusive scan of the per-segment head counts plus a ballot prefix inside every 32 hits.
Checked against a plain group-by."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

SEG = 256


def runs_by_segments(base, mass_bits, q_begin, q_count):
    """Python model of the kernels: returns (pep_off[nq + 1], pep_hit_off[n_runs + 1], run_entry[n_runs])."""
    nq = len(q_begin)
    hit_off = np.concatenate(([0], np.cumsum(q_count))).astype(np.int64)
    nseg = (np.asarray(q_count) + SEG - 1) // SEG
    seg_off = np.concatenate(([0], np.cumsum(nseg))).astype(np.int64)
    seg_q = np.repeat(np.arange(nq), nseg)

    def head(e, i):  # run_head()
        return i == 0 or base[e] != base[e - 1] or mass_bits[e] != mass_bits[e - 1]

    def segment(s):  # hit_segment()
        q = seg_q[s]
        i0 = (s - seg_off[q]) * SEG
        return q_begin[q] + i0, hit_off[q] + i0, i0, min(SEG, q_count[q] - i0)

    nruns = np.zeros(len(seg_q), np.int64)
    for s in range(len(seg_q)):  # K10a
        b, _, i0, n = segment(s)
        nruns[s] = sum(head(b + i, i0 + i) for i in range(n))
    seg_run_off = np.concatenate(([0], np.cumsum(nruns))).astype(np.int64)
    pep_off = seg_run_off[seg_off]  # hits_pep_off_kernel
    n_runs = int(seg_run_off[-1])
    pep_hit_off = np.zeros(n_runs + 1, np.int64)
    run_entry = np.zeros(n_runs, np.int64)
    for s in range(len(seg_q)):  # K10b: warp-uniform loop over 32 hits, ballot prefix
        b, h0, i0, n = segment(s)
        done = seg_run_off[s]
        for j in range(0, n, 32):
            heads = [j + lane < n and head(b + j + lane, i0 + j + lane) for lane in range(32)]
            for lane in range(32):
                if heads[lane]:
                    p = done + sum(heads[:lane])
                    pep_hit_off[p] = h0 + j + lane
                    run_entry[p] = b + j + lane
            done += sum(heads)
    pep_hit_off[n_runs] = hit_off[nq]
    return pep_off, pep_hit_off, run_entry


def runs_by_groupby(base, mass_bits, q_begin, q_count):
    pep_off, pep_hit_off, run_entry, h = [0], [], [], 0
    for b, n in zip(q_begin, q_count):
        prev = None
        for e in range(b, b + n):
            key = (base[e], mass_bits[e])
            if key != prev:
                pep_hit_off.append(h)
                run_entry.append(e)
                prev = key
            h += 1
        pep_off.append(len(run_entry))
    return np.array(pep_off), np.array(pep_hit_off + [h]), np.array(run_entry)


@settings(max_examples=60, deadline=None)
@given(st.data())
def test_segment_numbering_equals_groupby(data):
    n = data.draw(st.integers(1, 2500))
    rng = np.random.default_rng(data.draw(st.integers(0, 2 ** 31)))
    # entries: variant groups of 1..40 entries (same peptide, same mass), sometimes the same peptide again with another mass
    base, mass = [], []
    while len(base) < n:
        g = int(rng.integers(1, 41))
        pep = int(rng.integers(0, 50))
        m = int(rng.integers(0, 1 << 20))
        base += [pep] * g
        mass += [m] * g
    base, mass = np.array(base[:n]), np.array(mass[:n])
    nq = data.draw(st.integers(0, 12))
    q_begin = rng.integers(0, n, size=nq)
    q_count = np.array([int(rng.integers(0, n - b + 1)) if rng.random() < 0.8 else 0 for b in q_begin], dtype=np.int64)
    a = runs_by_segments(base, mass, q_begin, q_count)
    b = runs_by_groupby(base, mass, q_begin, q_count)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_runs_straddling_segment_and_warp_boundaries():
    base = np.zeros(1000, np.int64)
    mass = np.zeros(1000, np.int64)
    base[300:] = 1                      # a run boundary inside the second segment
    mass[511:] = 7                      # ... and one right before a segment boundary (entry 511 = hit 511 of query 0)
    q_begin, q_count = np.array([0, 5, 256]), np.array([1000, 0, 300])
    pep_off, pep_hit_off, run_entry = runs_by_segments(base, mass, q_begin, q_count)
    assert pep_off.tolist() == [0, 3, 3, 6]
    assert run_entry.tolist() == [0, 300, 511, 256, 300, 511]
    assert pep_hit_off.tolist() == [0, 300, 511, 1000, 1044, 1255, 1300]
