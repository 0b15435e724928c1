"""world_size-2 gloo tests of the N>1 path on CPU: the orchestration of dbindex_b200/multigpu.py
(FASTA shards, histograms -> plan -> count matrix, the two exchanges, unique offsets, query routing)
with a CPU engine, against the single-process oracle; and the host arithmetic of dbi_mg_plan against its
numpy twin."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from dbindex_b200.multigpu import MG_BINS, pick_splitters, plan_exchange, route_queries, shard_proteins, splitter_masses

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pick_splitters_equal_count_and_monotone():
    rng = np.random.default_rng(0)
    hist = rng.integers(0, 1000, size=4096)
    hist[:300] = 0
    for world in (1, 2, 4, 8):
        s = pick_splitters(hist, world)
        assert len(s) == world - 1 and np.all(np.diff(s.astype(np.int64)) >= 0)
        edges = np.concatenate(([0], s, [4096])).astype(int)
        loads = [hist[edges[i]:edges[i + 1]].sum() for i in range(world)]
        assert sum(loads) == hist.sum()
        assert max(loads) - min(loads) <= 2 * hist.max()  # bins are never split
    assert pick_splitters(np.zeros(4096, np.int64), 4).tolist() == [0, 0, 0]


def test_route_queries_straddling():
    sm = splitter_masses(np.array([1000, 2000], np.uint32), 40, 600.0)
    assert sm[0] < sm[1]
    lo = np.array([600.0, sm[0] - 1, sm[0], sm[1] + 5, 0.0])
    hi = np.array([700.0, sm[0] + 1, sm[0], 9000.0, 1e9])
    assert route_queries(lo, hi, sm, 0, 3).tolist() == [0, 1, 4]
    assert route_queries(lo, hi, sm, 1, 3).tolist() == [1, 2, 4]
    assert route_queries(lo, hi, sm, 2, 3).tolist() == [3, 4]


def test_folded_slices_route_and_own():
    """2 * world folded slices: slice s lives on rank s < world ? s : 2 * world - 1 - s."""
    from dbindex_b200.multigpu import owned_mask, slice_owner
    assert slice_owner(np.arange(4), 4, 2).tolist() == [0, 1, 1, 0]
    assert slice_owner(np.arange(6), 6, 3).tolist() == [0, 1, 2, 2, 1, 0]
    assert slice_owner(np.arange(3), 3, 3).tolist() == [0, 1, 2]
    sm = np.array([1000.0, 2000.0, 3000.0])  # 4 slices on 2 ranks: rank 0 holds [.., 1000) and [3000, ..)
    m = np.array([600.0, 999.9, 1000.0, 2500.0, 3000.0, 5000.0])
    assert owned_mask(m, sm, 0, 2).tolist() == [True, True, False, False, True, True]
    assert owned_mask(m, sm, 1, 2).tolist() == [False, False, True, True, False, False]
    lo = np.array([600.0, 990.0, 1500.0, 2990.0, 3500.0, 0.0])
    hi = np.array([700.0, 1010.0, 1600.0, 3010.0, 3600.0, 9000.0])
    assert route_queries(lo, hi, sm, 0, 2).tolist() == [0, 1, 3, 4, 5]
    assert route_queries(lo, hi, sm, 1, 2).tolist() == [1, 2, 3, 5]
    # the planner with folded slices: send / receive counts follow the owners
    rng = np.random.default_rng(7)
    for world in (2, 3, 8):
        plain = rng.integers(0, 50, size=MG_BINS).astype(np.uint64)
        hg = np.concatenate([plain * np.uint64(7), plain, plain * np.uint64(3)])
        split, send, recv = plan_exchange(world, hg, hg, 42, 600.0, cost=[400.0, 125.0, 17.5, 20.0], n_slices=2 * world)
        assert len(split) == 2 * world - 1 and np.all(np.diff(split.astype(np.int64)) >= 0)
        edges = np.concatenate(([0], split, [MG_BINS])).astype(int)
        want = np.zeros(world, np.int64)
        for s_ in range(2 * world):
            want[int(slice_owner(s_, 2 * world, world))] += int(plain[edges[s_]:edges[s_ + 1]].sum())
        assert send.tolist() == want.tolist() and recv.tolist() == want.tolist()


@pytest.mark.parametrize("name", ["cfg1_tryptic", "cfg2_mods", "semi_nocut_mods"])
def test_two_rank_gloo_build_matches_oracle(name, tmp_path):
    out = tmp_path / "result.json"
    port = 29600 + (abs(hash(name)) % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), name, str(out)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    r = json.loads(out.read_text())
    assert r["ok"] and all(c > 0 for c in r["counts"]) and r["a2a_bytes"] > 0


def _phase_objective(hg, split, cost, shift=42, min_mass=600.0):
    """max over ranks of every phase's cost, summed over the phases (what dbi_mg_plan minimises)."""
    w, items, grp = (hg[i * MG_BINS:(i + 1) * MG_BINS].astype(np.float64) for i in range(3))
    base = np.float64(min_mass).view(np.uint64)
    edges_m = ((np.arange(MG_BINS + 1, dtype=np.uint64) << np.uint64(shift)) + base).view(np.float64)
    width = np.diff(edges_m)
    q = np.where(w > 0, cost[3] * w * w * 0.5 * (edges_m[:-1] + edges_m[1:]) / (width * max(w.sum(), 1.0)), 0.0)
    phases = [cost[0] * items, cost[1] * grp + cost[2] * w, q]
    e = np.concatenate(([0], split, [MG_BINS])).astype(int)
    return sum(max(ph[e[d]:e[d + 1]].sum() for d in range(len(e) - 1)) for ph in phases)


def test_dbi_mg_plan_cost_models():
    """dbi_mg_plan: with the equal-weight model the splitters equal the numpy reference (pick_splitters);
    send / receive counts partition the plain histograms; the query-aware model moves the cuts towards the
    dense bins (fewer entries for the ranks that will see more hits) and stays monotone; with several phases
    the cuts minimise the sum of the per-phase maxima at least as well as equal-sum cuts do."""
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8, 16):
        plain = [rng.integers(0, 50, size=MG_BINS).astype(np.uint64) for _ in range(world)]
        for p in plain:
            p[:100] = 0
        weighted = [p * rng.integers(1, 2000, size=MG_BINS).astype(np.uint64) for p in plain]
        groups = [p * rng.integers(1, 15, size=MG_BINS).astype(np.uint64) for p in plain]
        hg = np.concatenate([sum(weighted), sum(plain), sum(groups)]).astype(np.uint64)
        for r in range(world):
            hl = np.concatenate([weighted[r], plain[r], groups[r]]).astype(np.uint64)
            split, send, recv = plan_exchange(world, hg, hl, 42, 600.0, cost=[0.0, 0.0, 1.0, 0.0])
            assert split.tolist() == pick_splitters(hg[:MG_BINS], world).tolist()
            edges = np.concatenate(([0], split, [MG_BINS])).astype(int)
            assert send.tolist() == [int(plain[r][edges[d]:edges[d + 1]].sum()) for d in range(world)]
            assert recv.tolist() == [int(hg[MG_BINS:2 * MG_BINS][edges[d]:edges[d + 1]].sum()) for d in range(world)]
            assert int(recv.sum()) == int(hg[MG_BINS:2 * MG_BINS].sum())
            # the two-histogram form (no group estimates) is accepted and means the same
            s2h, _, _ = plan_exchange(world, hg[:2 * MG_BINS], hl[:2 * MG_BINS], 42, 600.0, cost=[0.0, 0.0, 1.0, 0.0])
            assert s2h.tolist() == split.tolist()
        for stage in (0, 1):
            s2, send, recv = plan_exchange(world, hg, np.concatenate([weighted[0], plain[0], groups[0]]), 42, 600.0,
                                           stage=stage, has_mods=True)
            assert np.all(np.diff(s2.astype(np.int64)) >= 0) and len(s2) == world - 1
            assert int(send.sum()) == int(plain[0].sum()) and int(recv.sum()) == int(hg[MG_BINS:2 * MG_BINS].sum())
    # a density peak: the query-aware model gives the rank holding the peak a narrower slice
    w = np.full(MG_BINS, 1000, dtype=np.uint64)
    w[3000:3100] = 40000
    hg = np.concatenate([w, np.full(MG_BINS, 100, dtype=np.uint64), np.zeros(MG_BINS, np.uint64)])
    flat, _, _ = plan_exchange(4, hg, hg, 42, 600.0, cost=[0.0, 0.0, 1.0, 0.0])
    aware, _, _ = plan_exchange(4, hg, hg, 42, 600.0, cost=[0.0, 0.0, 1.0, 50.0])
    width = lambda s: np.diff(np.concatenate(([0], s.astype(np.int64), [MG_BINS])))  # noqa: E731
    k = int(np.searchsorted(flat, 3050))
    k2 = int(np.searchsorted(aware, 3050))
    assert width(aware)[k2] < width(flat)[k]
    # records crowd the light end, entries the heavy end (what differential mods do): cuts of equal SUM leave
    # the first rank with the slowest base phase and the last with the slowest variant phase; the planner must
    # do at least as well on the sum of the phase maxima, and better here
    x = np.arange(MG_BINS, dtype=np.float64) / MG_BINS
    items = (2000 * np.exp(-3 * x)).astype(np.uint64)
    wgt = (items * (1 + 400 * x ** 3)).astype(np.uint64)
    grp = (items * (1 + 12 * x)).astype(np.uint64)
    hg = np.concatenate([wgt, items, grp])
    cost = [400.0, 125.0, 17.5, 0.0]
    total = cost[0] * items + cost[1] * grp + cost[2] * wgt
    cum = np.cumsum(total)
    for world in (2, 4, 8):
        equal_sum = np.array([int(np.searchsorted(cum, cum[-1] * d / world)) + 1 for d in range(1, world)])
        got, _, _ = plan_exchange(world, hg, hg, 42, 600.0, cost=cost)
        assert np.all(np.diff(got.astype(np.int64)) >= 0)
        assert _phase_objective(hg, got, cost) <= _phase_objective(hg, equal_sum, cost) * (1 + 1e-9)
    assert _phase_objective(hg, got, cost) < 0.97 * _phase_objective(hg, equal_sum, cost)


def test_shard_proteins_covers_the_proteome_in_order():
    from dbindex_b200 import synth
    res, off = synth.synth_proteome(57, 3, median_len=80, min_len=5)
    for world in (1, 2, 5, 8):
        got, first = [], 0
        for r in range(world):
            sres, soff, p0 = shard_proteins(res, off, r, world)
            assert p0 == first and soff[0] == 0
            first += len(soff) - 1
            got.append(sres)
        assert first == len(off) - 1 and np.array_equal(np.concatenate(got), res)


def test_plan_matrix_equals_per_rank_plans():
    """dbi_mg_plan_matrix (every rank's local histograms -> cuts + whole count matrix) agrees with dbi_mg_plan run
    on the summed histogram with each rank's own one."""
    from dbindex_b200.multigpu import plan_matrix
    rng = np.random.default_rng(11)
    for world, fold in ((1, 1), (2, 2), (3, 1), (8, 2), (16, 2)):
        hists = np.stack([np.concatenate([p * np.uint64(9), p, p * np.uint64(4)]) for p in
                          (rng.integers(0, 40, size=MG_BINS).astype(np.uint64) for _ in range(world))])
        split, matrix = plan_matrix(world, hists, 42, 600.0, stage=0, has_mods=True, n_slices=fold * world)
        hg = hists.sum(axis=0, dtype=np.uint64)
        for r in range(world):
            s2, send, recv = plan_exchange(world, hg, hists[r], 42, 600.0, stage=0, has_mods=True, n_slices=fold * world)
            assert s2.tolist() == split.tolist()
            assert send.tolist() == matrix[r].tolist() and recv.tolist() == matrix.sum(axis=0).tolist()


def test_planner_degenerate_histograms():
    """dbi_mg_plan / dbi_mg_plan_matrix on inputs a build can produce at the edges: nothing at all, everything in
    one bin, items only in the last bin, more ranks than occupied bins -- cuts stay ascending and inside the bin
    range, counts are conserved, nothing divides by zero."""
    from dbindex_b200.multigpu import plan_matrix
    B = MG_BINS
    cases = {
        "empty": np.zeros(3 * B, np.uint64),
        "one_bin": np.concatenate([np.eye(1, B, 1234, dtype=np.uint64)[0] * np.uint64(10 ** 6)] * 3),
        "last_bin": np.concatenate([np.eye(1, B, B - 1, dtype=np.uint64)[0] * np.uint64(7)] * 3),
        "three_bins": np.concatenate([(np.eye(1, B, 5, dtype=np.uint64)[0] + np.eye(1, B, 2000, dtype=np.uint64)[0] +
                                       np.eye(1, B, 4000, dtype=np.uint64)[0]) * np.uint64(100)] * 3),
    }
    for name, hg in cases.items():
        for world in (1, 2, 8, 16):
            for fold in (1, 2):
                if world == 1 and fold == 2:
                    continue
                for cost in ([0.0, 0.0, 1.0, 0.0], [400.0, 125.0, 17.5, 20.0], [1.0, 0.0, 0.0, 0.0]):
                    split, send, recv = plan_exchange(world, hg, hg, 42, 600.0, cost=cost, n_slices=fold * world)
                    assert len(split) == fold * world - 1 and np.all(np.diff(split.astype(np.int64)) >= 0), (name, world, fold)
                    assert np.all(split <= B)
                    assert int(send.sum()) == int(hg[B:2 * B].sum()) == int(recv.sum()), (name, world, fold)
                    s2, matrix = plan_matrix(world, np.stack([hg] + [np.zeros_like(hg)] * (world - 1)), 42, 600.0, cost=cost,
                                             n_slices=fold * world)
                    assert s2.tolist() == split.tolist() and matrix[0].tolist() == send.tolist()
                    assert int(matrix[1:].sum()) == 0
