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
    assert route_queries(lo, hi, sm, 0).tolist() == [0, 1, 4]
    assert route_queries(lo, hi, sm, 1).tolist() == [1, 2, 4]
    assert route_queries(lo, hi, sm, 2).tolist() == [3, 4]


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


def test_dbi_mg_plan_equals_the_numpy_plan():
    """dbi_mg_plan (C, what dbi_mg_build_local and a Java host use) == plan_exchange (numpy, what the
    torch.distributed orchestration uses): splitters, send counts, receive totals."""
    import ctypes as C
    import dbindex_b200 as dbi
    lib = dbi.load_library()
    lib.dbi_mg_plan.restype = C.c_int
    lib.dbi_mg_plan.argtypes = [C.c_int] + [C.c_void_p] * 5
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 8, 16):
        plain = [rng.integers(0, 50, size=MG_BINS).astype(np.uint64) for _ in range(world)]
        for p in plain:
            p[:100] = 0
        weighted = [p * rng.integers(1, 2000, size=MG_BINS).astype(np.uint64) for p in plain]
        hg = np.concatenate([sum(weighted), sum(plain)]).astype(np.uint64)
        for r in range(world):
            hl = np.concatenate([weighted[r], plain[r]]).astype(np.uint64)
            split, send, recv = plan_exchange(world, hg, hl)
            c_split = np.zeros(max(world - 1, 1), np.uint32)
            c_send, c_recv = np.zeros(world, np.uint64), np.zeros(world, np.uint64)
            assert lib.dbi_mg_plan(world, hg.ctypes.data, hl.ctypes.data, c_split.ctypes.data, c_send.ctypes.data,
                                   c_recv.ctypes.data) == 0
            assert c_split[:world - 1].tolist() == split.tolist()
            assert c_send.tolist() == send.tolist() and c_recv.tolist() == recv.tolist()
            assert int(c_recv.sum()) == int(hg[MG_BINS:].sum())


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
