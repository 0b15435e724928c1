"""world_size-2 gloo tests of the N>1 path on CPU: the orchestration of dbindex_b200/multigpu.py
(splitters, all-to-all-v, the fused all-gather of (gpos, len), group exchange by variant mass,
query routing, resolution of hits whose base peptide lives on another rank) with a CPU engine,
against the single-process oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from dbindex_b200.multigpu import pick_splitters, route_queries, splitter_masses

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


def test_fetch_resolved_two_ranks(tmp_path):
    """Hits whose base peptide is held by the other rank come back as DBI_REMOTE_BASE + global id and
    are completed by their owner (multigpu.fetch_resolved), lists and first occurrences intact."""
    out = tmp_path / "ok.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29871", os.path.join(ROOT, "tests", "resolve_worker.py"), str(out)]
    p = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"), capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert out.read_text() == "ok"
