"""Worker of tests/test_dist.py: one gloo rank of the sharded build, checked against the
single-process oracle.  Launched with torch.distributed.run."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dbindex_b200 as dbi  # noqa: E402
from dbindex_b200 import synth  # noqa: E402
from dbindex_b200.multigpu import build_sharded, owned_mask, route_queries, shard_proteins, shard_sizes  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402
from tests.cpu_engine import OracleShardEngine  # noqa: E402
from tests.util import PARAM_SETS, bits  # noqa: E402


class GpuEntries:
    """Adapter giving the GPU handle the two calls the checks below use."""

    def __init__(self, g, info):
        self.g = g
        self.info = info

    def entries(self):
        n = self.g.stats()["n_entries"]
        f = self.g.fetch(0, n)  # base peptides of other ranks are read through their mapped windows
        assert not np.any(f["first_prot"] == 0xFFFFFFFF), "a base peptide of another rank was not resolved"
        plo = f["prot_list_off"].astype(np.int64)
        f["plist"] = [tuple(f["prot_ids"][plo[i]:plo[i + 1]].tolist()) for i in range(n)]
        return f

    def query(self, lo, hi):
        b, c = self.g.query(lo, hi)
        return b.astype(np.int64), c.astype(np.int64)


def main():
    name, out_path = sys.argv[1], sys.argv[2]
    engine_kind = sys.argv[3] if len(sys.argv) > 3 else "cpu"
    n_prot = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    if engine_kind == "gpu":
        import torch
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    params = dbi.default_params(**PARAM_SETS[name])
    res, off = synth.synth_proteome(n_prot, 4711, median_len=150 if n_prot <= 100 else 400, min_len=5)
    # duplicated proteins in different shards: the merge must still be global
    seqs = [res[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
    seqs = seqs + [seqs[0], seqs[1]] + seqs[20:23]
    res = np.concatenate(seqs)
    off = np.zeros(len(seqs) + 1, np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])

    if engine_kind == "gpu":
        from dbindex_b200.multigpu import GpuShardEngine
        params.device = local
        g = dbi.GpuIndex(params)
        sres, soff, _ = shard_proteins(res, off, rank, world)  # a rank only ever sees its own shard of the FASTA
        g.add_proteins(sres, soff)
        info = build_sharded(GpuShardEngine(g, torch.device("cuda", local)), shard_sizes(off, world))
        eng = GpuEntries(g, info)
    else:
        eng = OracleShardEngine(params, res, off)
        # both forms: the shard sizes gathered by the build itself, or handed over by the caller that cut the FASTA
        info = build_sharded(eng, shard_sizes(off, world) if name == "cfg2_mods" else None)
    mine = eng.entries()

    # the single-process answer
    o = Oracle(params, threads=max(1, (os.cpu_count() or 2) // world))
    o.add_proteins(res, off)
    assert o.build() == 0
    exp = o.entries()
    n_mine = len(mine["mass"])
    counts = [None] * world
    dist.all_gather_object(counts, n_mine)
    assert sum(counts) == len(exp["mass"]), (counts, len(exp["mass"]))
    # what this rank must hold: the oracle's entries whose mass falls into one of its slices, in mass order
    sm = info["split_mass"]
    sl = np.nonzero(owned_mask(exp["mass"], sm, rank, world))[0]
    assert len(sl) == n_mine, (len(sl), n_mine)
    assert np.array_equal(bits(mine["mass"]), bits(exp["mass"][sl])), "slice masses differ from the global index"
    plo = exp["prot_list_off"].astype(np.int64)
    exp_t = sorted(zip(bits(exp["mass"][sl]).tolist(), exp["first_prot"][sl].tolist(), exp["first_off"][sl].tolist(),
                       exp["len"][sl].tolist(), exp["modpat"][sl].tolist(),
                       [tuple(exp["prot_ids"][plo[i]:plo[i + 1]].tolist()) for i in sl]))
    got_t = sorted(zip(bits(mine["mass"]).tolist(), mine["first_prot"].tolist(), mine["first_off"].tolist(),
                       mine["len"].tolist(), mine["modpat"].tolist(), mine["plist"]))
    assert got_t == exp_t, "slice entries differ (first occurrence / protein lists must be global)"
    # routed queries: each rank answers the queries touching its slice; the sum is the global answer
    _, _, lo, hi = synth.synth_queries(exp["mass"], 400, 3, da_fraction=0.5)
    ob, oc, _ = o.query(lo, hi)
    sel = route_queries(lo, hi, sm, rank, world)
    part = np.zeros(len(lo), np.int64)
    b, c = eng.query(lo[sel], hi[sel])
    part[sel] = c
    # queries NOT routed here must have no hits here
    nb, nc = eng.query(lo, hi)
    assert np.array_equal(nc, part), "routing dropped hits"
    allp = [None] * world
    dist.all_gather_object(allp, part)
    assert np.array_equal(np.sum(allp, axis=0), oc.astype(np.int64)), "routed hit counts differ from the global answer"
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump({"ok": True, "counts": counts, "a2a_bytes": info["a2a_bytes"], "engine": engine_kind}, f)
    if engine_kind == "gpu":
        dist.barrier()  # nobody unmaps / frees its windows while another rank may still read them
        g.close()
        print("dist_worker OK", name, engine_kind, "world", world, "entries per rank", counts, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
