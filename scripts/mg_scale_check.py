"""cfg4 / cfg5 on N GPUs (torchrun): ONE index over --proteins synthetic proteins (TrEMBL-like
config 4 seed), trypsin + 3 variable mods.  The FASTA is sharded by protein (every rank adds only its
shard), the records and variant groups travel inside the two fused multisplit / peer-memory scatter
kernels into folded mass slices, then a routed query sweep (cfg5: half 10 ppm, half +-3 Da).
Size-independent checks (bench.sharded_parity): every rank's entries are sorted and lie inside the
slices it holds, entry counts add up over the ranks, and the entries the oracle derives from sampled
proteins alone are held by exactly one rank (zero-tolerance query, same peptide string + mod pattern,
protein in the list).  Rank 0 prints one JSON line.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/mg_scale_check.py --proteins P
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dbindex_b200 as dbi  # noqa: E402
from bench import CFG2, sharded_parity  # noqa: E402
from dbindex_b200 import synth  # noqa: E402
from dbindex_b200.multigpu import GpuShardEngine, build_sharded, route_queries, shard_proteins, shard_sizes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--proteins", type=int, default=400_000, help="total over all ranks")
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--builds", type=int, default=2)
    ap.add_argument("--sample", type=int, default=16)
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t0 = time.perf_counter()
    res, off = synth.config_proteome(4, args.proteins)  # generated on every rank; each keeps only its shard
    t_synth = time.perf_counter() - t0
    sres, soff, _ = shard_proteins(res, off, rank, world)
    params = dbi.default_params(**CFG2)
    params.device = local
    params.profile = 1
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g = dbi.GpuIndex(params)
    g.set_stream(stream.cuda_stream)
    g.add_proteins(sres, soff)
    g.upload()
    ms, a2a_ms, a2a_bytes, info = [], [], [], None
    for i in range(1 + args.builds):
        g.reset_index()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        info = build_sharded(GpuShardEngine(g, dev), shard_sizes(off, world))
        b.record(stream)
        torch.cuda.synchronize()
        if i:
            ms.append(a.elapsed_time(b))
            a2a_ms.append(info["a2a_ms"])
            a2a_bytes.append(info["a2a_bytes"])
    st = g.stats()
    n = st["n_entries"]
    free, total = torch.cuda.mem_get_info()
    ends = [None] * world
    dist.all_gather_object(ends, (n, float(min(ms)), float(max(a2a_ms)), int(a2a_bytes[-1]), (total - free) / 1e9))
    # ---- order, slices, membership of oracle-derived entries (the check every bench line carries)
    parity = sharded_parity(torch, dist, dbi, g, info, res, off, params, rank, world, n_sample=args.sample)
    assert parity["failed"] == 0, parity
    # ---- cfg5: routed query sweep, half 10 ppm / half +-3 Da
    nq = args.queries
    rng = np.random.default_rng(11 + rank)
    samp = (np.concatenate([g.fetch(int(s), 2048, with_ids=False)["mass"] for s in rng.integers(0, max(1, n - 2048), size=32)])
            if n > 2048 else g.fetch(0, n, with_ids=False)["mass"])
    gathered = [None] * world
    dist.all_gather_object(gathered, samp[::8].tolist())
    allm = np.sort(np.concatenate([np.asarray(x) for x in gathered]))
    _, _, lo, hi = synth.synth_queries(allm, nq, 20240605, da_fraction=0.5)
    sel = route_queries(lo, hi, info["split_mass"], rank, world)
    d_lo, d_hi = torch.from_numpy(lo[sel]).cuda(), torch.from_numpy(hi[sel]).cuda()
    d_b = torch.empty(max(1, len(sel)), dtype=torch.int64, device="cuda")
    d_c = torch.zeros(max(1, len(sel)), dtype=torch.int64, device="cuda")
    q_ms = []
    for i in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        if len(sel):
            g.query_device(d_lo.data_ptr(), d_hi.data_ptr(), len(sel), d_b.data_ptr(), d_c.data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
        if i:
            q_ms.append(a.elapsed_time(b))
    qt = torch.tensor([min(q_ms), float(d_c.sum().item()), float(len(sel))], device="cuda", dtype=torch.float64)
    qmax = qt.clone()
    dist.all_reduce(qmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(qt)
    if rank == 0:
        n_all = sum(e[0] for e in ends)
        t_max = max(e[1] for e in ends)
        bytes_all = sum(e[3] for e in ends)
        a2a_t = max(e[2] for e in ends)
        print(json.dumps({
            "workload": "cfg4-style: trypsin, 2 missed cleavages, 600-6000 Da, static C, variable M / STY <= 3; "
                        "cfg5 sweep: half 10 ppm, half +-3 Da, routed to the owning GPU",
            "n_gpus": world, "proteins": args.proteins, "residues": int(len(res)), "entries_total": int(n_all),
            "entries_per_rank": [e[0] for e in ends], "build_ms_max_rank": t_max,
            "entries_per_s": n_all / (t_max / 1e3), "n_slices": info.get("n_slices", world),
            "all_to_all": {"bytes_all_ranks": bytes_all, "ms_max_rank": a2a_t,
                           "bus_gbs_per_gpu": bytes_all / world / 1e9 / max(a2a_t / 1e3, 1e-12)},
            "gpu_mem_used_gb": [round(e[4], 1) for e in ends],
            "host_stage_ms_rank0": {k: round(v, 2) for k, v in info["t"].items()},
            "parity": parity, "synth_s": t_synth,
            "queries": {"n": nq, "routed": int(qt[2].item()), "ms_max_rank": float(qmax[0].item()),
                        "queries_per_s": nq / (float(qmax[0].item()) / 1e3), "hits": int(qt[1].item())},
        }), flush=True)
    dist.barrier()  # nobody releases its windows while another rank may still read them
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
