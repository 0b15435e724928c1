"""cfg4 / cfg5 on N GPUs (torchrun): ONE index over --proteins synthetic proteins (TrEMBL-like
config 4 seed), trypsin + 3 variable mods, mass-range sharded with the NCCL all-to-all, then a
routed query sweep.  Size-independent checks: every rank's slice is sorted, slices are ordered
across ranks and cut at the splitter masses, entry counts add up over ranks, and the entries the
oracle derives from sampled proteins alone are found (zero-tolerance query on the owning rank,
same peptide string + mod pattern, protein in the list).  Rank 0 prints one JSON line.

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
from bench import CFG2  # noqa: E402
from dbindex_b200 import synth  # noqa: E402
from dbindex_b200.multigpu import GpuShardEngine, build_sharded, fetch_resolved, route_queries  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402  (checker only)


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
    res, off = synth.config_proteome(4, args.proteins)  # identical on every rank
    t_synth = time.perf_counter() - t0
    params = dbi.default_params(**CFG2)
    params.device = local
    params.profile = 1
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g = dbi.GpuIndex(params)
    g.set_stream(stream.cuda_stream)
    g.add_proteins(res, off)
    g.upload()
    ms, a2a_ms, a2a_bytes, info = [], [], [], None
    for i in range(1 + args.builds):
        g.reset_index()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        info = build_sharded(GpuShardEngine(g, dev))
        b.record(stream)
        torch.cuda.synchronize()
        if i:
            ms.append(a.elapsed_time(b))
            a2a_ms.append(info["a2a_ms"])
            a2a_bytes.append(info["a2a_bytes"])
    st = g.stats()
    n = st["n_entries"]
    free, total = torch.cuda.mem_get_info()
    # ---- slice checks
    rng = np.random.default_rng(11 + rank)
    chunk = 1 << 20
    first = last = None
    for s in sorted(set([0, max(0, n - chunk)] + [int(x) for x in rng.integers(0, max(1, n - chunk), size=8)])):
        m = g.fetch(s, min(chunk, n - s), with_ids=False)["mass"]
        assert np.all(np.diff(m) >= 0), "slice not sorted"
        if s == 0:
            first = float(m[0])
        if s + len(m) == n:
            last = float(m[-1])
    ends = [None] * world
    dist.all_gather_object(ends, (n, first, last, float(min(ms)), float(max(a2a_ms)), int(a2a_bytes[-1]),
                                  (total - free) / 1e9))
    sm = info["split_mass"]
    if n:
        from dbindex_b200.multigpu import owned_mask
        assert bool(np.all(owned_mask(np.array([first, last]), sm, rank, world)))
    # ---- membership of oracle-derived entries (same sample on every rank; the owner must hold each)
    srng = np.random.default_rng(5)
    P = len(off) - 1
    found = np.zeros(0, np.int64)
    checked = 0
    for pid in np.sort(srng.choice(P, size=min(args.sample, P), replace=False)):
        r0, r1 = int(off[pid]), int(off[pid + 1])
        o = Oracle(params, threads=1)
        o.add_proteins(res[r0:r1], np.array([0, r1 - r0], dtype=np.uint64))
        assert o.build() == 0
        e = o.entries()
        if not len(e["mass"]):
            continue
        sel = srng.choice(len(e["mass"]), size=min(24, len(e["mass"])), replace=False)
        lo = e["mass"][sel].copy()
        bq, cq = g.query(lo, lo)
        ok = np.zeros(len(sel), np.int64)
        for k2, (b0, c0) in enumerate(zip(bq, cq)):
            i2 = sel[k2]
            want = res[r0 + int(e["first_off"][i2]):r0 + int(e["first_off"][i2]) + int(e["len"][i2])].tobytes()
            # COLLECTIVE: every rank fetches its (possibly empty) hit range; base peptides held by other
            # ranks are resolved by their owners
            hit = fetch_resolved(g, info, int(b0), int(min(c0, 1 << 16)))
            plo = hit["prot_list_off"].astype(np.int64)
            for h in np.nonzero((hit["len"] == len(want)) & (hit["modpat"] == e["modpat"][i2]))[0]:
                fp, fo = int(hit["first_prot"][h]), int(hit["first_off"][h])
                if res[int(off[fp]) + fo:int(off[fp]) + fo + len(want)].tobytes() == want and \
                        pid in hit["prot_ids"][plo[h]:plo[h + 1]]:
                    ok[k2] = 1
                    break
        found = np.concatenate([found, ok])
        checked += len(sel)
    tot = torch.from_numpy(found).cuda()
    dist.all_reduce(tot)
    assert bool((tot == 1).all()), "an oracle entry is missing (or held twice) across the ranks"
    # ---- cfg5: routed query sweep, half 10 ppm / half +-3 Da
    nq = args.queries
    samp = np.concatenate([g.fetch(int(s), 2048, with_ids=False)["mass"] for s in rng.integers(0, max(1, n - 2048), size=32)])
    gathered = [None] * world
    dist.all_gather_object(gathered, samp[::8].tolist())
    allm = np.sort(np.concatenate([np.asarray(x) for x in gathered]))
    _, _, lo, hi = synth.synth_queries(allm, nq, 20240605, da_fraction=0.5)
    sel = route_queries(lo, hi, sm, rank, world)
    d_lo, d_hi = torch.from_numpy(lo[sel]).cuda(), torch.from_numpy(hi[sel]).cuda()
    d_b = torch.empty(len(sel), dtype=torch.int64, device="cuda")
    d_c = torch.empty(len(sel), dtype=torch.int64, device="cuda")
    q_ms = []
    for i in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
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
        t_max = max(e[3] for e in ends)
        for r in range(world - 1):
            if ends[r][0] and ends[r + 1][0]:
                assert ends[r][2] <= ends[r + 1][1], "slices overlap across ranks"
        bytes_all = sum(e[5] for e in ends)
        a2a_t = max(e[4] for e in ends)
        print(json.dumps({
            "workload": "cfg4-style: trypsin, 2 missed cleavages, 600-6000 Da, static C, variable M / STY <= 3; "
                        "cfg5 sweep: half 10 ppm, half +-3 Da, routed to the owning GPU",
            "n_gpus": world, "proteins": args.proteins, "residues": int(len(res)), "entries_total": int(n_all),
            "entries_per_rank": [e[0] for e in ends], "build_ms_max_rank": t_max,
            "entries_per_s": n_all / (t_max / 1e3),
            "all_to_all": {"bytes_all_ranks": bytes_all, "ms_max_rank": a2a_t,
                           "bus_gbs_per_gpu": bytes_all / world / 1e9 / (a2a_t / 1e3)},
            "gpu_mem_used_gb": [round(e[6], 1) for e in ends],
            "host_stage_ms_rank0": {k: round(v, 2) for k, v in info["t"].items()},
            "oracle_entries_checked": int(checked), "synth_s": t_synth,
            "queries": {"n": nq, "routed": int(qt[2].item()), "ms_max_rank": float(qmax[0].item()),
                        "queries_per_s": nq / (float(qmax[0].item()) / 1e3), "hits": int(qt[1].item())},
        }), flush=True)
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
