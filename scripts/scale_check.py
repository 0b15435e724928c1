"""Large single-GPU runs of the BASELINE.json configs the oracle cannot finish in seconds:
cfg3 (200 k proteins, semi-tryptic), a per-GPU share of cfg4 (10^6 proteins, 3 variable mods) and
cfg5 (10^6 queries at 10 ppm and +-3 Da against that index).  Parity at these sizes is checked
through size-independent properties:
  * sortedness of the index over sampled chunks and across chunk boundaries;
  * membership: every index entry the oracle derives from a sample of proteins ALONE (same
    params) is found in the big index by a zero-tolerance query, with the same peptide string
    and mod pattern, and its protein list contains the sampled protein;
  * query bounds: for sampled queries, mass[begin-1] < lo <= mass[begin] and
    mass[begin+count-1] <= hi < mass[begin+count].
Prints one JSON line; run under gpurun, results are copied to profiles/.

usage: python scripts/scale_check.py --cfg 2|3 --proteins N [--queries Q] [--sample 64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dbindex_b200 as dbi  # noqa: E402
from dbindex_b200 import synth  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402  (checker only)

CFG = {
    2: dict(static_mods={"C": 57.02146}, diff_mods=[("M", 15.9949), ("STY", 79.96633)], max_mods_per_peptide=3),
    3: dict(semi=1),
}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2)
    ap.add_argument("--proteins", type=int, default=200_000)
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--sample", type=int, default=48, help="proteins checked against the oracle")
    ap.add_argument("--builds", type=int, default=2)
    args = ap.parse_args()

    t0 = time.perf_counter()
    res, off = synth.config_proteome(4 if args.cfg == 2 and args.proteins > 20000 else args.cfg, args.proteins)
    t_synth = time.perf_counter() - t0
    params = dbi.default_params(**CFG[args.cfg])
    params.profile = 1
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g = dbi.GpuIndex(params)
    g.set_stream(stream.cuda_stream)
    g.add_proteins(res, off)
    g.upload()
    build_ms = []
    for i in range(1 + args.builds):
        g.reset_index()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.build()
        b.record(stream)
        torch.cuda.synchronize()
        if i:
            build_ms.append(a.elapsed_time(b))
    st = g.stats()
    n = st["n_entries"]
    out = {"cfg": args.cfg, "proteins": args.proteins, "residues": int(len(res)), "emitted": st["n_emitted"],
           "unique": st["n_unique"], "entries": n, "build_ms": build_ms, "stage_ms": st["stage_ms"],
           "entries_per_s": n / (min(build_ms) / 1e3), "device_bytes": st["device_bytes"],
           "torch_mem_peak_gb": None, "synth_s": t_synth}
    free, total = torch.cuda.mem_get_info()
    out["gpu_mem_used_gb_after_build"] = (total - free) / 1e9
    rng = np.random.default_rng(7)

    # ---- sortedness over sampled chunks + chunk boundaries
    chunk = 1 << 20
    starts = sorted(set([0, max(0, n - chunk)] + [int(x) for x in rng.integers(0, max(1, n - chunk), size=24)]))
    prev_end = None
    for s in starts:
        c = min(chunk, n - s)
        m = g.fetch(s, c, with_ids=False)["mass"]
        assert np.all(np.diff(m) >= 0), f"index not sorted inside chunk at {s}"
        if prev_end is not None and s >= prev_end[0]:
            assert m[0] >= prev_end[1], "index not sorted across chunks"
        prev_end = (s + c, m[-1])
    out["sorted_chunks_checked"] = len(starts)

    # ---- membership of oracle-derived entries
    P = len(off) - 1
    pick = np.sort(rng.choice(P, size=min(args.sample, P), replace=False))
    checked = 0
    for pid in pick:
        r0, r1 = int(off[pid]), int(off[pid + 1])
        o = Oracle(params, threads=1)
        o.add_proteins(res[r0:r1], np.array([0, r1 - r0], dtype=np.uint64))
        assert o.build() == 0
        e = o.entries()
        if len(e["mass"]) == 0:
            continue
        sel = rng.choice(len(e["mass"]), size=min(40, len(e["mass"])), replace=False)
        lo = e["mass"][sel].copy()
        bq, cq = g.query(lo, lo)  # zero tolerance: the oracle's mass must be in the index bit for bit
        seq = res[r0:r1]
        for k, (b0, c0) in enumerate(zip(bq, cq)):
            i = sel[k]
            assert c0 >= 1, f"oracle entry of protein {pid} not found (mass {lo[k]!r})"
            want = seq[int(e["first_off"][i]):int(e["first_off"][i]) + int(e["len"][i])].tobytes()
            ok = False
            for h0 in range(int(b0), int(b0 + c0), 1 << 18):  # isomers share the mass bit for bit: runs can be long
                hit = g.fetch(h0, int(min(b0 + c0 - h0, 1 << 18)))
                plo = hit["prot_list_off"].astype(np.int64)
                cand = np.nonzero((hit["len"] == len(want)) & (hit["modpat"] == e["modpat"][i]))[0]
                for h in cand:
                    fp, fo = int(hit["first_prot"][h]), int(hit["first_off"][h])
                    got = res[int(off[fp]) + fo:int(off[fp]) + fo + len(want)].tobytes()
                    if got == want and pid in hit["prot_ids"][plo[h]:plo[h + 1]]:
                        ok = True
                        break
                if ok:
                    break
            assert ok, f"entry ({want!r}, pat {int(e['modpat'][i]):#x}) of protein {pid} missing from the index ({int(c0)} entries at that mass)"
            checked += 1
    out["oracle_entries_found"] = checked
    out["oracle_proteins_sampled"] = int(len(pick))

    # ---- cfg5: query sweep (half at 10 ppm, half at +-3 Da), queries resident in HBM
    nq = args.queries
    sample_m = np.concatenate([g.fetch(int(s), 4096, with_ids=False)["mass"]
                               for s in rng.integers(0, max(1, n - 4096), size=64)])
    _, _, lo, hi = synth.synth_queries(sample_m, nq, 20240605, da_fraction=0.5)
    d_lo, d_hi = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    d_b = torch.empty(nq, dtype=torch.int64, device="cuda")
    d_c = torch.empty(nq, dtype=torch.int64, device="cuda")
    q_ms = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.query_device(d_lo.data_ptr(), d_hi.data_ptr(), nq, d_b.data_ptr(), d_c.data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
        if i:
            q_ms.append(a.elapsed_time(b))
    hb, hc = d_b.cpu().numpy(), d_c.cpu().numpy()
    out["queries"] = {"n": nq, "ms": q_ms, "queries_per_s": nq / (min(q_ms) / 1e3), "hits": int(hc.sum()),
                      "mix": "half 10 ppm, half +-3 Da; half indexed masses +-5 ppm, half uniform decoys"}
    # host-call path (H2D + D2H inside)
    t1 = time.perf_counter()
    b2, c2 = g.query(lo, hi)
    out["queries"]["host_call_ms"] = 1e3 * (time.perf_counter() - t1)
    assert np.array_equal(b2.astype(np.int64), hb) and np.array_equal(c2.astype(np.int64), hc)
    for qi in rng.integers(0, nq, size=300):
        b0, c0 = int(hb[qi]), int(hc[qi])
        lo_i, hi_i = lo[qi], hi[qi]
        a0, a1 = max(0, b0 - 1), min(n, b0 + c0 + 1)
        m = g.fetch(a0, a1 - a0, with_ids=False)["mass"] if a1 - a0 <= (1 << 22) else None
        if m is None:  # very wide hit range: check the two ends only
            m_lo = g.fetch(a0, min(2, n - a0), with_ids=False)["mass"]
            m_hi = g.fetch(max(0, b0 + c0 - 1), min(2, n - max(0, b0 + c0 - 1)), with_ids=False)["mass"]
            if b0 > 0:
                assert m_lo[0] < lo_i
            if c0:
                assert m_lo[b0 - a0] >= lo_i and m_hi[0] <= hi_i
            if b0 + c0 < n:
                assert m_hi[-1] > hi_i
            continue
        if b0 > 0:
            assert m[0] < lo_i, "an entry below the hit range is >= lo"
        inner = m[b0 - a0:b0 - a0 + c0]
        assert np.all(inner >= lo_i) and np.all(inner <= hi_i), "hit outside [lo, hi]"
        if b0 + c0 < n:
            assert m[-1] > hi_i, "an entry above the hit range is <= hi"
    out["query_bounds_checked"] = 300
    print(json.dumps(out), flush=True)
    g.close()


if __name__ == "__main__":
    main()
