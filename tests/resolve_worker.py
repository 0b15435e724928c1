"""Worker of tests/test_dist.py::test_fetch_resolved_two_ranks: two gloo ranks, fake index handles.
Rank r owns the unique peptides with global ids [10 r, 10 r + 10); each rank's slice holds entries
whose base peptides belong to BOTH ranks.  fetch_resolved must hand back every entry with its
owner's first occurrence and protein list."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dbindex_b200.multigpu import REMOTE_BASE, _gather_tables, fetch_resolved  # noqa: E402


def truth(gid):  # what the owner knows about unique peptide gid
    return {"first_prot": 100 + gid, "first_off": 7 * gid, "len": 6 + gid % 5, "ids": [100 + gid] + [gid] * (gid % 3)}


class FakeIndex:
    def __init__(self, rank, bases):
        self.rank, self.bases = rank, np.asarray(bases, dtype=np.int64)

    def fetch(self, begin, count):
        b = self.bases[begin:begin + count]
        own = (b // 10) == self.rank
        t = [truth(int(g)) for g in b]
        lists = [x["ids"] if o else [] for x, o in zip(t, own)]
        return {"mass": 1000.0 + b.astype(np.float64), "modpat": np.zeros(len(b), np.uint32),
                "len": np.array([x["len"] for x in t], np.uint16),
                "first_prot": np.array([x["first_prot"] if o else REMOTE_BASE for x, o in zip(t, own)], np.uint32),
                "first_off": np.array([x["first_off"] if o else g for x, o, g in zip(t, own, b)], np.uint32),
                "prot_list_off": np.concatenate(([0], np.cumsum([len(x) for x in lists]))).astype(np.uint64),
                "prot_ids": np.array([i for x in lists for i in x], np.uint32)}


def fake_lookup(g, gids):
    assert all(int(x) // 10 == g.rank for x in gids), "a lookup reached the wrong owner"
    t = [truth(int(x)) for x in gids]
    return {"first_prot": np.array([x["first_prot"] for x in t], np.uint32),
            "first_off": np.array([x["first_off"] for x in t], np.uint32),
            "len": np.array([x["len"] for x in t], np.uint16),
            "prot_list_off": np.concatenate(([0], np.cumsum([len(x["ids"]) for x in t]))).astype(np.uint64),
            "prot_ids": np.array([i for x in t for i in x["ids"]], np.uint32)}


def main():
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    bases = [3, 15, 15, 4, 19, 10, 2] if rank == 0 else [12, 0, 9, 9, 18]
    g = FakeIndex(rank, bases)
    info = {"unique_off": np.array([0, 10, 20])}
    for begin, count in ((0, len(bases)), (2, 3), (1, 0)):  # same number of collective calls on both ranks
        f = fetch_resolved(g, info, begin, count, lookup=fake_lookup)
        plo = f["prot_list_off"].astype(np.int64)
        for i, gid in enumerate(bases[begin:begin + count]):
            t = truth(gid)
            assert int(f["first_prot"][i]) == t["first_prot"] and int(f["first_off"][i]) == t["first_off"], (rank, i, gid)
            assert int(f["len"][i]) == t["len"]
            assert f["prot_ids"][plo[i]:plo[i + 1]].tolist() == t["ids"], (rank, i, gid)
        assert len(plo) == count + 1
    # the fused all-gather of several tables with different widths and uneven (also empty) sizes
    for sizes in ([5, 3], [0, 4], [7, 0], [0, 0]):
        n = sizes[rank]
        a = torch.arange(n, dtype=torch.int32) + 1000 * rank
        b = (torch.arange(n, dtype=torch.int16) * 3 + rank).to(torch.int16)
        c = torch.arange(2 * n, dtype=torch.int64) - 7 * rank           # a table with its own counts
        ga, gb, gc = _gather_tables([a, b, c], [sizes, sizes, [2 * x for x in sizes]])
        ea = torch.cat([torch.arange(sizes[r], dtype=torch.int32) + 1000 * r for r in range(2)])
        eb = torch.cat([(torch.arange(sizes[r], dtype=torch.int16) * 3 + r).to(torch.int16) for r in range(2)])
        ec = torch.cat([torch.arange(2 * sizes[r], dtype=torch.int64) - 7 * r for r in range(2)])
        assert torch.equal(ga, ea) and torch.equal(gb, eb) and torch.equal(gc, ec), (rank, sizes)
    dist.barrier()
    if rank == 0:
        open(sys.argv[1], "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
