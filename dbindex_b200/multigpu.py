"""Multi-GPU index build and query routing: one process per GPU, torch.distributed
for the plumbing (NCCL over NVLink on GPUs, gloo on CPU in the tests).

The reference shards its index by mass into `indexFactor` SQLite files
(DBIndexStoreSQLiteMult.java:55-56,215-217) and answers a query from the buckets its range
touches (:333-343).  Here a bucket is a GPU:

  1. every rank holds the whole residue buffer (3 GB even at TrEMBL scale) and digests its own
     range of start positions;
  2. global key histogram (all-reduce) -> equal-count splitters -> all-to-all of the records, so
     rank d receives one contiguous mass slice, rank-ordered = global emission order, which keeps
     "first occurrence"  global;
  3. local sort + merge: every rank owns the unique peptides (first occurrence, protein lists) of
     its base-mass slice; a peptide is named by its global id = rank offset + local row;
  4. differential mods: every rank lists the variant GROUPS of its own peptides (one record per
     peptide and sequence of shift classes -- they share one mass); the groups go through a second
     histogram -> splitters (weighted by variant count) -> all-to-all by VARIANT mass, and the
     receiver sorts and expands them.  To expand a foreign peptide's group a rank only needs where
     its residues are, so (gpos, len) of all unique peptides -- 6 bytes each -- are all-gathered;
     nothing else is replicated.  A hit whose base peptide lives elsewhere is resolved by its owner
     (`fetch_resolved`), the way a query is answered by the rank that owns its mass;
  5. queries are routed on the host with the same splitters; a range that straddles a splitter is
     answered by both neighbours, exactly like Mult.getSequences walking two buckets.

`ShardEngine` is the device side of one rank.  `GpuShardEngine` drives the C ABI (dbi_mg_*); the
tests plug in a CPU engine so that the orchestration runs under gloo without a GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

MG_BINS = 4096


# ---- pure host logic (unit-tested on CPU) -------------------------------------------------------
def pick_splitters(hist: np.ndarray, world: int) -> np.ndarray:
    """Equal-count bin splitters: rank d receives the bins [s[d-1], s[d]).  Bins are never split, so
    equal keys (equal masses) stay on one rank and the merge of equal peptides remains local."""
    total = int(hist.sum())
    cum = np.cumsum(hist.astype(np.int64))
    out = np.empty(max(world - 1, 0), dtype=np.uint32)
    for d in range(1, world):
        target = total * d // world
        # first bin boundary at which at least `target` items lie below
        out[d - 1] = int(np.searchsorted(cum, target, side="left")) + 1 if total else 0
    np.minimum(out, len(hist), out=out)
    return np.maximum.accumulate(out) if len(out) else out


def splitter_masses(bin_splitters: np.ndarray, shift: int, min_mass: float) -> np.ndarray:
    """The mass at which each splitter sits: radix key = bits(mass) - bits(min_mass)."""
    base = np.float64(min_mass).view(np.uint64)
    keys = (bin_splitters.astype(np.uint64) << np.uint64(shift)) + base
    return keys.view(np.float64)


def route_queries(lo: np.ndarray, hi: np.ndarray, split_mass: np.ndarray, rank: int) -> np.ndarray:
    """Indices of the queries whose [lo, hi] intersects rank's slice [split[rank-1], split[rank])."""
    left = split_mass[rank - 1] if rank > 0 else -np.inf
    right = split_mass[rank] if rank < len(split_mass) else np.inf
    return np.nonzero((hi >= left) & (lo < right))[0]


# ---- collectives on raw bytes -------------------------------------------------------------------
def _world() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def _all_to_all_rows(send: torch.Tensor, send_counts: Sequence[int], recv_counts: Sequence[int]) -> torch.Tensor:
    """all-to-all-v of a 1-D typed tensor split by element counts (rank-ordered receive).  Moved as
    raw bytes: NCCL has no 16-bit integer type and the payload is opaque anyway."""
    out = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
    if _world() == 1:
        out.copy_(send)
        return out
    w = send.element_size()
    dist.all_to_all_single(out.view(torch.uint8), send.contiguous().view(torch.uint8),
                           output_split_sizes=[int(c) * w for c in recv_counts],
                           input_split_sizes=[int(c) * w for c in send_counts])
    return out


def _gather_concat(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """Rank-order concatenation of every rank's 1-D tensor on every rank (variable sizes)."""
    return _gather_tables([local], [counts])[0]


def _gather_tables(tables: Sequence[torch.Tensor], counts: Sequence[Sequence[int]]) -> List[torch.Tensor]:
    """Rank-order concatenation of several 1-D tensors at once: every rank packs its tables into one
    byte buffer (padded to the largest rank), ONE all-gather moves them, and each table is cut out
    of the gathered rows.  counts[i][r] = elements of table i on rank r."""
    if _world() == 1:
        return list(tables)
    world = dist.get_world_size()
    dev = tables[0].device
    widths = [t.element_size() for t in tables]
    # byte offset of table i inside rank r's row (16-byte aligned so that typed views stay aligned)
    offs = []
    row_bytes = 0
    for r in range(world):
        o, cur = [], 0
        for i, w in enumerate(widths):
            o.append(cur)
            cur += (int(counts[i][r]) * w + 15) & ~15
        offs.append(o)
        row_bytes = max(row_bytes, cur)
    row_bytes = max(row_bytes, 16)
    rank = dist.get_rank()
    row = torch.empty(row_bytes, dtype=torch.uint8, device=dev)
    for i, t in enumerate(tables):
        n = t.numel() * widths[i]
        if n:
            row[offs[rank][i]:offs[rank][i] + n].copy_(t.contiguous().view(torch.uint8))
    rows = torch.empty(world * row_bytes, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(rows, row)
    rows = rows.view(world, row_bytes)
    out = []
    for i, t in enumerate(tables):
        parts = [rows[r, offs[r][i]:offs[r][i] + int(counts[i][r]) * widths[i]] for r in range(world) if counts[i][r]]
        cat = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=dev)
        out.append(cat.view(t.dtype))
    return out


def _exchange_counts(send_counts: np.ndarray, device) -> np.ndarray:
    world = _world()
    s = torch.tensor(send_counts.astype(np.int64), device=device)
    r = torch.empty(world, dtype=torch.int64, device=device)
    if world == 1:
        r.copy_(s)
    else:
        dist.all_to_all_single(r, s)
    return r.cpu().numpy()


def _all_gather_ints(vals: Sequence[int], device) -> np.ndarray:
    world = _world()
    t = torch.tensor(list(vals), dtype=torch.int64, device=device)
    if world == 1:
        return t.cpu().numpy()[None, :]
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.stack(out).cpu().numpy()


class ShardEngine:
    """Device side of one rank.  Arrays are 1-D torch tensors on `device` with signed dtypes of
    the same width as the C types (int64 for u64/f64 bits, int32 for u32, int16 for u16)."""

    device = torch.device("cpu")
    has_mods = False
    min_mass = 0.0

    def begin(self, rank: int, world: int): ...
    def digest(self) -> int: ...
    def histogram(self, stage: int) -> Tuple[torch.Tensor, int]: ...          # (int64[MG_BINS], shift)
    def partition(self, stage: int, splitters: np.ndarray) -> np.ndarray: ...  # send counts [world]
    def pack_send(self, stage: int) -> List[torch.Tensor]: ...
    def index_base(self, mass, gpos, prot, length): ...
    def export_unique(self) -> List[torch.Tensor]: ...   # gpos (int32), len (int16) of the own unique peptides
    def import_unique(self, rank_unique, tables: List[torch.Tensor]): ...   # their rank-order concatenation
    def finish(self): ...
    def own_tiles(self) -> Tuple[int, int]: ...                                # (tile_begin, n_tiles) of the own peptides
    def expand(self, tile_begin: int, n_tiles: int) -> int: ...
    def index_variants(self, key, payload): ...


def build_sharded(engine: ShardEngine) -> dict:
    """Run the staged multi-rank build on this rank.  Returns routing info:
    {"split_mass": masses at which the entry slices are cut, "bytes_sent": ..., ...}."""
    import time
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    dev = engine.device
    info = {"rank": rank, "world": world, "a2a_bytes": 0, "t": {}}
    t_last = [time.perf_counter()]

    def lap(name):  # host wall-clock per stage (every engine call ends synchronised)
        now = time.perf_counter()
        info["t"][name] = info["t"].get(name, 0.0) + 1e3 * (now - t_last[0])
        t_last[0] = now

    engine.begin(rank, world)
    engine.digest()
    lap("digest")

    info["a2a_ms"] = 0.0
    cuda = dev.type == "cuda"

    def exchange(stage: int, widths: Sequence[int]):
        hist, shift = engine.histogram(stage)
        if world > 1:
            dist.all_reduce(hist)
        splitters = pick_splitters(hist.cpu().numpy(), world)
        lap(f"hist{stage}")
        send_counts = engine.partition(stage, splitters)
        lap(f"partition{stage}")
        recv_counts = _exchange_counts(send_counts, dev)
        bufs = engine.pack_send(stage)
        lap(f"pack{stage}")
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        recv = [_all_to_all_rows(b, send_counts, recv_counts) for b in bufs]
        if cuda:
            e1.record()
            e1.synchronize()
            info["a2a_ms"] += e0.elapsed_time(e1)
        sent_off_rank = int(send_counts.sum() - send_counts[rank])
        info["a2a_bytes"] += sent_off_rank * int(sum(widths))
        lap(f"a2a{stage}")
        return recv, splitters, shift

    recv, base_split, shift = exchange(0, (8, 4, 4, 2))
    engine.index_base(*recv)
    del recv
    lap("index_base")
    if not engine.has_mods:
        n_u = engine.n_unique()
        info["unique_off"] = np.concatenate(([0], np.cumsum(_all_gather_ints([n_u], dev)[:, 0])))
        info["n_unique"] = int(info["unique_off"][-1])
        engine.finish()
        info["split_mass"] = splitter_masses(base_split, shift, engine.min_mass)
        return info

    tables = engine.export_unique()
    rank_unique = _all_gather_ints([int(tables[0].numel())], dev)[:, 0]
    gathered = _gather_tables(tables, [rank_unique] * len(tables))
    lap("gather_gpos_len")
    engine.import_unique(rank_unique, gathered)
    del tables, gathered
    lap("import_unique")
    info["unique_off"] = np.concatenate(([0], np.cumsum(rank_unique)))
    info["n_unique"] = int(rank_unique.sum())
    # every rank lists the groups of its own peptides: that work is small and nearly even; the
    # expensive part (sort + expansion) is balanced by the variant-weighted splitters below
    first_tile = 0
    tb, tn = engine.own_tiles()
    engine.expand(first_tile + tb, tn)
    lap("expand")
    recv, var_split, shift = exchange(1, (8, 8))
    engine.index_variants(*recv)
    lap("index_variants")
    info["split_mass"] = splitter_masses(var_split, shift, engine.min_mass)
    return info


# ---- the GPU engine ---------------------------------------------------------------------------
class GpuShardEngine(ShardEngine):
    """Drives libdbindex_gpu.so's dbi_mg_* entry points for one rank."""

    def __init__(self, index, device: torch.device):
        import ctypes as C
        self.C = C
        self.g = index           # dbindex_b200.GpuIndex with the (replicated) proteins added
        self.lib = index.lib
        self.device = device
        p = index.params
        self.has_mods = p.n_mods > 0 and p.max_mods_per_peptide > 0
        self.min_mass = float(p.min_mass)
        self._send = None
        vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
        sig = {
            "dbi_mg_begin": [vp, C.c_int, C.c_int],
            "dbi_mg_digest": [vp, u64p],
            "dbi_mg_histogram": [vp, C.c_int, vp, C.POINTER(C.c_int)],
            "dbi_mg_partition": [vp, C.c_int, vp, vp],
            "dbi_mg_pack_send": [vp, C.c_int, vp, vp, vp, vp],
            "dbi_mg_index_base": [vp, vp, vp, vp, vp, C.c_uint64],
            "dbi_mg_unique_counts": [vp, u64p, u64p],
            "dbi_mg_export_unique": [vp, vp, vp],
            "dbi_mg_import_unique": [vp, vp, vp, vp],
            "dbi_mg_finish": [vp],
            "dbi_mg_own_tiles": [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
            "dbi_mg_expand": [vp, C.c_uint32, C.c_uint32, u64p],
            "dbi_mg_index_variants": [vp, vp, vp, C.c_uint64],
        }
        for name, args in sig.items():
            fn = getattr(self.lib, name)
            fn.restype = C.c_int
            fn.argtypes = args

    def _ck(self, rc):
        self.g._check(rc)

    @staticmethod
    def _p(t: Optional[torch.Tensor]):
        return None if t is None or t.numel() == 0 else t.data_ptr()

    def _empty(self, n, dtype):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def begin(self, rank, world):
        self.world = world
        self._ck(self.lib.dbi_mg_begin(self.g._h, rank, world))

    def digest(self):
        n = self.C.c_uint64()
        self._ck(self.lib.dbi_mg_digest(self.g._h, self.C.byref(n)))
        self.n_local = n.value
        return n.value

    def histogram(self, stage):
        hist = torch.zeros(MG_BINS, dtype=torch.int64, device=self.device)
        shift = self.C.c_int()
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_histogram(self.g._h, stage, hist.data_ptr(), self.C.byref(shift)))
        return hist, shift.value

    def partition(self, stage, splitters):
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        counts = np.zeros(self.world, dtype=np.uint64)
        self._ck(self.lib.dbi_mg_partition(self.g._h, stage, sp.ctypes.data if len(sp) else None, counts.ctypes.data))
        self._n_stage = int(counts.sum())
        return counts

    def pack_send(self, stage):
        n = self._n_stage
        if stage == 0:
            bufs = [self._empty(n, torch.int64), self._empty(n, torch.int32), self._empty(n, torch.int32),
                    self._empty(n, torch.int16)]
        else:
            bufs = [self._empty(n, torch.int64), self._empty(n, torch.int64), None, None]
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_pack_send(self.g._h, stage, *[self._p(b) for b in bufs]))
        return [b for b in bufs if b is not None]

    def index_base(self, mass, gpos, prot, length):
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_index_base(self.g._h, self._p(mass), self._p(gpos), self._p(prot), self._p(length),
                                            int(mass.numel())))

    def n_unique(self):
        u, p = self.C.c_uint64(), self.C.c_uint64()
        self._ck(self.lib.dbi_mg_unique_counts(self.g._h, self.C.byref(u), self.C.byref(p)))
        return u.value

    def export_unique(self):
        u = self.n_unique()
        t = [self._empty(u, torch.int32), self._empty(u, torch.int16)]
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_export_unique(self.g._h, *[self._p(x) for x in t]))
        return t

    def import_unique(self, rank_unique, tables):
        ru = np.ascontiguousarray(rank_unique, dtype=np.uint64)
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_import_unique(self.g._h, ru.ctypes.data, *[self._p(x) for x in tables]))

    def finish(self):
        self._ck(self.lib.dbi_mg_finish(self.g._h))

    def own_tiles(self):
        t0, nt = self.C.c_uint32(), self.C.c_uint32()
        self._ck(self.lib.dbi_mg_own_tiles(self.g._h, self.C.byref(t0), self.C.byref(nt)))
        return t0.value, nt.value

    def expand(self, tile_begin, n_tiles):
        v = self.C.c_uint64()
        self._ck(self.lib.dbi_mg_expand(self.g._h, tile_begin, n_tiles, self.C.byref(v)))
        return v.value

    def index_variants(self, key, payload):
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_index_variants(self.g._h, self._p(key), self._p(payload), int(key.numel())))


# ---- hits whose base peptide lives on another rank ---------------------------------------------
REMOTE_BASE = 0xFFFFFFFF


def lookup_unique(g, gids: np.ndarray) -> dict:
    """dbi_mg_lookup_unique on the rank that owns `gids`: first occurrence and protein list of each."""
    import ctypes as C
    lib = g.lib
    lib.dbi_mg_lookup_unique.restype = C.c_int
    lib.dbi_mg_lookup_unique.argtypes = [C.c_void_p] * 2 + [C.c_uint64] + [C.c_void_p] * 5 + [C.c_uint64, C.c_void_p]
    gids = np.ascontiguousarray(gids, dtype=np.uint32)
    n = len(gids)
    prot, off = np.empty(n, np.uint32), np.empty(n, np.uint32)
    ln, plo = np.empty(n, np.uint16), np.empty(n + 1, np.uint64)
    n_ids = C.c_uint64()
    p = lambda a: a.ctypes.data if a.size else None  # noqa: E731
    g._check(lib.dbi_mg_lookup_unique(g._h, p(gids), n, None, None, None, None, None, 0, C.byref(n_ids)))
    ids = np.empty(n_ids.value, np.uint32)
    g._check(lib.dbi_mg_lookup_unique(g._h, p(gids), n, p(prot), p(off), p(ln), plo.ctypes.data, p(ids), len(ids),
                                      C.byref(n_ids)))
    return {"first_prot": prot, "first_off": off, "len": ln, "prot_list_off": plo, "prot_ids": ids}


def fetch_resolved(g, info: dict, begin: int, count: int, lookup=None) -> dict:
    """COLLECTIVE dbi_fetch of entries [begin, begin + count) of this rank's slice with every base
    peptide resolved: hits whose base lives on another rank (first_prot == DBI_REMOTE_BASE, first_off
    = global id) are answered by their owners.  A process that holds all the handles (the Java host)
    calls the owner's handle directly instead of this exchange."""
    f = g.fetch(begin, count)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return f
    lookup = lookup or lookup_unique
    world, rank = dist.get_world_size(), dist.get_rank()
    uoff = np.asarray(info["unique_off"], dtype=np.int64)
    remote = np.nonzero(f["first_prot"] == REMOTE_BASE)[0]
    gids = f["first_off"][remote].astype(np.int64)
    owner = np.searchsorted(uoff, gids, side="right") - 1
    ask = [np.unique(gids[owner == r]).astype(np.uint32) for r in range(world)]
    asked = [None] * world
    dist.all_gather_object(asked, ask)           # asked[src][dst] = ids src wants from dst
    answers = [lookup(g, asked[src][rank]) if len(asked[src][rank]) else None for src in range(world)]
    got = [None] * world
    dist.all_gather_object(got, answers)         # got[dst][src] = dst's answer to src
    plo = f["prot_list_off"].astype(np.int64)
    lists = [f["prot_ids"][plo[i]:plo[i + 1]] for i in range(len(plo) - 1)]
    for r in range(world):
        ans = got[r][rank]
        if ans is None:
            continue
        pos = {int(gid): k for k, gid in enumerate(ask[r])}
        alo = ans["prot_list_off"].astype(np.int64)
        for i in remote[owner == r]:
            k = pos[int(f["first_off"][i])]
            f["first_prot"][i], f["first_off"][i] = ans["first_prot"][k], ans["first_off"][k]
            lists[i] = ans["prot_ids"][alo[k]:alo[k + 1]]
    sizes = np.array([len(x) for x in lists], dtype=np.int64)
    f["prot_list_off"] = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
    f["prot_ids"] = np.concatenate(lists).astype(np.uint32) if len(lists) and sizes.sum() else np.zeros(0, np.uint32)
    return f
