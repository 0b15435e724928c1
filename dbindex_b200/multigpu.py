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
  3. local sort + merge; the unique tables are then replicated to every rank (they are ~20 B per
     unique peptide), so any rank can materialise any hit locally;
  4. differential mods: base tiles are re-dealt by VARIANT count (the heavy slice has orders of
     magnitude more variants per peptide), expanded, and the variants go through a second
     histogram -> splitters -> all-to-all by variant mass, then a local sort;
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


def balance_tiles(tile_counts: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous tile ranges with near-equal total count: [(begin, n_tiles)] per rank."""
    n = len(tile_counts)
    cum = np.concatenate(([0], np.cumsum(tile_counts.astype(np.int64))))
    total = int(cum[-1])
    cuts = [0]
    for d in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * d // world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.minimum(cuts, n))
    return [(int(cuts[d]), int(cuts[d + 1] - cuts[d])) for d in range(world)]


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
    def export_unique(self) -> List[torch.Tensor]: ...   # mass, gpos, prot, len, pcnt, plist
    def import_unique(self, rank_unique, rank_plist, tables: List[torch.Tensor]): ...
    def finish(self): ...
    def own_tiles(self) -> Tuple[int, int]: ...                                # (tile_begin, n_tiles) of the own slice
    def mod_tile_counts(self) -> Tuple[int, torch.Tensor]: ...                # (tile_begin, int32 counts)
    def expand(self, tile_begin: int, n_tiles: int) -> int: ...
    def index_variants(self, key, payload): ...


def build_sharded(engine: ShardEngine, rebalance: bool = False) -> dict:
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
    tables = engine.export_unique()
    n_u, n_p = int(tables[0].numel()), int(tables[5].numel())
    cnt = _all_gather_ints([n_u, n_p], dev)
    rank_unique, rank_plist = cnt[:, 0], cnt[:, 1]
    gathered = _gather_tables(tables, [rank_plist if i == 5 else rank_unique for i in range(len(tables))])
    lap("replicate_tables")
    engine.import_unique(rank_unique, rank_plist, gathered)
    del tables, gathered
    lap("import_unique")
    info["n_unique"] = int(rank_unique.sum())
    if not engine.has_mods:
        engine.finish()
        info["split_mass"] = splitter_masses(base_split, shift, engine.min_mass)
        return info

    if rebalance:
        # re-deal the base tiles by variant count (the heavy slices hold more variants per peptide)
        t0, tc = engine.mod_tile_counts()
        meta = _all_gather_ints([t0, int(tc.numel())], dev)
        all_counts = _gather_concat(tc, meta[:, 1]).cpu().numpy()
        # ranks own ascending slices, so the concatenation is in tile order starting at tile meta[0, 0]
        first_tile = int(meta[0, 0]) if len(all_counts) else 0
        # cost model of the expansion: one unit per variant plus a fixed per-peptide part (site scan,
        # table loads) worth ~64 variants -- light slices hold many cheap peptides
        ranges = balance_tiles(all_counts.astype(np.int64) + 64 * 256, world)
        tb, tn = ranges[rank]
        lap("tile_counts")
    else:
        # every rank lists the groups of its own slice: that work is small and nearly even; the
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
            "dbi_mg_export_unique": [vp, vp, vp, vp, vp, vp, vp],
            "dbi_mg_import_unique": [vp, vp, vp, vp, vp, vp, vp, vp, vp],
            "dbi_mg_finish": [vp],
            "dbi_mg_mod_tile_counts": [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp],
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

    def export_unique(self):
        u, p = self.C.c_uint64(), self.C.c_uint64()
        self._ck(self.lib.dbi_mg_unique_counts(self.g._h, self.C.byref(u), self.C.byref(p)))
        t = [self._empty(u.value, torch.int64), self._empty(u.value, torch.int32), self._empty(u.value, torch.int32),
             self._empty(u.value, torch.int16), self._empty(u.value, torch.int32), self._empty(p.value, torch.int32)]
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_export_unique(self.g._h, *[self._p(x) for x in t]))
        return t

    def import_unique(self, rank_unique, rank_plist, tables):
        ru = np.ascontiguousarray(rank_unique, dtype=np.uint64)
        rp = np.ascontiguousarray(rank_plist, dtype=np.uint64)
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_import_unique(self.g._h, ru.ctypes.data, rp.ctypes.data, *[self._p(x) for x in tables]))

    def finish(self):
        self._ck(self.lib.dbi_mg_finish(self.g._h))

    def own_tiles(self):
        t0, nt = self.C.c_uint32(), self.C.c_uint32()
        self._ck(self.lib.dbi_mg_own_tiles(self.g._h, self.C.byref(t0), self.C.byref(nt)))
        return t0.value, nt.value

    def mod_tile_counts(self):
        st = self.g.stats()
        cap = int(st["n_unique"]) // 256 + 4
        buf = torch.zeros(cap, dtype=torch.int32, device=self.device)
        t0, nt = self.C.c_uint32(), self.C.c_uint32()
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_mod_tile_counts(self.g._h, self.C.byref(t0), self.C.byref(nt), buf.data_ptr()))
        return t0.value, buf[:nt.value].clone()

    def expand(self, tile_begin, n_tiles):
        v = self.C.c_uint64()
        self._ck(self.lib.dbi_mg_expand(self.g._h, tile_begin, n_tiles, self.C.byref(v)))
        return v.value

    def index_variants(self, key, payload):
        torch.cuda.current_stream().synchronize()
        self._ck(self.lib.dbi_mg_index_variants(self.g._h, self._p(key), self._p(payload), int(key.numel())))
