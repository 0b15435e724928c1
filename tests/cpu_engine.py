"""A CPU ShardEngine backed by the oracle and numpy -- TEST INFRASTRUCTURE.  Lets the multi-rank
orchestration of dbindex_b200/multigpu.py (splitters, all-to-all, the (gpos, len) all-gather, group
exchange under global peptide ids, query routing) run under gloo without a GPU.  It shards by PROTEIN ranges (the GPU
engine shards by start-position tiles; both concatenate to the global emission order)."""
from __future__ import annotations

import numpy as np
import torch

import dbindex_b200 as dbi
from dbindex_b200.multigpu import MG_BINS, ShardEngine
from oracle.oracle_py import Oracle

from . import pyref


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


class OracleShardEngine(ShardEngine):
    def __init__(self, params, residues, offsets):
        self.params = params
        self.residues = np.ascontiguousarray(residues, np.uint8)
        self.offsets = np.ascontiguousarray(offsets, np.uint64)
        self.has_mods = params.n_mods > 0 and params.max_mods_per_peptide > 0
        self.min_mass = float(params.min_mass)
        self.nomod = params.copy()
        self.nomod.n_mods = 0
        self.nomod.max_mods_per_peptide = 0
        self.base_bits = int(np.float64(params.min_mass).view(np.uint64))
        span = int(np.float64(params.max_mass).view(np.uint64)) - self.base_bits
        self.nbits = span.bit_length()
        self.shift = max(0, self.nbits - 12)
        self.pstart = (self.offsets[:-1] + np.arange(len(self.offsets) - 1, dtype=np.uint64) + 1).astype(np.int64)

    # ---- helpers
    def _gpos(self, prot, off):
        return (self.pstart[prot.astype(np.int64)] + off.astype(np.int64)).astype(np.uint32)

    def _prot_off(self, gpos, prot):
        return (gpos.astype(np.int64) - self.pstart[prot.astype(np.int64)]).astype(np.uint32)

    def _seq(self, prot, off, ln):
        a = int(self.offsets[prot]) + int(off)
        return self.residues[a:a + int(ln)].tobytes().decode()

    # ---- stages
    def begin(self, rank, world):
        self.rank, self.world = rank, world

    def digest(self):
        P = len(self.offsets) - 1
        p0, p1 = P * self.rank // self.world, P * (self.rank + 1) // self.world
        o = Oracle(self.nomod)
        sub = self.offsets[p0:p1 + 1] - self.offsets[p0]
        o.add_proteins(self.residues[int(self.offsets[p0]):int(self.offsets[p1])], sub)
        assert o.build() == 0
        e = o.emitted()
        prot = e["prot"] + np.uint32(p0)
        self.rec = [_bits(e["mass"]).copy(), self._gpos(prot, e["off"]), prot, e["len"].copy()]
        return len(prot)

    def _keys(self, stage):
        return (self.rec[0] - np.uint64(self.base_bits)) if stage == 0 else self.var[0]

    def histogram(self, stage):
        bins = (self._keys(stage) >> np.uint64(self.shift)).astype(np.int64)
        return torch.from_numpy(np.bincount(bins, minlength=MG_BINS).astype(np.int64)), self.shift

    def partition(self, stage, splitters):
        thr = splitters.astype(np.uint64) << np.uint64(self.shift)
        dest = np.searchsorted(thr, self._keys(stage), side="right")
        self.perm = np.argsort(dest, kind="stable")
        return np.bincount(dest, minlength=self.world).astype(np.uint64)

    def pack_send(self, stage):
        src = self.rec if stage == 0 else self.var
        dt = [np.int64, np.int32, np.int32, np.int16] if stage == 0 else [np.int64, np.int64]
        return [torch.from_numpy(a[self.perm].view(d).copy()) for a, d in zip(src, dt)]

    def index_base(self, mass, gpos, prot, length):
        mass = mass.numpy().view(np.float64)
        gpos, prot = gpos.numpy().view(np.uint32), prot.numpy().view(np.uint32)
        length = length.numpy().view(np.uint16)
        o = Oracle(self.nomod)
        o.add_proteins(self.residues, self.offsets)
        assert o.build_from_records(mass, prot, self._prot_off(gpos, prot), length) == 0
        e = o.entries()
        self.local = [_bits(e["mass"]).view(np.int64), self._gpos(e["first_prot"], e["first_off"]).view(np.int32),
                      e["first_prot"].view(np.int32), e["len"].view(np.int16),
                      np.diff(e["prot_list_off"].astype(np.int64)).astype(np.int32), e["prot_ids"].view(np.int32)]

    def n_unique(self):
        return len(self.local[0])

    def export_unique(self):
        return [torch.from_numpy(np.ascontiguousarray(self.local[1])), torch.from_numpy(np.ascontiguousarray(self.local[3]))]

    def import_unique(self, rank_unique, tables):
        # all a rank learns about foreign peptides: where their residues are
        self.g_gpos, self.g_len = tables[0].numpy().view(np.uint32), tables[1].numpy().view(np.uint16)
        self.uoff = int(sum(rank_unique[:self.rank]))
        assert int(rank_unique[self.rank]) == self.n_unique()

    def finish(self):
        self.e_mass = self.local[0].view(np.float64)
        self.e_base = None  # entry i = own unique peptide i
        self.e_pat = np.zeros(self.n_unique(), np.uint32)

    def _variants(self, u):
        pr = int(self.local[2].view(np.uint32)[u])
        off = int(self.local[1].view(np.uint32)[u]) - int(self.pstart[pr])
        return pyref.expand_set(self.params, self._seq(pr, off, self.local[3].view(np.uint16)[u]),
                                float(self.local[0].view(np.float64)[u]))

    def own_tiles(self):
        return 0, (self.n_unique() + 255) // 256

    def expand(self, tile_begin, n_tiles):
        U = self.n_unique()
        keys, pay = [], []
        for u in range(tile_begin * 256, min(U, (tile_begin + n_tiles) * 256)):
            for m, pos in self._variants(u):
                pat = sum((q + 1) << (8 * k) for k, q in enumerate(pos))
                keys.append(int(np.float64(m).view(np.uint64)) - self.base_bits)
                pay.append(((self.uoff + u) << 32) | pat)  # peptides travel under their global id
        self.var = [np.array(keys, dtype=np.uint64), np.array(pay, dtype=np.uint64)]
        return len(keys)

    def index_variants(self, key, payload):
        key, payload = key.numpy().view(np.uint64), payload.numpy().view(np.uint64)
        order = np.argsort(key, kind="stable")
        self.e_mass = (key[order] + np.uint64(self.base_bits)).view(np.float64)
        self.e_base = (payload[order] >> np.uint64(32)).astype(np.uint32)
        self.e_pat = (payload[order] & np.uint64(0xFFFFFFFF)).astype(np.uint32)

    # ---- what a query sees (COLLECTIVE: base peptides of other ranks are looked up in their tables)
    def entries(self):
        import torch.distributed as dist
        tabs = [None] * self.world
        dist.all_gather_object(tabs, [np.asarray(a) for a in self.local])
        u_gpos = np.concatenate([t[1].view(np.uint32) for t in tabs])
        u_prot = np.concatenate([t[2].view(np.uint32) for t in tabs])
        u_len = np.concatenate([t[3].view(np.uint16) for t in tabs])
        u_plo = np.concatenate(([0], np.cumsum(np.concatenate([t[4].astype(np.int64) for t in tabs]))))
        plist = np.concatenate([t[5].view(np.uint32) for t in tabs])
        my_off = sum(len(t[0]) for t in tabs[:self.rank])
        b = (np.arange(len(self.e_mass)) + my_off) if self.e_base is None else self.e_base.astype(np.int64)
        return {"mass": self.e_mass, "first_prot": u_prot[b], "len": u_len[b], "modpat": self.e_pat,
                "first_off": self._prot_off(u_gpos[b], u_prot[b]),
                "plist": [tuple(plist[u_plo[x]:u_plo[x + 1]].tolist()) for x in b]}

    def query(self, lo, hi):
        b = np.searchsorted(self.e_mass, lo, side="left")
        e = np.searchsorted(self.e_mass, hi, side="right")
        return b, np.maximum(e - b, 0)
