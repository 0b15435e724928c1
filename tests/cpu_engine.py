"""A CPU ShardEngine backed by the oracle and numpy -- TEST INFRASTRUCTURE.  Lets the multi-rank
orchestration of dbindex_b200/multigpu.py (shard layout, histograms -> plan -> count matrix, the two
exchanges, unique offsets, query routing) run under gloo without a GPU.  Windows do not exist here:
`scatter` delivers with a gloo all-to-all and the descriptors are dummies."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from dbindex_b200.multigpu import DESC_BYTES, MG_BINS, ShardEngine, shard_proteins, slice_owner
from oracle.oracle_py import Oracle

from . import pyref


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


class OracleShardEngine(ShardEngine):
    def __init__(self, params, residues, offsets):
        """residues / offsets: the WHOLE proteome; begin() keeps only this rank's shard (like a rank that
        parsed its own part of the FASTA) until pull_proteome() gathers the others."""
        self.params = params
        self._all = (np.ascontiguousarray(residues, np.uint8), np.ascontiguousarray(offsets, np.uint64))
        self.has_mods = params.n_mods > 0 and params.max_mods_per_peptide > 0
        self.min_mass = float(params.min_mass)
        self.nomod = params.copy()
        self.nomod.n_mods = 0
        self.nomod.max_mods_per_peptide = 0
        self.base_bits = int(np.float64(params.min_mass).view(np.uint64))
        span = int(np.float64(params.max_mass).view(np.uint64)) - self.base_bits
        self.nbits = span.bit_length()
        self.shift = max(0, self.nbits - 12)
        self.rank_unique = None
        self._nu = 0

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
        self.shard_res, self.shard_off, self.p0 = shard_proteins(self._all[0], self._all[1], rank, world)
        self._all = None  # from here on a rank only knows its own shard

    def shard_info(self):
        return len(self.shard_off) - 1, int(self.shard_off[-1])

    def set_shards(self, shard_proteins_, shard_residues):
        assert int(shard_proteins_[self.rank]) == len(self.shard_off) - 1
        self.sp = np.asarray(shard_proteins_, dtype=np.int64)
        return np.zeros(DESC_BYTES, np.uint8)

    def window(self, window, nbytes):
        return np.zeros(DESC_BYTES, np.uint8)

    def import_window(self, window, rank, desc):
        pass

    def layout_bytes(self, window, stage, n_items):
        return 0

    def group_bytes(self):
        return 16

    def pull_proteome(self):
        parts = [None] * self.world
        dist.all_gather_object(parts, (self.shard_res, self.shard_off))
        assert [len(o) - 1 for _, o in parts] == self.sp.tolist()
        self.residues = np.concatenate([r for r, _ in parts])
        lens = np.concatenate([np.diff(o.astype(np.int64)) for _, o in parts])
        self.offsets = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
        self.pstart = (self.offsets[:-1] + np.arange(len(self.offsets) - 1, dtype=np.uint64) + 1).astype(np.int64)

    def digest(self):
        o = Oracle(self.nomod)
        o.add_proteins(self.shard_res, self.shard_off)
        assert o.build() == 0
        e = o.emitted()
        prot = e["prot"] + np.uint32(self.p0)
        self.rec = [_bits(e["mass"]).copy(), self._gpos(prot, e["off"]), prot, e["len"].copy()]
        return len(prot)

    def _keys(self, stage):
        return (self.rec[0] - np.uint64(self.base_bits)) if stage == 0 else self.var[0]

    def hist(self, stage):
        bins = (self._keys(stage) >> np.uint64(self.shift)).astype(np.int64)
        plain = np.bincount(bins, minlength=MG_BINS).astype(np.int64)
        return torch.from_numpy(np.concatenate([plain, plain, np.zeros_like(plain)])), self.shift  # every item weighs 1 here, no group estimates

    def _dest(self, stage, splitters, n_slices):
        thr = np.asarray(splitters, dtype=np.uint64) << np.uint64(self.shift)
        return slice_owner(np.searchsorted(thr, self._keys(stage), side="right"), n_slices, self.world)

    def count(self, stage, splitters, n_slices):
        dest = self._dest(stage, splitters, n_slices)
        return np.bincount(dest, minlength=self.world).astype(np.uint64)

    def scatter(self, stage, splitters, n_slices, matrix):
        dest = self._dest(stage, splitters, n_slices)
        perm = np.argsort(dest, kind="stable")
        send = np.bincount(dest, minlength=self.world)
        assert send.tolist() == matrix[self.rank].astype(np.int64).tolist(), "count matrix row differs from the partition"
        recv = matrix[:, self.rank].astype(np.int64)
        src = self.rec if stage == 0 else self.var
        if stage == 1:  # peptides travel under their global id
            src = [src[0], src[1] + (np.uint64(self.uoff) << np.uint64(32))]
        out = []
        for a in src:
            raw = np.ascontiguousarray(a[perm]).view(np.uint8)
            w = a.dtype.itemsize
            got = torch.empty(int(recv.sum()) * w, dtype=torch.uint8)
            if self.world == 1:
                got.copy_(torch.from_numpy(raw.copy()))
            else:
                dist.all_to_all_single(got, torch.from_numpy(raw.copy()), output_split_sizes=[int(c) * w for c in recv],
                                       input_split_sizes=[int(c) * w for c in send])
            out.append(got.numpy().view(a.dtype))
        self.arena = out

    def index_base(self):
        mass, gpos, prot, length = self.arena
        o = Oracle(self.nomod)
        o.add_proteins(self.residues, self.offsets)
        assert o.build_from_records(mass.view(np.float64), prot, self._prot_off(gpos, prot), length) == 0
        e = o.entries()
        self.local = [_bits(e["mass"]).view(np.int64), self._gpos(e["first_prot"], e["first_off"]).view(np.int32),
                      e["first_prot"].view(np.int32), e["len"].view(np.int16),
                      np.diff(e["prot_list_off"].astype(np.int64)).astype(np.int32), e["prot_ids"].view(np.int32)]
        self._nu = len(self.local[0])

    def n_unique(self):
        return self._nu

    def set_unique(self, rank_unique):
        self.rank_unique = np.asarray(rank_unique, dtype=np.uint64)
        assert int(rank_unique[self.rank]) == self._nu
        self.uoff = int(sum(int(x) for x in rank_unique[:self.rank]))

    def finish(self):
        if not self.has_mods:
            self.e_mass = self.local[0].view(np.float64)
            self.e_base = None  # entry i = own unique peptide i
            self.e_pat = np.zeros(self._nu, np.uint32)

    def _variants(self, u):
        pr = int(self.local[2].view(np.uint32)[u])
        off = int(self.local[1].view(np.uint32)[u]) - int(self.pstart[pr])
        return pyref.expand_set(self.params, self._seq(pr, off, self.local[3].view(np.uint16)[u]),
                                float(self.local[0].view(np.float64)[u]))

    def groups(self):
        keys, pay = [], []
        for u in range(self._nu):
            for m, pos in self._variants(u):
                pat = sum((q + 1) << (8 * k) for k, q in enumerate(pos))
                keys.append(int(np.float64(m).view(np.uint64)) - self.base_bits)
                pay.append((u << 32) | pat)  # LOCAL row; the exchange adds the rank's offset
        self.var = [np.array(keys, dtype=np.uint64), np.array(pay, dtype=np.uint64)]
        return len(keys), len(keys)

    def index_variants(self):
        key, payload = self.arena
        order = np.argsort(key, kind="stable")
        self.e_mass = (key[order] + np.uint64(self.base_bits)).view(np.float64)
        self.e_base = (payload[order] >> np.uint64(32)).astype(np.uint32)
        self.e_pat = (payload[order] & np.uint64(0xFFFFFFFF)).astype(np.uint32)

    # ---- what a query sees (COLLECTIVE: base peptides of other ranks are looked up in their tables)
    def entries(self):
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
