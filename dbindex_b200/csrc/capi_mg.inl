// capi_mg.inl -- sharded (multi-GPU) build behind the C ABI; included at the end of capi.cu.
//
// The reference shards its index by mass: DBIndexStoreSQLiteMult keeps `indexFactor` SQLite files,
// one per mass bucket (DBIndexStoreSQLiteMult.java:55-56,215-217) and a query walks the buckets its
// range touches (:333-343).  Here a bucket is a GPU with one contiguous slice of the mass axis:
//
//   every rank adds ITS OWN shard of the FASTA (dbi_add_proteins), packs it into its place of the
//   global residue buffer and pulls the other shards over NVLink (window 0); digests its share of the
//   start positions; exchange 0 moves the records to the owners of their base-mass slices, exchange 1
//   moves the variant groups (with their site masks) to the owners of their variant-mass slices.
//   Both exchanges are ONE kernel each: a stable multisplit whose output pointers are the other
//   GPUs' arenas (mapped peer memory), mg.cu.
//
// Small collectives (two 32 KB histograms, a world x world count matrix, barriers) are the caller's:
// torch.distributed / NCCL between processes (dbindex_b200/multigpu.py), plain host code when one
// process holds every handle (dbi_mg_build_local below, what a Java host calls).

#include <unistd.h>

#include <array>

namespace {

constexpr int kWinProt = 0, kWinArena = 1, kWinUniq = 2;

// Exportable allocations and peer mappings outlive handles: a later handle of the same process gets
// the same blocks back, so peers keep their mappings (and cudaIpcOpenMemHandle is paid once).
struct WinBlock {
  void* p;
  uint64_t cap;
};
struct WinCaches {
  std::mutex mu;
  std::vector<WinBlock> free_[64][3];
  std::map<std::array<uint8_t, 64>, void*> opened;  // IPC handle -> mapped base
};
WinCaches& win_caches() {
  static WinCaches c;
  return c;
}

void* win_get(int device, int window, uint64_t bytes, uint64_t* cap) {
  WinCaches& c = win_caches();
  {
    std::lock_guard<std::mutex> lk(c.mu);
    auto& v = c.free_[device & 63][window];
    int best = -1;
    for (int i = 0; i < (int)v.size(); ++i)
      if (v[i].cap >= bytes && (best < 0 || v[i].cap < v[best].cap)) best = i;
    if (best >= 0) {
      const WinBlock b = v[best];
      v.erase(v.begin() + best);
      *cap = b.cap;
      return b.p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    DevCache::of(device).trim();
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) throw CudaError{e, "cudaMalloc (window)", __FILE__, __LINE__};
  *cap = bytes;
  return p;
}

// the largest cached block of a window, if any (no allocation): a fresh handle adopts what an earlier
// handle of this process left, so its peers keep their mappings and nothing has to grow again
void* win_adopt(int device, int window, uint64_t* cap) {
  WinCaches& c = win_caches();
  std::lock_guard<std::mutex> lk(c.mu);
  auto& v = c.free_[device & 63][window];
  int best = -1;
  for (int i = 0; i < (int)v.size(); ++i)
    if (best < 0 || v[i].cap > v[best].cap) best = i;
  if (best < 0) return nullptr;
  const WinBlock b = v[best];
  v.erase(v.begin() + best);
  *cap = b.cap;
  return b.p;
}

void win_put(int device, int window, void* p, uint64_t cap) {
  if (!p) return;
  WinCaches& c = win_caches();
  std::lock_guard<std::mutex> lk(c.mu);
  c.free_[device & 63][window].push_back({p, cap});
}

void mg_release(dbi_handle* h) {
  for (int w = 0; w < 3; ++w) {
    win_put(h->device, w, h->win[w].p, h->win[w].cap);
    h->win[w] = dbi_handle::MgWindow();
  }
}

// arena layouts: what an exchange delivers to a rank that receives n items
struct RecLayout {
  uint64_t mass, gpos, prot, len, total;
  explicit RecLayout(uint64_t n) {
    mass = 0;
    gpos = mass + al256(8 * n);
    prot = gpos + al256(4 * n);
    len = prot + al256(4 * n);
    total = len + al256(2 * n);
  }
};
struct GrpLayout {
  uint64_t key, pay, gid, mask, total;
  GrpLayout(uint64_t n, int C) {
    key = 0;
    pay = key + al256(8 * n);
    gid = pay + al256(8 * n);
    mask = gid + (C > 0 ? al256(4 * n) : 0);
    total = mask + (C > 0 ? al256(8 * n * (uint64_t)C) : 0);
  }
};

// classes travelling with a group record: the group path ships C masks, the per-variant path none
int mg_side_classes(const dbi_handle* h) { return h->cfg.n_seq > 0 ? h->cfg.n_classes : 0; }

int mg_shift(const KeySpace& ks) { return ks.nbits > 12 ? ks.nbits - 12 : 0; }

// fills the thresholds of a plan from bin splitters; false (error set) if they are malformed
bool mg_fill_plan(MgPlan& pl, int W, int n_slices, const uint32_t* bin_splitters, int sh) {
  std::memset(&pl, 0, sizeof(pl));
  if ((n_slices != W && n_slices != 2 * W) || (n_slices > 1 && !bin_splitters)) {
    set_error("n_slices must be world or 2 * world (%d for world %d)", n_slices, W);
    return false;
  }
  pl.world = W;
  pl.n_thr = n_slices - 1;
  for (int d = 0; d + 1 < n_slices; ++d) {
    if (d > 0 && bin_splitters[d] < bin_splitters[d - 1]) {
      set_error("bin_splitters must be ascending");
      return false;
    }
    pl.thr[d] = (uint64_t)bin_splitters[d] << sh;
  }
  return true;
}

// the main stream waits for the shards pulled on the side stream
void mg_join_pull(dbi_handle* h) {
  if (!h->pull_pending) return;
  DBI_CUDA(cudaStreamWaitEvent(h->stream, h->pull_done, 0));
  h->pull_pending = false;
}

}  // namespace

extern "C" {

int dbi_mg_begin(dbi_handle* h, int rank, int world) {
  DBI_API_BEGIN(h)
  if (h->built) {
    set_error("index already built");
    return DBI_EALREADY;
  }
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
    set_error("bad rank %d / world %d (1..%d)", rank, world, kMaxRanks);
    return DBI_EINVAL;
  }
  h->mg_rank = rank;
  h->mg_world = world;
  return DBI_OK;
  DBI_API_END
}

uint64_t dbi_mg_layout_bytes(int window, int stage, uint64_t n_items, int n_classes) {
  if (window == kWinUniq) return UniqLayout(n_items).total;
  if (window == kWinArena) return stage == 0 ? RecLayout(n_items).total : GrpLayout(n_items, n_classes).total;
  return 0;
}

int dbi_mg_side_classes(dbi_handle* h) { return h ? mg_side_classes(h) : 0; }

int dbi_mg_window_ensure(dbi_handle* h, int window, uint64_t bytes, dbi_mg_window* desc) {
  DBI_API_BEGIN(h)
  if (window < 0 || window > 2) {
    set_error("bad window %d", window);
    return DBI_EINVAL;
  }
  dbi_handle::MgWindow& w = h->win[window];
  if (!w.p) w.p = win_adopt(h->device, window, &w.cap);
  if (bytes > w.cap) {
    if (window == kWinProt && h->mg_layout) {
      set_error("window 0 cannot grow once the shards are laid out");
      return DBI_EINVAL;
    }
    DBI_CUDA(cudaStreamSynchronize(h->stream));
    win_put(h->device, window, w.p, w.cap);  // the old block stays allocated: peers may still map it
    w.p = nullptr;
    w.cap = 0;
    w.p = win_get(h->device, window, bytes + bytes / 4 + 4096, &w.cap);
  }
  w.peer[h->mg_rank] = w.p;
  w.peer_cap[h->mg_rank] = w.cap;
  if (desc) {
    std::memset(desc, 0, sizeof(*desc));
    if (w.p) {
      cudaIpcMemHandle_t ih;
      DBI_CUDA(cudaIpcGetMemHandle(&ih, w.p));
      static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
      std::memcpy(desc->ipc, &ih, 64);
    }
    desc->ptr = (uint64_t)(uintptr_t)w.p;
    desc->bytes = w.cap;
    desc->device = h->device;
    desc->pid = (int32_t)getpid();
  }
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_window_import(dbi_handle* h, int window, int rank, const dbi_mg_window* desc) {
  DBI_API_BEGIN(h)
  if (window < 0 || window > 2 || rank < 0 || rank >= h->mg_world || !desc) {
    set_error("bad argument");
    return DBI_EINVAL;
  }
  dbi_handle::MgWindow& w = h->win[window];
  if (rank == h->mg_rank) return DBI_OK;
  void* base = nullptr;
  if (desc->ptr == 0) {
    base = nullptr;
  } else if (desc->pid == (int32_t)getpid()) {
    // same process: the pointer itself, after enabling peer access between the two devices
    if (desc->device != h->device) {
      int can = 0;
      DBI_CUDA(cudaDeviceCanAccessPeer(&can, h->device, desc->device));
      if (!can) {
        set_error("device %d cannot access device %d (no P2P path)", h->device, desc->device);
        return DBI_ECUDA;
      }
      const cudaError_t e = cudaDeviceEnablePeerAccess(desc->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) throw CudaError{e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__};
      cudaGetLastError();
    }
    base = (void*)(uintptr_t)desc->ptr;
  } else {
    WinCaches& c = win_caches();
    std::array<uint8_t, 64> key;
    std::memcpy(key.data(), desc->ipc, 64);
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.opened.find(key);
    if (it != c.opened.end()) {
      base = it->second;
    } else {
      cudaIpcMemHandle_t ih;
      std::memcpy(&ih, desc->ipc, 64);
      DBI_CUDA(cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess));
      c.opened.emplace(key, base);
    }
  }
  w.peer[rank] = base;
  w.peer_cap[rank] = desc->bytes;
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_set_shards(dbi_handle* h, const uint64_t* shard_proteins, const uint64_t* shard_residues,
                      uint64_t* window0_bytes) {
  DBI_API_BEGIN(h)
  if (h->built) {
    set_error("index already built");
    return DBI_EALREADY;
  }
  if (!shard_proteins || !shard_residues) {
    set_error("null argument");
    return DBI_EINVAL;
  }
  const int W = h->mg_world, r = h->mg_rank;
  const uint64_t my_prot = h->h_off.size() - 1;
  if (shard_proteins[r] != my_prot || shard_residues[r] != h->n_res) {
    set_error("shard %d is declared as %llu proteins / %llu residues but this rank holds %llu / %llu", r,
              (unsigned long long)shard_proteins[r], (unsigned long long)shard_residues[r],
              (unsigned long long)my_prot, (unsigned long long)h->n_res);
    return DBI_EINVAL;
  }
  h->prot_off[0] = 0;
  h->pos_off[0] = 0;
  for (int s = 0; s < W; ++s) {
    h->prot_off[s + 1] = h->prot_off[s] + shard_proteins[s];
    h->pos_off[s + 1] = h->pos_off[s] + shard_residues[s] + shard_proteins[s];
  }
  const uint64_t P = h->prot_off[W], res_end = h->pos_off[W] + 1;
  if (res_end + 4096 >= (1ull << 32) || P >= (1ull << 31)) {
    set_error("more than 2^32 residues in the sharded proteome");
    return DBI_ERANGE;
  }
  const uint64_t padded = (res_end + 64 + 15) & ~15ull;
  const uint64_t pstart_at = al256(padded);
  const uint64_t total = pstart_at + (P + 1) * 4;
  if (window0_bytes) {  // sizing call: the caller ensures window 0 and calls again with NULL
    *window0_bytes = total;
    return DBI_OK;
  }
  dbi_handle::MgWindow& w = h->win[kWinProt];
  if (w.cap < total) {
    set_error("window 0 holds %llu bytes, the proteome needs %llu", (unsigned long long)w.cap, (unsigned long long)total);
    return DBI_EINVAL;
  }
  cudaStream_t st = h->stream;
  ensure_uploaded(h);
  const uint32_t zero = 0;
  DBI_CUDA(cudaMemcpyAsync(h->d_err.p, &zero, 4, cudaMemcpyHostToDevice, st));
  h->res_end = (uint32_t)res_end;
  h->res_alloc = padded;
  h->d_res.borrow(w.p, padded);
  h->d_pstart.borrow((uint8_t*)w.p + pstart_at, (P + 1) * 4);
  if (r == W - 1) DBI_CUDA(cudaMemsetAsync((uint8_t*)w.p + res_end, 0, padded - res_end, st));
  {
    Stage sg(h, DBI_STAGE_PACK);
    launch_pack(h->d_raw.as<uint8_t>(), h->d_off.as<uint64_t>(), (uint32_t)my_prot, h->n_res,
                h->d_res.as<uint8_t>() + h->pos_off[r], h->d_pstart.as<uint32_t>() + h->prot_off[r],
                (uint32_t)h->pos_off[r], h->d_err.as<uint32_t>(), st);
    h->st.algo_bytes[DBI_STAGE_PACK] += 2 * h->n_res + my_prot * 12;
  }
  DBI_CUDA(cudaStreamSynchronize(st));
  h->mg_layout = true;
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_pull_proteome(dbi_handle* h) {
  DBI_API_BEGIN(h)
  if (!h->mg_layout) {
    set_error("dbi_mg_set_shards first");
    return DBI_ENOTINIT;
  }
  const int W = h->mg_world, r = h->mg_rank;
  // The copies run on the handle's side stream, concurrently with the digest of the own shard (which reads
  // nothing else); whatever reads foreign residues afterwards waits for pull_done (mg_join_pull).
  if (!h->side_stream) {
    DBI_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    DBI_CUDA(cudaEventCreateWithFlags(&h->pull_done, cudaEventDisableTiming));
  }
  cudaStream_t st = h->side_stream;
  // the own shard is packed (main stream, synchronised by dbi_mg_set_shards); the last rank also zeroed the tail
  const uint64_t pstart_at = (uint64_t)((uint8_t*)h->d_pstart.p - (uint8_t*)h->d_res.p);
  for (int k = 1; k < W; ++k) {
    const int s = (r + k) % W;  // start with the next rank: the pulls of the ranks spread over the links
    const uint8_t* peer = (const uint8_t*)h->win[kWinProt].peer[s];
    if (!peer) {
      set_error("window 0 of rank %d is not mapped", s);
      return DBI_EINVAL;
    }
    // shard s: its separator, residues and inner separators; the last shard also owns the final separator
    const uint64_t b0 = h->pos_off[s], b1 = s == W - 1 ? (uint64_t)h->res_alloc : h->pos_off[s + 1];
    const uint64_t p0 = h->prot_off[s], p1 = h->prot_off[s + 1] + (s == W - 1 ? 1 : 0);
    if (b1 > b0)
      DBI_CUDA(cudaMemcpyAsync(h->d_res.as<uint8_t>() + b0, peer + b0, b1 - b0, cudaMemcpyDefault, st));
    if (p1 > p0)
      DBI_CUDA(cudaMemcpyAsync((uint8_t*)h->d_pstart.p + p0 * 4, peer + pstart_at + p0 * 4, (p1 - p0) * 4,
                               cudaMemcpyDefault, st));
    h->st.algo_bytes[DBI_STAGE_PACK] += 2 * (b1 - b0) + 8 * (p1 - p0);
  }
  DBI_CUDA(cudaEventRecord(h->pull_done, st));
  h->pull_pending = true;
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_digest(dbi_handle* h, uint64_t* n_records) {
  DBI_API_BEGIN(h)
  if (h->built) {
    set_error("index already built");
    return DBI_EALREADY;
  }
  if (!h->mg_layout) {
    set_error("dbi_mg_set_shards first");
    return DBI_ENOTINIT;
  }
  cudaStream_t s = h->stream;
  // Every rank digests the start positions of ITS OWN shard: contiguous, ascending ranges, so rank order ==
  // global emission order, and nothing outside the shard is read for the result -- the other shards may
  // still be arriving (dbi_mg_pull_proteome runs on the side stream).
  const int W = h->mg_world, r = h->mg_rank;
  const uint64_t lo_pos = h->pos_off[r], hi_pos = r == W - 1 ? (uint64_t)h->res_end : h->pos_off[r + 1];
  const DigestRange rg{(uint32_t)lo_pos, (uint32_t)hi_pos, (uint32_t)h->prot_off[r], (uint32_t)h->prot_off[r + 1]};
  const uint32_t t0 = (uint32_t)(lo_pos / kDigestTile);
  const uint32_t t1 = (uint32_t)((hi_pos + kDigestTile - 1) / kDigestTile);
  const uint32_t nt = hi_pos > lo_pos ? t1 - t0 : 0u;
  DevBuf tile_counts, tile_offs, start_cnt;
  tile_counts.alloc((uint64_t)nt * 4, h->arena);
  tile_offs.alloc(((uint64_t)nt + 1) * 8, h->arena);
  start_cnt.alloc(std::max<uint64_t>(1, nt) * kDigestTile, h->arena);
  uint64_t N = 0;
  {
    Stage sg(h, DBI_STAGE_DIGEST_COUNT);
    launch_digest_count(h->d_res.as<uint8_t>(), rg, h->res_alloc, h->d_tables.as<DevTables>(), h->cfg, t0, nt,
                        start_cnt.as<uint8_t>(), tile_counts.as<uint32_t>(), h->d_err.as<uint32_t>(), s);
    launch_scan_u32_to_u64(tile_counts.as<uint32_t>(), nt, tile_offs.as<uint64_t>(), s);
    N = read_u64(h, tile_offs.as<uint64_t>() + nt);
    h->st.algo_bytes[DBI_STAGE_DIGEST_COUNT] += 2ull * nt * kDigestTile + (uint64_t)nt * 16;
  }
  if (int rc = check_err_bits(read_err(h))) {
    free_index(h);
    return rc;
  }
  if (N >= (1ull << 32)) {
    set_error("more than 2^32 emitted records on one GPU");
    return DBI_ERANGE;
  }
  h->mg_mass.alloc(N * 8, h->arena);
  h->mg_gpos.alloc(N * 4, h->arena);
  h->mg_prot.alloc(N * 4, h->arena);
  h->mg_len.alloc(N * 2, h->arena);
  if (h->cfg.max_mods > 0) h->mg_nmod.alloc(std::max<uint64_t>(N, 1), h->arena);
  {
    Stage sg(h, DBI_STAGE_DIGEST_EMIT);
    launch_digest_emit(h->d_res.as<uint8_t>(), rg, h->res_alloc, h->d_tables.as<DevTables>(), h->cfg, t0, nt,
                       start_cnt.as<uint8_t>(), tile_offs.as<uint64_t>(), h->d_pstart.as<uint32_t>(),
                       h->mg_mass.as<uint64_t>(), h->mg_gpos.as<uint32_t>(), h->mg_prot.as<uint32_t>(),
                       h->mg_len.as<uint16_t>(), h->mg_nmod.as<uint8_t>(), h->d_err.as<uint32_t>(), s);
    h->st.algo_bytes[DBI_STAGE_DIGEST_EMIT] += 2ull * nt * kDigestTile + N * 18;
  }
  DBI_CUDA(cudaStreamSynchronize(s));
  h->mg_n = N;
  h->n_emitted = N;
  h->st.n_emitted = N;
  if (n_records) *n_records = N;
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_hist(dbi_handle* h, int stage, void* d_hist, int* shift) {
  DBI_API_BEGIN(h)
  if (!d_hist || (stage != 0 && stage != 1)) {
    set_error("bad argument");
    return DBI_EINVAL;
  }
  const KeySpace ks(h->p.min_mass, h->p.max_mass);
  const int sh = mg_shift(ks);
  if (stage == 0) {
    // weighted = index entries the records will expand to, estimated from their mod-site counts n: sum over
    // k <= K of C(n, k); groups = variant groups they will list: sum over k <= min(n, K) of min(C^k, C(n, k));
    // plain = records.  The cuts of exchange 0 are the cuts of the whole index, so they must anticipate the
    // expansion.
    const uint32_t* wtab = nullptr;
    const uint32_t* gtab = nullptr;
    if (h->cfg.max_mods > 0 && h->mg_nmod.p) {
      if (!h->mg_wtab.p) {
        uint32_t tab[512];
        const int K = h->cfg.max_mods, C = std::max(1, h->cfg.n_classes);
        for (int n = 0; n < 256; ++n) {
          uint64_t t = 0, g = 0, c = 1, ck = 1;  // C(n, 0), C^0
          for (int k = 0; k <= K; ++k) {
            t += c;
            g += std::min<uint64_t>(c, ck);
            c = c * (uint64_t)(n - k) / (uint64_t)(k + 1);
            ck = std::min<uint64_t>(ck * (uint64_t)C, 1u << 20);
            if (n - k <= 0) break;
          }
          tab[n] = (uint32_t)std::min<uint64_t>(t, 0xffffull);  // < 2^16: stays in the 32-bit shared bins
          tab[256 + n] = (uint32_t)std::min<uint64_t>(g, h->cfg.n_seq > 0 ? (uint64_t)h->cfg.n_seq : 0xffull);
        }
        h->mg_wtab.alloc(sizeof(tab), h->arena);
        DBI_CUDA(cudaMemcpyAsync(h->mg_wtab.p, tab, sizeof(tab), cudaMemcpyHostToDevice, h->stream));
        DBI_CUDA(cudaStreamSynchronize(h->stream));
      }
      wtab = h->mg_wtab.as<uint32_t>();
      gtab = wtab + 256;
    }
    launch_mg_hist(h->mg_mass.as<uint64_t>(), h->mg_n, ks.base_bits, sh, nullptr, 0, 0u, h->mg_nmod.as<uint8_t>(), wtab,
                   gtab, (unsigned long long*)d_hist, h->stream);
  } else {  // group records: weigh by their variant count
    const bool weighted = h->cfg.n_seq > 0;
    // weighted = index entries the groups stand for, plain = groups; dbi_mg_plan turns both into a cost
    launch_mg_hist(h->mg_vkey.as<uint64_t>(), h->mg_v, 0, sh, weighted ? h->mg_vpay.as<uint64_t>() : nullptr,
                   kGrpCntMask, 0u, nullptr, nullptr, nullptr, (unsigned long long*)d_hist, h->stream);
  }
  if (shift) *shift = sh;
  DBI_CUDA(cudaStreamSynchronize(h->stream));  // the caller reduces d_hist on ITS stream next
  return DBI_OK;
  DBI_API_END
}

// Pure host arithmetic: the cuts of the mass axis from the global histograms [weighted | plain | groups]
// (bins are never split, so equal masses -- hence equal peptides -- meet on one rank), this rank's send
// counts from its own plain histogram, every rank's receive total from the global plain histogram.
//
// n_slices = world: slice s lives on rank s.  n_slices = 2 * world: FOLDED slices, slice s lives on rank
// s < world ? s : 2 * world - 1 - s -- every rank holds one light and one heavy slice.  With differential
// mods the light end of the axis is crowded with records (base phase) and the heavy end with variants
// (variant and search phases); contiguous slices can balance the sum of the phases but not each of them,
// and a sharded build lasts as long as the slowest rank of EVERY phase (the exchanges are barriers).
//
// Per bin:
//   base phase    B = cost[0] * items                      (sort + merge of the records / of the items)
//   variant phase V = cost[1] * groups + cost[2] * weight  (group sort; expansion writes per index entry)
//   search phase  Q = cost[3] * weight^2 * mass / (width * total weight)
//     = expected HITS of precursor queries that follow the indexed mass density at a relative (ppm) tolerance:
//     queries landing in the bin ~ weight / total, hits per query ~ (weight / width) * mass.  An index is built
//     once and searched many times, so the slices are cut for the search load too.
// The cuts start where the SUM B + V + Q is equal per slice; when more than one phase has a cost they are then
// moved, one at a time, to wherever  max_r B + max_r V + max_r Q  gets smallest (coordinate descent on the
// prefix sums: a few thousand operations).  cost == NULL: {0, 0, 1, 0} (equal weight).  Every rank runs this on
// the same global histogram and gets the same splitters.
int dbi_mg_plan(int world, const uint64_t* hist_global, const uint64_t* hist_local, int shift, double min_mass,
                const double* cost, int n_slices, uint32_t* bin_splitters, uint64_t* send_counts,
                uint64_t* recv_totals) {
  if (world < 1 || world > kMaxRanks || !hist_global || !hist_local || !send_counts || !recv_totals ||
      (n_slices != world && n_slices != 2 * world) || (n_slices > 1 && !bin_splitters) || shift < 0 || shift > 52) {
    set_error("bad argument");
    return DBI_EINVAL;
  }
  const int B = kMgBins, S = n_slices, W = world;
  const double c_item = cost ? cost[0] : 0.0, c_grp = cost ? cost[1] : 0.0, c_w = cost ? cost[2] : 1.0,
               c_hit = cost ? cost[3] : 0.0;
  double w_total = 0;
  for (int b = 0; b < B; ++b) w_total += (double)hist_global[b];
  const uint64_t base_bits = dbits(min_mass);
  auto mass_at = [&](int b) {  // mass at the lower edge of bin b
    const uint64_t bits = base_bits + ((uint64_t)b << shift);
    double m;
    std::memcpy(&m, &bits, 8);
    return m;
  };
  // prefix sums of the three phase costs: P[ph][b] = cost of bins [0, b)
  std::vector<double> P[3];
  for (auto& v : P) v.assign(B + 1, 0.0);
  for (int b = 0; b < B; ++b) {
    const double items = (double)hist_global[B + b], w = (double)hist_global[b], g = (double)hist_global[2 * B + b];
    double q = 0;
    if (c_hit > 0 && w > 0 && w_total > 0) {
      const double m0 = mass_at(b), m1 = mass_at(b + 1);
      const double width = m1 - m0;
      if (width > 0) q = c_hit * w * w * (0.5 * (m0 + m1)) / (width * w_total);
    }
    P[0][b + 1] = P[0][b] + c_item * items;
    P[1][b + 1] = P[1][b] + c_grp * g + c_w * w;
    P[2][b + 1] = P[2][b] + q;
  }
  const double total = P[0][B] + P[1][B] + P[2][B];
  std::vector<int> cut(S + 1, 0);  // slice s = bins [cut[s], cut[s + 1])
  cut[S] = B;
  {
    int b = 0;
    for (int d = 1; d < S; ++d) {
      const double target = total * (double)d / (double)S;
      // first bin boundary at which at least `target` of the cost lies below
      while (b < B && (P[0][b + 1] + P[1][b + 1] + P[2][b + 1]) < target) ++b;
      int c = total > 0 ? std::min(b + 1, B) : 0;
      if (c < cut[d - 1]) c = cut[d - 1];
      cut[d] = c;
    }
  }
  auto owner = [&](int sl) { return (int)mg_slice_owner((uint32_t)sl, (uint32_t)S, (uint32_t)W); };
  const int phases = (P[0][B] > 0) + (P[1][B] > 0) + (P[2][B] > 0);
  const char* no_refine = std::getenv("DBI_MG_REFINE");  // "0": keep the equal-sum cuts (diagnostic)
  if (S > 1 && phases > 1 && !(no_refine && no_refine[0] == '0')) {
    double load[3][kMaxRanks];  // per phase and rank, under the current cuts
    auto slice_cost = [&](int ph, int sl) { return P[ph][cut[sl + 1]] - P[ph][cut[sl]]; };
    auto reload = [&]() {
      for (int ph = 0; ph < 3; ++ph) {
        for (int r = 0; r < W; ++r) load[ph][r] = 0;
        for (int sl = 0; sl < S; ++sl) load[ph][owner(sl)] += slice_cost(ph, sl);
      }
    };
    auto objective = [&]() {
      double o = 0;
      for (int ph = 0; ph < 3; ++ph) {
        double mx = 0;
        for (int r = 0; r < W; ++r) mx = std::max(mx, load[ph][r]);
        o += mx;
      }
      return o;
    };
    // moving cut d shifts cost between slices d - 1 and d only
    auto move_cut = [&](int d, int c) {
      const int ra = owner(d - 1), rb = owner(d);
      for (int ph = 0; ph < 3; ++ph) {
        const double delta = P[ph][c] - P[ph][cut[d]];
        load[ph][ra] += delta;
        load[ph][rb] -= delta;
      }
      cut[d] = c;
    };
    reload();
    double best = objective();
    for (int sweep = 0; sweep < 5; ++sweep) {
      bool moved = false;
      for (int d = 1; d < S; ++d) {
        const int lo = cut[d - 1], hi = cut[d + 1];
        int keep = cut[d];
        // coarse grid over the whole gap, then two refinements around the best candidate
        int span = hi - lo, centre = keep;
        for (int level = 0; level < 3 && span > 0; ++level) {
          const int steps = 16;
          const int a = level == 0 ? lo : std::max(lo, centre - span), z = level == 0 ? hi : std::min(hi, centre + span);
          for (int i = 0; i <= steps; ++i) {
            move_cut(d, a + (int)((int64_t)(z - a) * i / steps));
            const double o = objective();
            if (o < best * (1.0 - 1e-12)) {
              best = o;
              keep = cut[d];
              moved = true;
            }
          }
          centre = keep;
          span = std::max(1, (z - a) / steps);
        }
        move_cut(d, keep);
      }
      reload();  // drop the rounding drift of the incremental updates
      if (!moved) break;
    }
  }
  for (int d = 1; d < S; ++d) bin_splitters[d - 1] = (uint32_t)cut[d];
  for (int d = 0; d < W; ++d) send_counts[d] = recv_totals[d] = 0;
  for (int sl = 0; sl < S; ++sl) {
    uint64_t sc = 0, rt = 0;
    for (int x = cut[sl]; x < cut[sl + 1]; ++x) {
      sc += hist_local[B + x];
      rt += hist_global[B + x];
    }
    send_counts[owner(sl)] += sc;
    recv_totals[owner(sl)] += rt;
  }
  return DBI_OK;
}

// dbi_mg_plan for a rank that holds EVERY rank's local histograms (one all-gather instead of an all-reduce plus
// a second exchange of the send counts): hist_all[world][3 * DBI_MG_BINS].  Sums the global histogram, places
// the cuts, and fills the whole count matrix: matrix[src * world + dst] = items rank src sends to rank dst.
int dbi_mg_plan_matrix(int world, const uint64_t* hist_all, int shift, double min_mass, const double* cost,
                       int n_slices, uint32_t* bin_splitters, uint64_t* matrix) {
  if (world < 1 || world > kMaxRanks || !hist_all || !matrix) {
    set_error("bad argument");
    return DBI_EINVAL;
  }
  const size_t HB = 3 * (size_t)kMgBins;
  std::vector<uint64_t> global(HB, 0);
  for (int r = 0; r < world; ++r)
    for (size_t i = 0; i < HB; ++i) global[i] += hist_all[(size_t)r * HB + i];
  uint64_t send[kMaxRanks], recv[kMaxRanks];
  uint32_t split_local[kMaxSlices];
  uint32_t* split = bin_splitters ? bin_splitters : split_local;
  if (int rc = dbi_mg_plan(world, global.data(), hist_all, shift, min_mass, cost, n_slices, split, send, recv)) return rc;
  for (int src = 0; src < world; ++src) {
    uint64_t* row = matrix + (size_t)src * world;
    for (int d = 0; d < world; ++d) row[d] = 0;
    const uint64_t* plain = hist_all + (size_t)src * HB + kMgBins;
    int lo = 0;
    for (int sl = 0; sl < n_slices; ++sl) {
      const int hi = sl + 1 < n_slices ? (int)split[sl] : kMgBins;
      uint64_t c = 0;
      for (int x = lo; x < hi; ++x) c += plain[x];
      row[mg_slice_owner((uint32_t)sl, (uint32_t)n_slices, (uint32_t)world)] += c;
      lo = hi;
    }
  }
  return DBI_OK;
}

// The cost model of an exchange for dbi_mg_plan {per item, per estimated group, per unit of weight (index
// entry), per expected hit}: measured on B200 (profiles/r02_*), in ps.  DBI_MG_COST0 / DBI_MG_COST =
// "item,group,weight,hit" override exchange 0's / exchange 1's.
void dbi_mg_default_cost(int stage, int has_mods, double* cost) {
  if (stage == 0 && !has_mods) {
    cost[0] = 1.0; cost[1] = 0.0; cost[2] = 0.0; cost[3] = 0.0;  // records are the entries: equal counts
    return;
  }
  if (stage == 0) {
    // The cuts of exchange 0 are the cuts of the whole index (the variant groups travel by the same cuts).
    // per record: hash + 11 radix passes + merge + group count (sort_base 0.63 + dedup 0.14 + mod_count 0.29 ms
    // per 2.7 M); per group: listing + 7 radix passes + the scatter (mod_emit 0.22 + sort_var 1.04 + 0.4 ms per
    // 13 M); per entry: the expansion write (1.21 ms per 69 M); per expected hit: 13 ps x 10 000 queries x 2e-5
    // (+-10 ppm) x ~8 (masses cluster inside a bin).
    cost[0] = 400.0; cost[1] = 125.0; cost[2] = 17.5; cost[3] = 20.0;
  } else {
    cost[0] = 125.0; cost[1] = 0.0; cost[2] = 17.5; cost[3] = 20.0;  // exchange 1 alone (only when it plans its own cuts)
  }
  if (const char* e = std::getenv(stage == 0 ? "DBI_MG_COST0" : "DBI_MG_COST")) {
    double a, g2, b2, c2;
    if (std::sscanf(e, "%lf,%lf,%lf,%lf", &a, &g2, &b2, &c2) == 4 && a >= 0 && g2 >= 0 && b2 >= 0 && c2 >= 0 &&
        a + g2 + b2 + c2 > 0) {
      cost[0] = a; cost[1] = g2; cost[2] = b2; cost[3] = c2;
    }
  }
}

int dbi_mg_count(dbi_handle* h, int stage, const uint32_t* bin_splitters, int n_slices, uint64_t* send_counts) {
  DBI_API_BEGIN(h)
  const int W = h->mg_world;
  if (!send_counts || (stage != 0 && stage != 1)) {
    set_error("bad argument");
    return DBI_EINVAL;
  }
  const KeySpace ks(h->p.min_mass, h->p.max_mass);
  const int sh = mg_shift(ks);
  MgPlan pl;
  if (!mg_fill_plan(pl, W, n_slices, bin_splitters, sh)) return DBI_EINVAL;
  DevBuf cnt;
  cnt.alloc(kMaxRanks * 8, h->arena);
  DBI_CUDA(cudaMemsetAsync(cnt.p, 0, kMaxRanks * 8, h->stream));
  if (stage == 0)
    launch_mg_count(h->mg_mass.as<uint64_t>(), h->mg_n, ks.base_bits, pl, (unsigned long long*)cnt.p, h->stream);
  else
    launch_mg_count(h->mg_vkey.as<uint64_t>(), h->mg_v, 0, pl, (unsigned long long*)cnt.p, h->stream);
  uint64_t host[kMaxRanks];
  DBI_CUDA(cudaMemcpyAsync(host, cnt.p, kMaxRanks * 8, cudaMemcpyDeviceToHost, h->stream));
  DBI_CUDA(cudaStreamSynchronize(h->stream));
  for (int d = 0; d < W; ++d) send_counts[d] = host[d];
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_scatter(dbi_handle* h, int stage, const uint32_t* bin_splitters, int n_slices, const uint64_t* matrix) {
  DBI_API_BEGIN(h)
  const int W = h->mg_world, r = h->mg_rank;
  if (!matrix || (stage != 0 && stage != 1)) {
    set_error("bad argument");
    return DBI_EINVAL;
  }
  cudaStream_t s = h->stream;
  const KeySpace ks(h->p.min_mass, h->p.max_mass);
  const int sh = mg_shift(ks);
  const int C = mg_side_classes(h);
  MgPlan pl;
  if (!mg_fill_plan(pl, W, n_slices, bin_splitters, sh)) return DBI_EINVAL;
  h->mg_nthr[stage] = pl.n_thr;
  for (int d = 0; d < pl.n_thr; ++d) h->mg_thr[stage][d] = pl.thr[d];
  uint64_t recv[kMaxRanks] = {}, sent = 0;
  for (int d = 0; d < W; ++d) {
    for (int src = 0; src < W; ++src) {
      if (src < r) pl.row0[d] += matrix[src * W + d];
      recv[d] += matrix[src * W + d];
    }
    sent += matrix[r * W + d];
  }
  const uint64_t n_local = stage == 0 ? h->mg_n : h->mg_v;
  if (sent != n_local) {
    set_error("count matrix row %d sums to %llu but this rank holds %llu items", r, (unsigned long long)sent,
              (unsigned long long)n_local);
    return DBI_EINVAL;
  }
  for (int d = 0; d < W; ++d) {
    if (recv[d] >= (1ull << 32)) {
      set_error("more than 2^32 items for rank %d", d);
      return DBI_ERANGE;
    }
    const uint64_t need = stage == 0 ? RecLayout(recv[d]).total : GrpLayout(recv[d], C).total;
    if (recv[d] && (!h->win[kWinArena].peer[d] || h->win[kWinArena].peer_cap[d] < need)) {
      set_error("arena of rank %d is not mapped or too small (%llu < %llu bytes)", d,
                (unsigned long long)h->win[kWinArena].peer_cap[d], (unsigned long long)need);
      return DBI_EINVAL;
    }
  }
  DevBuf tmp;
  tmp.alloc(mg_scatter_tmp_bytes(n_local), h->arena);
  Stage sg(h, DBI_STAGE_OTHER);
  if (stage == 0) {
    MgRecDst dst;
    std::memset(&dst, 0, sizeof(dst));
    for (int d = 0; d < W; ++d) {
      uint8_t* base = (uint8_t*)h->win[kWinArena].peer[d];
      const RecLayout L(recv[d]);
      dst.mass[d] = (uint64_t*)(base + L.mass);
      dst.gpos[d] = (uint32_t*)(base + L.gpos);
      dst.prot[d] = (uint32_t*)(base + L.prot);
      dst.len[d] = (uint16_t*)(base + L.len);
      h->uniq_cap[d] = recv[d];  // window 2 of every rank is laid out by what exchange 0 delivers to it
    }
    launch_mg_scatter_records(h->mg_mass.as<uint64_t>(), h->mg_gpos.as<uint32_t>(), h->mg_prot.as<uint32_t>(),
                              h->mg_len.as<uint16_t>(), n_local, ks.base_bits, pl, dst, tmp.p, s);
    h->st.algo_bytes[DBI_STAGE_OTHER] += n_local * 2 * 18;
    h->mg_mass.release(); h->mg_gpos.release(); h->mg_prot.release(); h->mg_len.release();
    h->mg_nmod.release();
  } else {
    MgGrpDst dst;
    std::memset(&dst, 0, sizeof(dst));
    for (int d = 0; d < W; ++d) {
      uint8_t* base = (uint8_t*)h->win[kWinArena].peer[d];
      const GrpLayout L(recv[d], C);
      dst.key[d] = (uint64_t*)(base + L.key);
      dst.pay[d] = (uint64_t*)(base + L.pay);
      dst.gid[d] = (uint32_t*)(base + L.gid);
      dst.mask[d] = (uint64_t*)(base + L.mask);
    }
    if (C > 0) ensure_site_masks(h);
    launch_mg_scatter_groups(h->mg_vkey.as<uint64_t>(), h->mg_vpay.as<uint64_t>(), h->u_cmask.as<uint64_t>(), C,
                             h->uoff[r], n_local, pl, dst, tmp.p, s);
    h->st.algo_bytes[DBI_STAGE_OTHER] += n_local * 2 * (20 + 8ull * C);
    h->mg_vkey.release(); h->mg_vpay.release();
  }
  h->mg_recv[stage] = recv[r];
  DBI_CUDA(cudaStreamSynchronize(s));  // the peers may read their arenas once every rank is past this point
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_index_base(dbi_handle* h) {
  DBI_API_BEGIN(h)
  mg_join_pull(h);  // the records delivered here name residues of every shard
  const uint64_t n = h->mg_recv[0];
  const KeySpace ks(h->p.min_mass, h->p.max_mass);
  const RecLayout L(n);
  const uint8_t* base = (const uint8_t*)h->win[kWinArena].p;
  const RecView rv{(const uint64_t*)(base + L.mass), (const uint32_t*)(base + L.gpos), (const uint32_t*)(base + L.prot),
                   (const uint16_t*)(base + L.len)};
  int rc = sort_dedup(h, rv, n, ks);
  if (rc == DBI_OK) rc = check_err_bits(read_err(h));
  if (rc == DBI_OK && n == 0) {  // empty slice: still needs (empty) tables
    h->u_mass.alloc(8, h->arena);
    if (!h->u_gpos.p) { h->u_gpos.alloc(8, h->arena); h->u_prot.alloc(8, h->arena); h->u_len.alloc(8, h->arena); h->plist.alloc(8, h->arena); }
  }
  DBI_CUDA(cudaStreamSynchronize(h->stream));
  return rc;
  DBI_API_END
}

int dbi_mg_unique_count(dbi_handle* h, uint64_t* n_unique) {
  if (!h || !n_unique) {
    set_error("null argument");
    return DBI_EINVAL;
  }
  *n_unique = h->n_unique;
  return DBI_OK;
}

int dbi_mg_set_unique(dbi_handle* h, const uint64_t* rank_unique) {
  DBI_API_BEGIN(h)
  if (!rank_unique) {
    set_error("null argument");
    return DBI_EINVAL;
  }
  if (rank_unique[h->mg_rank] != h->n_unique) {
    set_error("rank_unique[%d] = %llu but this rank holds %llu unique peptides", h->mg_rank,
              (unsigned long long)rank_unique[h->mg_rank], (unsigned long long)h->n_unique);
    return DBI_EINVAL;
  }
  h->uoff[0] = 0;
  for (int r = 0; r < h->mg_world; ++r) h->uoff[r + 1] = h->uoff[r] + rank_unique[r];
  if (h->uoff[h->mg_world] >= (1ull << 32)) {
    set_error("more than 2^32 unique peptides in total");
    return DBI_ERANGE;
  }
  h->ent_base_off = h->uoff[h->mg_rank];
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_groups(dbi_handle* h, uint64_t* n_items, uint64_t* n_variants) {
  DBI_API_BEGIN(h)
  const KeySpace ks(h->p.min_mass, h->p.max_mass);
  const uint32_t utiles = (uint32_t)((h->n_unique + kModTile - 1) / kModTile);
  uint64_t V = 0, NG = 0;
  int rc;
  const uint64_t keep = h->ent_base_off;
  h->ent_base_off = 0;  // the records name LOCAL peptide rows; the exchange turns them into global ids
  if (h->cfg.n_seq > 0) {
    rc = emit_groups(h, 0, utiles, ks, h->mg_vkey, h->mg_vpay, &NG, &V);
  } else {
    rc = emit_variants(h, 0, utiles, ks, h->mg_vkey, h->mg_vpay, &V);
    NG = V;
  }
  h->ent_base_off = keep;
  DBI_CUDA(cudaStreamSynchronize(h->stream));
  h->mg_v = NG;  // items that travel: group records, or variants on the per-variant path
  if (n_items) *n_items = NG;
  if (n_variants) *n_variants = V;
  return rc;
  DBI_API_END
}

int dbi_mg_index_variants(dbi_handle* h) {
  DBI_API_BEGIN(h)
  TR("begin");
  const uint64_t n = h->mg_recv[1];
  const KeySpace ks(h->p.min_mass, h->p.max_mass);
  const int C = mg_side_classes(h);
  const GrpLayout L(n, C);
  uint8_t* base = (uint8_t*)h->win[kWinArena].p;
  int rc;
  if (h->cfg.n_seq > 0) {
    const UniqView uv = uniq_view(h);
    GroupSide side;
    side.cmask = (const uint64_t*)(base + L.mask);
    side.gid = (const uint32_t*)(base + L.gid);
    side.uv = &uv;
    rc = sort_expand_groups(h, (uint64_t*)(base + L.key), (uint64_t*)(base + L.pay), n, ks, &side);
  } else {
    rc = sort_variants(h, (uint64_t*)(base + L.key), (uint64_t*)(base + L.pay), n, ks);
  }
  if (rc == DBI_OK) rc = check_err_bits(read_err(h));
  DBI_CUDA(cudaStreamSynchronize(h->stream));
  TR("final_sync");
  g_trace.dump();
  return rc;
  DBI_API_END
}

int dbi_mg_finish(dbi_handle* h) {
  DBI_API_BEGIN(h)
  if (h->cfg.max_mods == 0) {
    h->e_mass.release(); h->e_base.release(); h->e_pat.release();
    h->n_entries = h->n_unique;  // the entries of this rank are its own unique peptides
    h->st.n_entries = h->n_entries;
  }
  h->built = true;
  finish_stats(h);
  return DBI_OK;
  DBI_API_END
}

int dbi_mg_slices(dbi_handle* h) {
  if (!h) return 0;
  const int stage = h->cfg.max_mods > 0 ? 1 : 0;
  return h->mg_world > 1 ? h->mg_nthr[stage] + 1 : 1;
}

int dbi_mg_split_masses(dbi_handle* h, double* split_mass) {
  if (!h || (h->mg_world > 1 && !split_mass)) {
    set_error("null argument");
    return DBI_EINVAL;
  }
  const KeySpace ks(h->p.min_mass, h->p.max_mass);
  const int stage = h->cfg.max_mods > 0 ? 1 : 0;
  if (h->mg_world > 1)
    for (int d = 0; d < h->mg_nthr[stage]; ++d) {
      const uint64_t bits = h->mg_thr[stage][d] + ks.base_bits;
      std::memcpy(&split_mass[d], &bits, 8);
    }
  return DBI_OK;
}

// The whole sharded build when ONE process holds a handle per GPU (what a Java host calls): the same
// stages as above with the small collectives done by the host.  handles[r] becomes rank r; every
// handle must have been given ITS shard of the FASTA (dbi_add_proteins) in rank order.
int dbi_mg_build_local(dbi_handle** hs, int n) {
  if (!hs || n < 1 || n > kMaxRanks) {
    set_error("bad argument");
    return DBI_EINVAL;
  }
  for (int r = 0; r < n; ++r)
    if (!hs[r]) {
      set_error("null handle");
      return DBI_EINVAL;
    }
#define MG_ALL(expr)                         \
  for (int r = 0; r < n; ++r) {              \
    dbi_handle* h = hs[r];                   \
    (void)h;                                 \
    if (int rc_ = (expr)) return rc_;        \
  }
  try {
    const int W = n;
    std::vector<uint64_t> sp(W), sr(W);
    for (int r = 0; r < W; ++r) {
      sp[r] = hs[r]->h_off.size() - 1;
      sr[r] = hs[r]->n_res;
    }
    MG_ALL(dbi_mg_begin(h, r, W));
    auto connect = [&](int window) -> int {  // every rank maps every rank's window
      std::vector<dbi_mg_window> d(W);
      for (int r = 0; r < W; ++r)
        if (int rc = dbi_mg_window_ensure(hs[r], window, 0, &d[r])) return rc;
      for (int r = 0; r < W; ++r)
        for (int q = 0; q < W; ++q)
          if (int rc = dbi_mg_window_import(hs[r], window, q, &d[q])) return rc;
      return DBI_OK;
    };
    {
      uint64_t bytes = 0;
      MG_ALL(dbi_mg_set_shards(h, sp.data(), sr.data(), &bytes));
      MG_ALL(dbi_mg_window_ensure(h, kWinProt, bytes, nullptr));
      MG_ALL(dbi_mg_set_shards(h, sp.data(), sr.data(), nullptr));
      if (int rc = connect(kWinProt)) return rc;
      MG_ALL(dbi_mg_pull_proteome(h));
    }
    MG_ALL(dbi_mg_digest(h, nullptr));
    const bool mods = hs[0]->cfg.max_mods > 0;
    const int C = mg_side_classes(hs[0]);
    std::vector<uint64_t> matrix((size_t)W * W), recv(W);
    // folded slices when differential mods make the phases of a build pull the cuts apart
    const int S = (mods && W > 1) ? 2 * W : W;
    std::vector<uint32_t> split((size_t)S);
    // exchange 0 plans the cuts (histograms -> equal-cost splitters); exchange 1 reuses them, so the variant
    // groups of a peptide mostly stay on the GPU that owns the peptide: only the counts are needed
    auto exchange = [&](int stage) -> int {
      if (stage == 0) {
        const size_t HB = 3 * kMgBins;
        std::vector<uint64_t> local((size_t)W * HB), global(HB, 0);
        int shift_of = 0;
        for (int r = 0; r < W; ++r) {
          dbi_handle* h = hs[r];
          DBI_CUDA(cudaSetDevice(h->device));
          DevBuf d;
          d.alloc(HB * 8, h->arena);
          DBI_CUDA(cudaMemsetAsync(d.p, 0, HB * 8, h->stream));
          if (int rc = dbi_mg_hist(h, stage, d.p, &shift_of)) return rc;
          DBI_CUDA(cudaMemcpyAsync(&local[(size_t)r * HB], d.p, HB * 8, cudaMemcpyDeviceToHost, h->stream));
          DBI_CUDA(cudaStreamSynchronize(h->stream));
        }
        for (int r = 0; r < W; ++r)
          for (size_t i = 0; i < HB; ++i) global[i] += local[(size_t)r * HB + i];
        double cost[4];
        dbi_mg_default_cost(stage, mods ? 1 : 0, cost);
        for (int r = 0; r < W; ++r)
          if (int rc = dbi_mg_plan(W, global.data(), &local[(size_t)r * HB], shift_of, hs[0]->p.min_mass, cost, S,
                                   split.data(), &matrix[(size_t)r * W], recv.data()))
            return rc;
      } else {
        for (int r = 0; r < W; ++r)
          if (int rc = dbi_mg_count(hs[r], stage, split.data(), S, &matrix[(size_t)r * W])) return rc;
        for (int d = 0; d < W; ++d) {
          recv[d] = 0;
          for (int r = 0; r < W; ++r) recv[d] += matrix[(size_t)r * W + d];
        }
      }
      for (int r = 0; r < W; ++r) {
        const uint64_t need = dbi_mg_layout_bytes(kWinArena, stage, recv[r], C);
        if (int rc = dbi_mg_window_ensure(hs[r], kWinArena, need, nullptr)) return rc;
        if (stage == 0)
          if (int rc = dbi_mg_window_ensure(hs[r], kWinUniq, dbi_mg_layout_bytes(kWinUniq, 0, recv[r], 0), nullptr)) return rc;
      }
      if (int rc = connect(kWinArena)) return rc;
      if (stage == 0)
        if (int rc = connect(kWinUniq)) return rc;
      MG_ALL(dbi_mg_scatter(h, stage, split.data(), S, matrix.data()));
      return DBI_OK;
    };
    if (int rc = exchange(0)) return rc;
    MG_ALL(dbi_mg_index_base(h));
    std::vector<uint64_t> ru(W);
    for (int r = 0; r < W; ++r) ru[r] = hs[r]->n_unique;
    MG_ALL(dbi_mg_set_unique(h, ru.data()));
    if (mods) {
      MG_ALL(dbi_mg_groups(h, nullptr, nullptr));
      if (int rc = exchange(1)) return rc;
      MG_ALL(dbi_mg_index_variants(h));
    }
    MG_ALL(dbi_mg_finish(h));
    return DBI_OK;
  } catch (const CudaError& e) {
    return fail_cuda(e);
  } catch (const std::bad_alloc&) {
    set_error("host allocation failed");
    return DBI_ENOMEM;
  }
#undef MG_ALL
}

}  // extern "C"
