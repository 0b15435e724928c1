// mg.cu -- device helpers of the multi-GPU exchange.
//
// The reference already shards its index by mass: DBIndexStoreSQLiteMult splits the mass axis
// into `indexFactor` equal-width buckets, one SQLite file each (DBIndexStoreSQLiteMult.java:55-56,
// 215-217) and answers a query from the buckets its range touches (:333-343).  Here a bucket is
// a GPU slice of the mass axis, the slice edges are taken from key histograms (the mass density is
// far from uniform; dbi_mg_plan balances every phase of the build), and records travel to their
// bucket inside the partition kernel itself: peer stores over NVLink into the owners' arenas.
#include "kernels.cuh"

namespace dbi {
namespace {

constexpr int MG_THREADS = 256;

// ---- histograms of an exchange stage ------------------------------------------------------
// hist[0 .. kMgBins)            += weight of the item  (stage 1: variant count of the group; stage 0: the
//                                  index entries a record will expand to, estimated from its mod sites; else 1)
// hist[kMgBins .. 2 kMgBins)    += 1
// hist[2 kMgBins .. 3 kMgBins)  += variant GROUPS a record will list (stage 0 with mods, estimated; else 0)
// over the bins min(kMgBins - 1, (key - sub) >> shift).  The plain one gives every rank the number of
// items it will send and receive; all three feed the cost model that places the cuts (dbi_mg_plan).
// 32-bit shared-memory bins (native atomics); a CTA sees at most 2^16 items and weights of 2^16 or more
// (a group can stand for 2^27 variants) go straight to the 64-bit global bins, so nothing can wrap.
constexpr uint32_t kMgMaxItemsPerCta = 1u << 16;

// destination rank of a key: its slice (thresholds ascending), folded onto the ranks
__device__ __forceinline__ uint32_t mg_dest(const MgPlan& pl, uint64_t k) {
  uint32_t sl = 0;
  for (int x = 0; x < pl.n_thr; ++x) sl += (pl.thr[x] <= k) ? 1u : 0u;
  return mg_slice_owner(sl, (uint32_t)pl.n_thr + 1u, (uint32_t)pl.world);
}

__global__ void __launch_bounds__(MG_THREADS)
    mg_hist_kernel(const uint64_t* __restrict__ key, uint64_t n, uint64_t per_cta, uint64_t sub, int shift,
                   const uint64_t* __restrict__ wpay, uint64_t wmask, uint32_t wadd,
                   const uint8_t* __restrict__ wcode, const uint32_t* __restrict__ wtab,
                   const uint32_t* __restrict__ gtab, unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh_w[kMgBins];
  __shared__ uint32_t sh_c[kMgBins];
  __shared__ uint32_t sh_g[kMgBins];  // 3 x 16 KB: exactly the 48 KB of static shared memory
  for (int i = threadIdx.x; i < kMgBins; i += MG_THREADS) {
    sh_w[i] = 0;
    sh_c[i] = 0;
    sh_g[i] = 0;
  }
  __syncthreads();
  const uint64_t i0 = (uint64_t)blockIdx.x * per_cta, i1 = min(n, i0 + per_cta);
  for (uint64_t i = i0 + threadIdx.x; i < i1; i += MG_THREADS) {
    uint64_t b = (key[i] - sub) >> shift;
    if (b >= kMgBins) b = kMgBins - 1;
    atomicAdd(&sh_c[b], 1u);
    const uint32_t code = wcode ? wcode[i] : 0u;
    const uint64_t w = wpay ? (wpay[i] & wmask) + wadd : (wtab ? (uint64_t)__ldg(wtab + code) : 1ull);
    if (w < (1ull << 16)) atomicAdd(&sh_w[b], (uint32_t)w);
    else atomicAdd(&hist[b], (unsigned long long)w);
    if (gtab) atomicAdd(&sh_g[b], __ldg(gtab + code));  // < 2^8 per record
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMgBins; i += MG_THREADS) {
    if (sh_w[i]) atomicAdd(&hist[i], (unsigned long long)sh_w[i]);
    if (sh_c[i]) atomicAdd(&hist[kMgBins + i], (unsigned long long)sh_c[i]);
    if (sh_g[i]) atomicAdd(&hist[2 * kMgBins + i], (unsigned long long)sh_g[i]);
  }
}

// items per destination for given thresholds (the variant exchange reuses the cuts of the record exchange,
// so it needs no histogram: only how many groups cross each cut)
__global__ void __launch_bounds__(MG_THREADS)
    mg_count_kernel(const uint64_t* __restrict__ key, uint64_t n, uint64_t sub, const __grid_constant__ MgPlan pl,
                    unsigned long long* __restrict__ counts) {
  __shared__ uint32_t sh[kMaxRanks];
  if (threadIdx.x < kMaxRanks) sh[threadIdx.x] = 0;
  __syncthreads();
  uint32_t mine[kMaxRanks];
#pragma unroll
  for (int d = 0; d < kMaxRanks; ++d) mine[d] = 0;
  const uint64_t stride = (uint64_t)gridDim.x * MG_THREADS;
  for (uint64_t i = (uint64_t)blockIdx.x * MG_THREADS + threadIdx.x; i < n; i += stride) {
    const uint32_t d = mg_dest(pl, key[i] - sub);
#pragma unroll
    for (int x = 0; x < kMaxRanks; ++x) mine[x] += (d == (uint32_t)x) ? 1u : 0u;
  }
#pragma unroll
  for (int d = 0; d < kMaxRanks; ++d)
    if (mine[d]) atomicAdd(&sh[d], mine[d]);
  __syncthreads();
  if (threadIdx.x < kMaxRanks && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// ---- fused multisplit + all-to-all over peer memory ------------------------------------------
// One pass over the local items of an exchange stage: every item finds its destination rank
// (thresholds on the radix key), gets its arrival row there -- stable: items keep their local
// order per destination, and rank s's items follow rank s-1's, so the receive side is in global
// emission order (SURVEY.md Q6 needs "first occurrence" to stay global) -- and is written STRAIGHT
// into the destination GPU's arena through NVLink peer stores (mapped peer pointers), staged through
// shared memory so that each destination receives contiguous, coalesced runs.  No pack buffer, no
// NCCL all-to-all: the partition kernel is the transport.
//
// Cross-tile offsets come from a decoupled look-back over the tiles of THIS GPU only (tiles are
// claimed in launch order); there is no waiting on another GPU inside the kernel.
constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_IPT = 8;
constexpr int SC_TILE = SC_THREADS * SC_IPT;  // 2048 items per tile

struct ScSmem {
  uint64_t stage[SC_TILE];                 // one array at a time, in destination order
  uint32_t wcnt[SC_WARPS][kMaxRanks];      // per-warp counts, then exclusive prefixes over the warps
  uint32_t dstart[kMaxRanks + 1];          // first staging slot of each destination
  uint64_t row0[kMaxRanks];                // arrival row of staging slot dstart[d]
  uint8_t sdest[SC_TILE];                  // destination of each staging slot
  uint32_t tile;
};

// look-back word: [63:62] flag, [61:0] count
constexpr uint64_t SLB_AGG = 1ull << 62;
constexpr uint64_t SLB_INC = 2ull << 62;
constexpr uint64_t SLB_MASK = (1ull << 62) - 1;

struct ScCtx {
  uint32_t tile, tile_n;
  uint64_t base;             // first local item of the tile
  uint32_t slot[SC_IPT];     // staging slot of this thread's items
  bool valid[SC_IPT];
};

// Ranks the tile's items by destination; leaves s.dstart / s.row0 / s.sdest and the slots in cx.
__device__ __forceinline__ void sc_rank_tile(ScSmem& s, ScCtx& cx, const uint64_t* __restrict__ key, uint64_t n,
                                             uint64_t sub, const MgPlan& pl, unsigned long long* lookback,
                                             uint32_t* tile_counter) {
  const int t = threadIdx.x, w = t >> 5;
  const unsigned l = lane_id();
  if (t == 0) s.tile = atomicAdd(tile_counter, 1u);
  if (t < SC_WARPS * kMaxRanks) (&s.wcnt[0][0])[t] = 0;
  __syncthreads();
  cx.tile = s.tile;
  cx.base = (uint64_t)cx.tile * SC_TILE;
  cx.tile_n = (uint32_t)min((uint64_t)SC_TILE, n - cx.base);
  uint32_t dest[SC_IPT], rank[SC_IPT];
#pragma unroll
  for (int i = 0; i < SC_IPT; ++i) {  // warp-striped: element e = w * 32 * IPT + i * 32 + lane
    const uint32_t e = (uint32_t)w * 32 * SC_IPT + i * 32 + l;
    cx.valid[i] = e < cx.tile_n;
    uint32_t d = 0;
    if (cx.valid[i]) {
      d = mg_dest(pl, key[cx.base + e] - sub);
    } else {
      d = (uint32_t)pl.world;  // filler: after every real destination
    }
    dest[i] = d;
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    uint32_t old = 0;
    if ((int)l == leader && d < (uint32_t)pl.world) {
      old = s.wcnt[w][d];
      s.wcnt[w][d] = old + (uint32_t)__popc(m);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[i] = old + (uint32_t)__popc(m & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();
  // thread d < world owns destination d: prefix over the warps, tile count, look-back
  if (t < pl.world) {
    uint32_t sum = 0;
    for (int ww = 0; ww < SC_WARPS; ++ww) {
      const uint32_t c = s.wcnt[ww][t];
      s.wcnt[ww][t] = sum;
      sum += c;
    }
    s.dstart[t] = sum;  // count for now
    volatile unsigned long long* lb = lookback;
    uint64_t excl = 0;
    if (cx.tile == 0) {
      lb[(size_t)cx.tile * kMaxRanks + t] = SLB_INC | (uint64_t)sum;
    } else {
      lb[(size_t)cx.tile * kMaxRanks + t] = SLB_AGG | (uint64_t)sum;
      int64_t j = (int64_t)cx.tile - 1;
      while (true) {
        const uint64_t st = lb[(size_t)j * kMaxRanks + t];
        if ((st >> 62) == 0) continue;  // not published yet
        excl += st & SLB_MASK;
        if ((st >> 62) == 2) break;
        --j;
      }
      lb[(size_t)cx.tile * kMaxRanks + t] = SLB_INC | (excl + sum);
    }
    s.row0[t] = pl.row0[t] + excl;
  }
  __syncthreads();
  if (t == 0) {  // counts -> exclusive starts
    uint32_t run = 0;
    for (int d = 0; d < pl.world; ++d) {
      const uint32_t c = s.dstart[d];
      s.dstart[d] = run;
      run += c;
    }
    s.dstart[pl.world] = run;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SC_IPT; ++i) {
    if (cx.valid[i]) {
      cx.slot[i] = s.dstart[dest[i]] + s.wcnt[w][dest[i]] + rank[i];
      s.sdest[cx.slot[i]] = (uint8_t)dest[i];
    }
  }
  __syncthreads();
}

// Moves one array: `load(e)` = value of local item e of the tile, `dst[d]` = the array at destination d.
// Row r of destination d lands at dst[d][r * stride + offset] (interleaved side tables).
template <typename T, typename L>
__device__ __forceinline__ void sc_move(ScSmem& s, const ScCtx& cx, L&& load, T* const* dst, uint32_t stride = 1,
                                        uint32_t offset = 0) {
  const int t = threadIdx.x, w = t >> 5;
  const unsigned l = lane_id();
  T* st = reinterpret_cast<T*>(s.stage);
#pragma unroll
  for (int i = 0; i < SC_IPT; ++i)
    if (cx.valid[i]) st[cx.slot[i]] = load((uint32_t)w * 32 * SC_IPT + i * 32 + l);
  __syncthreads();
  for (uint32_t j = t; j < cx.tile_n; j += SC_THREADS) {
    const uint32_t d = s.sdest[j];
    dst[d][(s.row0[d] + (j - s.dstart[d])) * stride + offset] = st[j];
  }
  __syncthreads();
}

// stage 0: digested records -> the owners of their mass slices
__global__ void __launch_bounds__(SC_THREADS)
    mg_scatter_records_kernel(const uint64_t* __restrict__ mass, const uint32_t* __restrict__ gpos,
                              const uint32_t* __restrict__ prot, const uint16_t* __restrict__ len, uint64_t n,
                              uint64_t sub, const __grid_constant__ MgPlan pl, const __grid_constant__ MgRecDst dst,
                              unsigned long long* lookback, uint32_t* tile_counter) {
  __shared__ ScSmem s;
  ScCtx cx;
  sc_rank_tile(s, cx, mass, n, sub, pl, lookback, tile_counter);
  sc_move<uint64_t>(s, cx, [&](uint32_t e) { return mass[cx.base + e]; }, dst.mass);
  sc_move<uint32_t>(s, cx, [&](uint32_t e) { return gpos[cx.base + e]; }, dst.gpos);
  sc_move<uint32_t>(s, cx, [&](uint32_t e) { return prot[cx.base + e]; }, dst.prot);
  sc_move<uint16_t>(s, cx, [&](uint32_t e) { return len[cx.base + e]; }, dst.len);
  __threadfence_system();  // peer stores visible before the kernel is reported complete
}

// stage 1: variant groups -> the owners of their VARIANT mass slices.  A group travels with
// everything its expansion needs: key, payload (the peptide field becomes the ARRIVAL ROW, which is
// how the receiver finds the side tables), the peptide's global id and its C site masks.
__global__ void __launch_bounds__(SC_THREADS)
    mg_scatter_groups_kernel(const uint64_t* __restrict__ key, const uint64_t* __restrict__ pay,
                             const uint64_t* __restrict__ cmask, int C, uint64_t id_off, uint64_t n,
                             const __grid_constant__ MgPlan pl, const __grid_constant__ MgGrpDst dst,
                             unsigned long long* lookback, uint32_t* tile_counter) {
  __shared__ ScSmem s;
  ScCtx cx;
  sc_rank_tile(s, cx, key, n, 0, pl, lookback, tile_counter);
  sc_move<uint64_t>(s, cx, [&](uint32_t e) { return key[cx.base + e]; }, dst.key);
  if (C == 0) {  // per-variant records (more than 32 class sequences): payload = peptide id | pattern, nothing else
    sc_move<uint64_t>(s, cx, [&](uint32_t e) { return pay[cx.base + e] + (id_off << 32); }, dst.pay);
    __threadfence_system();
    return;
  }
  {  // payload: peptide field := arrival row at the destination
    const int t = threadIdx.x, w = t >> 5;
    const unsigned l = lane_id();
#pragma unroll
    for (int i = 0; i < SC_IPT; ++i)
      if (cx.valid[i]) s.stage[cx.slot[i]] = pay[cx.base + (uint32_t)w * 32 * SC_IPT + i * 32 + l];
    __syncthreads();
    for (uint32_t j = t; j < cx.tile_n; j += SC_THREADS) {
      const uint32_t d = s.sdest[j];
      const uint64_t row = s.row0[d] + (j - s.dstart[d]);
      dst.pay[d][row] = (row << 32) | (s.stage[j] & 0xffffffffull);
    }
    __syncthreads();
  }
  sc_move<uint32_t>(s, cx, [&](uint32_t e) { return (uint32_t)((pay[cx.base + e] >> 32) + id_off); }, dst.gid);
  for (int c = 0; c < C; ++c)
    sc_move<uint64_t>(s, cx, [&](uint32_t e) {
      return cmask[(pay[cx.base + e] >> 32) * (uint64_t)C + c];  // the payload names the LOCAL row of the (own) peptide
    }, dst.mask, (uint32_t)C, (uint32_t)c);
  __threadfence_system();
}

__global__ void __launch_bounds__(MG_THREADS)
    plo_to_counts_kernel(const uint64_t* __restrict__ plo, uint64_t n, uint32_t* __restrict__ cnt) {
  const uint64_t i = (uint64_t)blockIdx.x * MG_THREADS + threadIdx.x;
  if (i < n) cnt[i] = (uint32_t)(plo[i + 1] - plo[i]);
}

constexpr int FS_IPT = kScanTile / MG_THREADS;

__global__ void __launch_bounds__(MG_THREADS)
    tile_sums_kernel(const uint32_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ tile_sums) {
  // 64-bit sums: the inputs are variant counts of up to 2^27 each (4096 of them wrap 32 bits)
  __shared__ unsigned long long scratch[MG_THREADS / 32 + 1];
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  unsigned long long sum = 0;
  for (int k = 0; k < FS_IPT; ++k) {
    const uint64_t i = tile_base + (uint64_t)k * MG_THREADS + threadIdx.x;
    if (i < n) sum += in[i];
  }
  unsigned long long total;
  block_exclusive_sum<unsigned long long, MG_THREADS>(sum, scratch, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// exclusive scan of u64 tile sums in place; offs[n] = total (single CTA)
__global__ void __launch_bounds__(1024)
    scan_u64_kernel(const unsigned long long* __restrict__ in, uint64_t n, unsigned long long* __restrict__ offs) {
  __shared__ unsigned long long scratch[1024 / 32 + 1];
  unsigned long long carry = 0;
  for (uint64_t base = 0; base < n; base += 1024) {
    const uint64_t i = base + threadIdx.x;
    const unsigned long long v = (i < n) ? in[i] : 0ull;
    unsigned long long total;
    const unsigned long long ex = block_exclusive_sum<unsigned long long, 1024>(v, scratch, &total);
    if (i < n) offs[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) offs[n] = carry;
}

__global__ void __launch_bounds__(MG_THREADS)
    tile_scan_kernel(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ tile_offs,
                     uint64_t ntiles, uint64_t* __restrict__ offs) {
  __shared__ unsigned long long scratch[MG_THREADS / 32 + 1];
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t running = tile_offs[blockIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) offs[n] = tile_offs[ntiles];
  for (int k = 0; k < FS_IPT; ++k) {
    const uint64_t i = tile_base + (uint64_t)k * MG_THREADS + threadIdx.x;
    const unsigned long long v = (i < n) ? in[i] : 0ull;
    unsigned long long total;
    const unsigned long long ex = block_exclusive_sum<unsigned long long, MG_THREADS>(v, scratch, &total);
    if (i < n) offs[i] = running + ex;
    running += total;
  }
}

}  // namespace

void launch_mg_hist(const uint64_t* key, uint64_t n, uint64_t sub, int shift, const uint64_t* wpay, uint64_t wmask,
                    uint32_t wadd, const uint8_t* wcode, const uint32_t* wtab, const uint32_t* gtab,
                    unsigned long long* hist, cudaStream_t s) {
  if (n == 0) return;
  uint64_t g = (uint64_t)kNumSMsB200 * 4;
  uint64_t per_cta = (n + g - 1) / g;
  if (per_cta > kMgMaxItemsPerCta) per_cta = kMgMaxItemsPerCta;
  per_cta = (per_cta + MG_THREADS - 1) / MG_THREADS * MG_THREADS;
  g = (n + per_cta - 1) / per_cta;
  DBI_LAUNCH(mg_hist_kernel, (unsigned)g, MG_THREADS, 0, s, key, n, per_cta, sub, shift, wpay, wmask, wadd, wcode, wtab,
             gtab, hist);
}

void launch_mg_count(const uint64_t* key, uint64_t n, uint64_t sub, const MgPlan& pl, unsigned long long* counts,
                     cudaStream_t s) {
  if (n == 0) return;
  uint64_t g = (n + MG_THREADS * 16 - 1) / (MG_THREADS * 16);
  if (g > (uint64_t)kNumSMsB200 * 8) g = (uint64_t)kNumSMsB200 * 8;
  DBI_LAUNCH(mg_count_kernel, (unsigned)g, MG_THREADS, 0, s, key, n, sub, pl, counts);
}

size_t mg_scatter_tmp_bytes(uint64_t n) {
  const uint64_t tiles = (n + SC_TILE - 1) / SC_TILE;
  return 256 + (size_t)tiles * kMaxRanks * 8;
}

void launch_mg_scatter_records(const uint64_t* mass, const uint32_t* gpos, const uint32_t* prot, const uint16_t* len,
                               uint64_t n, uint64_t sub, const MgPlan& pl, const MgRecDst& dst, void* tmp,
                               cudaStream_t s) {
  if (n == 0) return;
  const uint64_t tiles = (n + SC_TILE - 1) / SC_TILE;
  DBI_CUDA(cudaMemsetAsync(tmp, 0, mg_scatter_tmp_bytes(n), s));
  DBI_LAUNCH(mg_scatter_records_kernel, (unsigned)tiles, SC_THREADS, 0, s, mass, gpos, prot, len, n, sub, pl, dst,
             (unsigned long long*)((uint8_t*)tmp + 256), (uint32_t*)tmp);
}

void launch_mg_scatter_groups(const uint64_t* key, const uint64_t* pay, const uint64_t* cmask, int C, uint64_t id_off,
                              uint64_t n, const MgPlan& pl, const MgGrpDst& dst, void* tmp, cudaStream_t s) {
  if (n == 0) return;
  const uint64_t tiles = (n + SC_TILE - 1) / SC_TILE;
  DBI_CUDA(cudaMemsetAsync(tmp, 0, mg_scatter_tmp_bytes(n), s));
  DBI_LAUNCH(mg_scatter_groups_kernel, (unsigned)tiles, SC_THREADS, 0, s, key, pay, cmask, C, id_off, n, pl, dst,
             (unsigned long long*)((uint8_t*)tmp + 256), (uint32_t*)tmp);
}

void launch_plo_to_counts(const uint64_t* plo, uint64_t n, uint32_t* cnt, cudaStream_t s) {
  if (n == 0) return;
  DBI_LAUNCH(plo_to_counts_kernel, (unsigned)((n + MG_THREADS - 1) / MG_THREADS), MG_THREADS, 0, s, plo, n, cnt);
}

size_t full_scan_tmp_bytes(uint64_t n) {
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  return (tiles + 1) * 8 + tiles * 8 + 64;
}

void launch_full_scan_u32_to_u64(const uint32_t* in, uint64_t n, uint64_t* offs, void* tmp, cudaStream_t s) {
  if (n == 0) {
    DBI_CUDA(cudaMemsetAsync(offs, 0, 8, s));
    return;
  }
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  uint64_t* tile_offs = (uint64_t*)tmp;
  unsigned long long* tile_sums = (unsigned long long*)((uint8_t*)tmp + (tiles + 1) * 8);
  DBI_LAUNCH(tile_sums_kernel, (unsigned)tiles, MG_THREADS, 0, s, in, n, tile_sums);
  DBI_LAUNCH(scan_u64_kernel, 1, 1024, 0, s, tile_sums, tiles, (unsigned long long*)tile_offs);
  DBI_LAUNCH(tile_scan_kernel, (unsigned)tiles, MG_THREADS, 0, s, in, n, tile_offs, tiles, offs);
}

}  // namespace dbi
