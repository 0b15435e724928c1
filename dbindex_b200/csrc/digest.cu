// digest.cu -- K1 pack, K2 digest-count, K3 scan, K4 digest-emit.
//
// Restates DBIndexer.cutSeq (DBIndexer.java:237-405) as a two-pass count / scan /
// emit over a device-resident residue buffer.  Work unit = one START residue:
// a lane walks `end` sequentially and adds residue masses in IEEE double in the
// reference's order (DBIndexer.java:265-271,308), so masses are bit-identical to
// the Java loop.  Output order is the reference's addSequence()
// call order (protein, start, end ascending) because the scan runs over starts in
// buffer order and a lane emits its ends in walking order.
//
// Layout: res[] = 0, protein0, 0, protein1, 0, ... ; a 0 byte is "outside the
// protein" on both sides, so `start == 0` and `end == length-1` of the reference
// become "previous / next byte is 0".
#include "kernels.cuh"

namespace dbi {
namespace {

constexpr int DG_THREADS = 256;
constexpr int DG_SPT = kDigestTile / DG_THREADS;  // starts per thread

struct TileTables {
  double mass[256];
  uint8_t flags[256];
};

__device__ __forceinline__ void load_tables(TileTables& tt, const DevTables* __restrict__ tb) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    tt.mass[i] = tb->mass[i];
    tt.flags[i] = tb->flags[i];
  }
}

// The body of the reference's `for start` iteration (DBIndexer.java:256-395) for the
// start at buffer position g.  F is called as emit(mass, len) for every record.
template <typename F>
__device__ __forceinline__ uint32_t walk_start(const uint8_t* __restrict__ res, uint32_t g, const TileTables& tt,
                                               const DigestCfg& cfg, uint32_t* err, F&& emit) {
  const uint8_t c0 = ld_res(res, g);
  if (c0 == 0) return 0;  // separator, not a residue
  const uint8_t prev = ld_res(res, g - 1);
  // Enzyme.checkCleavage, N side (contract of SURVEY.md 8c): start == 0 or the previous
  // residue is an enzyme residue and this one is not a no-cut residue
  const bool n_ok = (prev == 0) || ((tt.flags[prev] & kFlagEnzyme) && !(tt.flags[c0] & kFlagNocut));
  // full specificity: cleavageStatus can never become true for this start, so the
  // reference walks it without ever reaching addSequence -- nothing to emit
  if (!cfg.semi && !n_ok) return 0;

  double mass = cfg.init_mass;  // DBIndexer.java:265-271
  int mc = -1;                  // intMisCleavageCount, :280
  uint32_t len = 0, count = 0;
  uint32_t pos = g;
  uint8_t c = c0;
  while (mass <= cfg.max_mass && c != 0) {  // :284  (end < length  <=>  c != 0)
    ++len;                                  // pepSize++, :285
    mass = __dadd_rn(mass, tt.mass[c]);     // precMass = precMass + aaMass, :308
    const uint8_t fc = tt.flags[c];
    if (fc & kFlagEnzyme) ++mc;             // :314-316
    const uint8_t nxt = ld_res(res, pos + 1);
    const bool c_ok = (nxt == 0) || ((fc & kFlagEnzyme) && !(tt.flags[nxt] & kFlagNocut));
    const bool cleavage = cfg.semi ? (n_ok || c_ok) : c_ok;  // n_ok is known true when !semi
    if (cleavage) {                          // :320
      if (mc > cfg.max_mc) break;            // :322-324
      if (mass > cfg.max_mass) break;        // :327-329
      if ((int)len >= cfg.min_len && mass >= cfg.min_mass) {  // :331
        if (len > DBI_MAX_PEP_LEN) {
          atomicOr(err, kErrPepTooLong);
          break;
        }
        emit(mass, len);
        ++count;
      }
    }
    ++pos;  // ++end, :394
    c = nxt;
  }
  return count;
}

// Can the start at buffer position g emit anything?  Separators never, and with full
// specificity only starts right after a cleavage site (or at the protein N-terminus): for
// any other start checkCleavage never becomes true, so the reference walks it in vain.
__device__ __forceinline__ bool start_is_live(const uint8_t* __restrict__ res, uint32_t g, const TileTables& tt,
                                              const DigestCfg& cfg) {
  const uint8_t c0 = ld_res(res, g);
  if (c0 == 0) return false;
  if (cfg.semi) return true;
  const uint8_t prev = ld_res(res, g - 1);
  return (prev == 0) || ((tt.flags[prev] & kFlagEnzyme) && !(tt.flags[c0] & kFlagNocut));
}

// Warp-cooperative cleavage-site scan: the CTA compacts the live starts of its tile, in buffer
// order, into s_list (tile-local positions).  For trypsin only ~11 % of the residues start a
// peptide; without this the walks below ran with 4 of 32 lanes busy (ncu, profiles/).
__device__ __forceinline__ uint32_t compact_live_starts(const uint8_t* __restrict__ res, uint32_t tile_base,
                                                        uint32_t res_end, const TileTables& tt, const DigestCfg& cfg,
                                                        uint16_t* s_list, uint32_t* scratch) {
  // thread t owns the DG_SPT consecutive starts t*DG_SPT .. : order is preserved by a block scan
  uint32_t live = 0, n_live = 0;
#pragma unroll
  for (int k = 0; k < DG_SPT; ++k) {
    const uint32_t g = tile_base + threadIdx.x * DG_SPT + k;
    if (g < res_end && start_is_live(res, g, tt, cfg)) {
      live |= 1u << k;
      ++n_live;
    }
  }
  uint32_t total;
  uint32_t o = block_exclusive_sum<uint32_t, DG_THREADS>(n_live, scratch, &total);
#pragma unroll
  for (int k = 0; k < DG_SPT; ++k)
    if (live & (1u << k)) s_list[o++] = (uint16_t)(threadIdx.x * DG_SPT + k);
  __syncthreads();
  return total;
}

// ---- K2 -----------------------------------------------------------------------------
__global__ void __launch_bounds__(DG_THREADS)
    digest_count_kernel(const uint8_t* __restrict__ res, uint32_t res_end, const DevTables* __restrict__ tb,
                        DigestCfg cfg, uint32_t tile0, uint32_t* __restrict__ tile_counts, uint32_t* err) {
  __shared__ TileTables tt;
  __shared__ uint16_t s_list[kDigestTile];
  __shared__ uint32_t scratch[DG_THREADS / 32 + 1];
  load_tables(tt, tb);
  __syncthreads();
  const uint32_t tile_base = 1u + (tile0 + blockIdx.x) * (uint32_t)kDigestTile;
  const uint32_t n_live = compact_live_starts(res, tile_base, res_end, tt, cfg, s_list, scratch);
  uint32_t cnt = 0;
  for (uint32_t i = threadIdx.x; i < n_live; i += DG_THREADS)
    cnt += walk_start(res, tile_base + s_list[i], tt, cfg, err, [](double, uint32_t) {});
  uint32_t total;
  block_exclusive_sum<uint32_t, DG_THREADS>(cnt, scratch, &total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// ---- K4 -----------------------------------------------------------------------------
__global__ void __launch_bounds__(DG_THREADS, 6)
    digest_emit_kernel(const uint8_t* __restrict__ res, uint32_t res_end, const DevTables* __restrict__ tb,
                       DigestCfg cfg, uint32_t tile0, const uint64_t* __restrict__ tile_offs,
                       const uint32_t* __restrict__ pstart, uint32_t n_prot, uint64_t* __restrict__ o_mass,
                       uint32_t* __restrict__ o_gpos, uint32_t* __restrict__ o_prot, uint16_t* __restrict__ o_len,
                       uint32_t* err) {
  __shared__ TileTables tt;
  __shared__ uint16_t s_list[kDigestTile];
  __shared__ uint32_t s_cnt[kDigestTile + 1];
  __shared__ uint32_t scratch[DG_THREADS / 32 + 1];
  load_tables(tt, tb);
  __syncthreads();
  const uint32_t tile_base = 1u + (tile0 + blockIdx.x) * (uint32_t)kDigestTile;
  const uint32_t n_live = compact_live_starts(res, tile_base, res_end, tt, cfg, s_list, scratch);
  // pass A: records per live start
  for (uint32_t i = threadIdx.x; i < n_live; i += DG_THREADS)
    s_cnt[i] = walk_start(res, tile_base + s_list[i], tt, cfg, err, [](double, uint32_t) {});
  __syncthreads();
  // exclusive scan over the live starts in buffer order (thread t scans DG_SPT consecutive entries)
  uint32_t local[DG_SPT];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < DG_SPT; ++k) {
    const uint32_t i = threadIdx.x * DG_SPT + k;
    local[k] = sum;
    sum += (i < n_live) ? s_cnt[i] : 0u;
  }
  uint32_t total;
  const uint32_t ex = block_exclusive_sum<uint32_t, DG_THREADS>(sum, scratch, &total);
#pragma unroll
  for (int k = 0; k < DG_SPT; ++k) {
    const uint32_t i = threadIdx.x * DG_SPT + k;
    if (i < n_live) s_cnt[i] = ex + local[k];
  }
  if (threadIdx.x == 0) s_cnt[n_live] = total;
  __syncthreads();
  if (total == 0) return;
  const uint64_t tile_off = tile_offs[blockIdx.x];
  // pass B: walk again and write
  for (uint32_t i = threadIdx.x; i < n_live; i += DG_THREADS) {
    if (s_cnt[i + 1] == s_cnt[i]) continue;  // this start emits nothing
    const uint32_t g = tile_base + s_list[i];
    // protein of this start: last p with pstart[p] <= g
    uint32_t lo = 0, hi = n_prot;  // invariant: pstart[lo] <= g < pstart[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(pstart + mid) <= g) lo = mid; else hi = mid;
    }
    const uint32_t prot = lo;
    uint64_t o = tile_off + s_cnt[i];
    walk_start(res, g, tt, cfg, err, [&](double m, uint32_t len) {
      o_mass[o] = (uint64_t)__double_as_longlong(m);
      o_gpos[o] = g;
      o_prot[o] = prot;
      o_len[o] = (uint16_t)len;
      ++o;
    });
  }
}

// ---- K1 -----------------------------------------------------------------------------
constexpr int PK_THREADS = 256;
constexpr int PK_BYTES = 16;  // raw bytes per thread

__global__ void __launch_bounds__(PK_THREADS)
    pack_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ off, uint32_t n_prot, uint64_t n_res,
                uint8_t* __restrict__ res, uint32_t* __restrict__ pstart, uint32_t* err) {
  const uint64_t tid = (uint64_t)blockIdx.x * PK_THREADS + threadIdx.x;
  // protein starts and separators
  if (tid <= n_prot) {
    const uint64_t ps = off[tid] + tid + 1;  // pstart[n_prot] = one past the last separator
    pstart[tid] = (uint32_t)ps;
    res[ps - 1] = 0;  // separator before protein tid (tid == n_prot: the final one)
  }
  const uint64_t b0 = tid * PK_BYTES;
  if (b0 >= n_res) return;
  // protein containing raw byte b0: last p with off[p] <= b0 (empty proteins share an offset)
  uint32_t lo = 0, hi = n_prot;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (off[mid] <= b0) lo = mid; else hi = mid;
  }
  uint32_t p = lo;
  uint64_t next = off[p + 1];
  const uint64_t b1 = min(b0 + PK_BYTES, n_res);
  for (uint64_t b = b0; b < b1; ++b) {
    while (b >= next) { ++p; next = off[p + 1]; }
    const uint8_t c = raw[b];
    if (c == 0) atomicOr(err, kErrZeroResidue);
    res[b + p + 1] = c;
  }
}

// ---- K3 -----------------------------------------------------------------------------
constexpr int SC_THREADS = 1024;

__global__ void __launch_bounds__(SC_THREADS)
    scan_u32_to_u64_kernel(const uint32_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ offs) {
  __shared__ unsigned long long scratch[SC_THREADS / 32 + 1];
  unsigned long long carry = 0;
  for (uint64_t base = 0; base < n; base += SC_THREADS) {
    const uint64_t i = base + threadIdx.x;
    const unsigned long long v = (i < n) ? in[i] : 0ull;
    unsigned long long total;
    const unsigned long long ex = block_exclusive_sum<unsigned long long, SC_THREADS>(v, scratch, &total);
    if (i < n) offs[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) offs[n] = carry;
}

}  // namespace

void launch_pack(const uint8_t* d_raw, const uint64_t* d_off, uint32_t n_prot, uint64_t n_res, uint8_t* d_res,
                 uint32_t* d_pstart, uint32_t* d_err, cudaStream_t s) {
  uint64_t threads = (n_res + PK_BYTES - 1) / PK_BYTES;
  if (threads < (uint64_t)n_prot + 1) threads = (uint64_t)n_prot + 1;
  const unsigned grid = (unsigned)((threads + PK_THREADS - 1) / PK_THREADS);
  DBI_LAUNCH(pack_kernel, grid, PK_THREADS, 0, s, d_raw, d_off, n_prot, n_res, d_res, d_pstart, d_err);
}

void launch_digest_count(const uint8_t* d_res, uint32_t res_end, const DevTables* d_tb, const DigestCfg& cfg,
                         uint32_t tile0, uint32_t ntiles, uint32_t* d_tile_counts, uint32_t* d_err, cudaStream_t s) {
  if (ntiles == 0) return;
  DBI_LAUNCH(digest_count_kernel, ntiles, DG_THREADS, 0, s, d_res, res_end, d_tb, cfg, tile0, d_tile_counts, d_err);
}

void launch_scan_u32_to_u64(const uint32_t* d_in, uint64_t n, uint64_t* d_offs, cudaStream_t s) {
  DBI_LAUNCH(scan_u32_to_u64_kernel, 1, SC_THREADS, 0, s, d_in, n, d_offs);
}

void launch_digest_emit(const uint8_t* d_res, uint32_t res_end, const DevTables* d_tb, const DigestCfg& cfg,
                        uint32_t tile0, uint32_t ntiles, const uint64_t* d_tile_offs, const uint32_t* d_pstart,
                        uint32_t n_prot, uint64_t* o_mass, uint32_t* o_gpos, uint32_t* o_prot, uint16_t* o_len,
                        uint32_t* d_err, cudaStream_t s) {
  if (ntiles == 0) return;
  DBI_LAUNCH(digest_emit_kernel, ntiles, DG_THREADS, 0, s, d_res, res_end, d_tb, cfg, tile0, d_tile_offs, d_pstart,
             n_prot, o_mass, o_gpos, o_prot, o_len, d_err);
}

}  // namespace dbi
