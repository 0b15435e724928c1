// mods.cu -- K5 mod-count, K6 mod-emit: differential-modification combinatorics.
//
// The reference only PARSES differential mods into a residue->shift table
// (model/DiffModification.java:11-54, io/SearchParamReader.java:631-667) and a
// per-peptide cap (max_num_differential_AA_per_mod, SearchParamReader.java:322); it
// never expands them into the index.  The expansion below is this project's SPEC
// (DESIGN.md "Mod expansion SPEC", restated by oracle/dbindex_oracle.cpp):
//   * runs over the UNIQUE base peptides (after K8), so protein lists are shared;
//   * eligible sites = peptide positions whose residue has a shift;
//   * a variant = a set of k <= max_mods distinct sites, enumerated k ascending then
//     lexicographically by site; k = 0 (unmodified) is always present;
//   * mass = base mass, then + shift per chosen site left to right (IEEE double);
//   * the mass gate min <= m <= max is re-applied to every modified variant.
//
// Two-pass count / scan / emit.  One WARP per base peptide: the lanes find the
// sites with a ballot, then each lane takes variant ranks lane, lane+32, ... and
// un-ranks them (combinatorial number system) so that consecutive lanes write
// consecutive output slots -- coalesced 8-byte stores, no per-thread serial runs.
// The count is closed-form (sum of binomials) whenever no variant of the peptide can
// reach a mass gate, which is the case for all but the heaviest few percent.
#include <cstdlib>

#include "kernels.cuh"

namespace dbi {
namespace {

constexpr int MD_THREADS = 256;
constexpr int MD_WARPS = MD_THREADS / 32;
constexpr int MD_PER_WARP = kModTile / MD_WARPS;  // bases per warp per tile
constexpr int MD_MAX_SITES = DBI_MAX_MOD_POS + 1;  // positions 0..254

struct ModTables {
  double diff[256];
  uint8_t flags[256];
  uint8_t cls[256];
  double cls_delta[16];
};

struct WarpSites {
  uint8_t pos[256];
  uint8_t res[256];
};

__device__ __forceinline__ void load_mod_tables(ModTables& mt, const DevTables* __restrict__ tb) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    mt.diff[i] = tb->diff[i];
    mt.flags[i] = tb->flags[i];
    mt.cls[i] = tb->cls[i];
    if (i < 16) mt.cls_delta[i] = tb->cls_delta[i];
  }
}

// Heaviest peptides first: the bases are mass-sorted and the variant count grows like the
// cube of the site count, so the last tiles carry orders of magnitude more work.
__device__ __forceinline__ uint32_t mod_tile_local() { return gridDim.x - 1 - blockIdx.x; }

// Gated count of one peptide with one lane per class SEQUENCE (n_seq <= 32).  The variant mass
// depends only on the sequence of shift classes along the chosen sites, so
//   #passing = sum over sequences s of [gate(f_s(base))] * #(occurrences of s as a subsequence
//              of the peptide's site-class string),
// and the occurrence counts follow from one pass over the sites: a site of class c extends every
// sequence's parent (heap numbering: children of v are v*C + c + 1).
// C(m, j) for j <= 4, m <= 255
__device__ __forceinline__ uint32_t binom(uint32_t m, int j) {
  switch (j) {
    case 0: return 1u;
    case 1: return m;
    case 2: return m < 2 ? 0u : m * (m - 1) / 2;
    case 3: return m < 3 ? 0u : m * (m - 1) * (m - 2) / 6;
    default: return m < 4 ? 0u : (uint32_t)((uint64_t)m * (m - 1) * (m - 2) * (m - 3) / 24);
  }
}

// number of subsets of size <= K of n sites
__device__ __forceinline__ uint32_t total_variants(uint32_t n, int K) {
  uint32_t t = 0;
  for (int k = 0; k <= K; ++k) t += binom(n, k);
  return t;
}

// Whole warp: eligible sites of the peptide res[g .. g+len) into ws (positions ascending).
// Returns the number of sites (<= 255); *bad is set if one lies beyond DBI_MAX_MOD_POS.
__device__ __forceinline__ int warp_collect_sites(const uint8_t* __restrict__ res, uint32_t g, uint32_t len,
                                                  const ModTables& mt, WarpSites& ws, bool* bad) {
  const unsigned l = lane_id();
  int n = 0;
  for (uint32_t b = 0; b < len; b += 32) {
    const uint32_t i = b + l;
    const uint8_t c = (i < len) ? ld_res(res, g + i) : (uint8_t)0;
    const bool is = (i < len) && (mt.flags[c] & kFlagDiffMod);
    const unsigned m = __ballot_sync(0xffffffffu, is);
    if (is) {
      if (i > DBI_MAX_MOD_POS) {
        *bad = true;
      } else {
        const int slot = n + __popc(m & lanemask_lt());
        ws.pos[slot] = (uint8_t)i;
        ws.res[slot] = c;
      }
    }
    n += __popc(m);
  }
  *bad = __any_sync(0xffffffffu, *bad);  // warp-uniform verdict
  __syncwarp();                          // the site list is read by other lanes next
  return n > MD_MAX_SITES ? MD_MAX_SITES : n;
}

// Lane v = class sequence v (heap numbering).  Outputs this lane's occurrence count, the mass
// its variants have, and whether that mass passes the gate.
__device__ __forceinline__ void warp_seq_dp(uint32_t n, double base_mass, const WarpSites& ws, const ModTables& mt,
                                            const DigestCfg& cfg, uint32_t* cnt_out, double* val_out, bool* pass_out) {
  const unsigned v = lane_id();
  const int C = cfg.n_classes;
  const bool active = (int)v < cfg.n_seq;
  int depth = 0;
  double val = base_mass;
  {
    int cl[DBI_MAX_MODS_PER_PEP];
    uint32_t x = v;
    while (active && x > 0) { cl[depth++] = (int)((x - 1) % C); x = (x - 1) / C; }
    for (int i = depth - 1; i >= 0; --i) val = __dadd_rn(val, mt.cls_delta[cl[i]]);  // first chosen site first
  }
  const int last = (active && v > 0) ? (int)((v - 1) % C) : -1;
  const int parent = (active && v > 0) ? (int)((v - 1) / C) : 0;
  uint32_t cnt = (v == 0) ? 1u : 0u;
  for (uint32_t j = 0; j < n; ++j) {
    const int c = mt.cls[ws.res[j]];
    const uint32_t up = __shfl_sync(0xffffffffu, cnt, parent);  // parent's count BEFORE this site
    if (last == c) cnt += up;
  }
  *cnt_out = cnt;
  *val_out = val;
  *pass_out = active && (v == 0 || (val >= cfg.min_mass && val <= cfg.max_mass));
}

__device__ __forceinline__ uint32_t warp_gated_count(uint32_t n, double base_mass, const WarpSites& ws,
                                                     const ModTables& mt, const DigestCfg& cfg) {
  uint32_t cnt;
  double val;
  bool pass;
  warp_seq_dp(n, base_mass, ws, mt, cfg, &cnt, &val, &pass);
  uint32_t tot = pass ? cnt : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  return tot;
}

// Variant of global rank r (0 = unmodified; then k = 1, 2, ... each in lexicographic order):
// mass and mod pattern.  cn[k] = C(n, k) is warp-uniform (computed once per peptide).  The
// combinatorial-number-system walk keeps its binomials incrementally
// (C(m-1, j) = C(m, j) - C(m-1, j-1)), so a step of the search is 3-5 integer instructions.
__device__ __forceinline__ void decode_variant(uint32_t r, uint32_t n, int K, const uint32_t* cn, double base_mass,
                                               const WarpSites& ws, const ModTables& mt, double* mass,
                                               uint32_t* pat) {
  int k = 0;
#pragma unroll
  for (int q = 0; q <= DBI_MAX_MODS_PER_PEP; ++q) {
    if (q <= K && k == q && r >= cn[q]) {
      r -= cn[q];
      k = q + 1;
    }
  }
  double m = base_mass;
  uint32_t p = 0;
  uint32_t x = 0;  // next candidate site ordinal
#pragma unroll
  for (int i = 0; i < DBI_MAX_MODS_PER_PEP; ++i) {
    if (i < k) {
      const int j = k - 1 - i;       // elements still to choose after this one
      const uint32_t rest = n - x;   // candidates x .. n-1
      if (j == 0) {
        x += r;
        r = 0;
      } else if (j == 1) {
        uint32_t w = rest - 1;       // C(rest-1, 1): subsets whose next element is x
        while (r >= w) { r -= w; --w; ++x; }
      } else if (j == 2) {
        uint32_t d = rest - 2;
        uint32_t w = (rest - 1) * d / 2;  // C(rest-1, 2)
        while (r >= w) { r -= w; w -= d; --d; ++x; }
      } else {
        uint32_t d = rest - 3;
        uint32_t w2 = (rest - 2) * d / 2;                 // C(rest-2, 2)
        uint32_t w = (uint32_t)((uint64_t)(rest - 1) * (rest - 2) * d / 6);  // C(rest-1, 3)
        while (r >= w) { r -= w; w -= w2; w2 -= d; --d; ++x; }
      }
      m = __dadd_rn(m, mt.diff[ws.res[x]]);  // left to right
      p |= ((uint32_t)ws.pos[x] + 1u) << (8 * i);
      ++x;
    }
  }
  *mass = m;
  *pat = p;
}

__global__ void __launch_bounds__(MD_THREADS)
    mod_count_kernel(const uint8_t* __restrict__ res, const DevTables* __restrict__ tb, DigestCfg cfg,
                     const double* __restrict__ u_mass, const uint32_t* __restrict__ u_gpos,
                     const uint16_t* __restrict__ u_len, uint64_t n_unique, uint32_t tile0,
                     uint32_t* __restrict__ counts, uint32_t* __restrict__ tile_counts, uint32_t* err) {
  __shared__ ModTables mt;
  __shared__ WarpSites wsites[MD_WARPS];
  __shared__ uint32_t wsum[MD_WARPS];
  load_mod_tables(mt, tb);
  __syncthreads();
  const int w = threadIdx.x >> 5;
  const unsigned l = lane_id();
  WarpSites& ws = wsites[w];
  const uint64_t u0 = (uint64_t)(tile0 + mod_tile_local()) * kModTile + (uint64_t)w * MD_PER_WARP;
  const int K = cfg.max_mods;
  uint32_t sum = 0;
  for (int q = 0; q < MD_PER_WARP; ++q) {
    const uint64_t u = u0 + q;
    if (u >= n_unique) break;
    const double bm = u_mass[u];
    bool bad = false;
    const uint32_t n = (uint32_t)warp_collect_sites(res, u_gpos[u], u_len[u], mt, ws, &bad);
    if (bad && l == 0) atomicOr(err, kErrModPos);
    const int kk = K < (int)n ? K : (int)n;
    uint32_t cn[DBI_MAX_MODS_PER_PEP + 1];
#pragma unroll
    for (int q = 0; q <= DBI_MAX_MODS_PER_PEP; ++q) cn[q] = binom(n, q);
    const uint32_t total = total_variants(n, kk);
    uint32_t c;
    // no variant can reach a gate: every subset counts
    if (bm + kk * cfg.mod_hi <= cfg.max_mass - 1e-6 && bm + kk * cfg.mod_lo >= cfg.min_mass + 1e-6) {
      c = total;
    } else if (cfg.n_seq > 0) {
      c = warp_gated_count(n, bm, ws, mt, cfg);
    } else {
      c = 0;
      for (uint32_t r0 = 0; r0 < total; r0 += 32) {
        const uint32_t r = r0 + l;
        bool pass = false;
        if (r < total) {
          double m;
          uint32_t pat;
          decode_variant(r, n, kk, cn, bm, ws, mt, &m, &pat);
          pass = (r == 0) || (m >= cfg.min_mass && m <= cfg.max_mass);
        }
        c += __popc(__ballot_sync(0xffffffffu, pass));
      }
    }
    if (l == 0) counts[u] = c;
    sum += c;
    __syncwarp();
  }
  if (l == 0) wsum[w] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < MD_WARPS; ++i) t += wsum[i];
    tile_counts[mod_tile_local()] = t;
  }
}

__global__ void __launch_bounds__(MD_THREADS)
    mod_emit_kernel(const uint8_t* __restrict__ res, const DevTables* __restrict__ tb, DigestCfg cfg,
                    const double* __restrict__ u_mass, const uint32_t* __restrict__ u_gpos,
                    const uint16_t* __restrict__ u_len, uint64_t n_unique, uint32_t tile0,
                    const uint32_t* __restrict__ counts, const uint64_t* __restrict__ tile_offs, uint64_t base_bits,
                    uint64_t* __restrict__ v_key, uint64_t* __restrict__ v_payload) {
  __shared__ ModTables mt;
  __shared__ WarpSites wsites[MD_WARPS];
  __shared__ uint32_t wsum[MD_WARPS];
  load_mod_tables(mt, tb);
  const int w = threadIdx.x >> 5;
  const unsigned l = lane_id();
  WarpSites& ws = wsites[w];
  const uint64_t u0 = (uint64_t)(tile0 + mod_tile_local()) * kModTile + (uint64_t)w * MD_PER_WARP;
  // output offset of this warp's first base: tile offset + counts of the earlier warps' bases
  uint32_t part = 0;
  for (int q = l; q < MD_PER_WARP; q += 32) {
    const uint64_t u = u0 + q;
    if (u < n_unique) part += counts[u];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (l == 0) wsum[w] = part;
  __syncthreads();
  uint64_t out = tile_offs[mod_tile_local()];
  for (int i = 0; i < w; ++i) out += wsum[i];

  const int K = cfg.max_mods;
  for (int q = 0; q < MD_PER_WARP; ++q) {
    const uint64_t u = u0 + q;
    if (u >= n_unique) break;
    const uint32_t cnt = counts[u];
    const double bm = u_mass[u];
    if (cnt == 1) {  // only the unmodified peptide
      if (l == 0) {
        v_key[out] = (uint64_t)__double_as_longlong(bm) - base_bits;
        v_payload[out] = u << 32;
      }
      out += 1;
      continue;
    }
    bool bad = false;
    const uint32_t n = (uint32_t)warp_collect_sites(res, u_gpos[u], u_len[u], mt, ws, &bad);
    const int kk = K < (int)n ? K : (int)n;
    uint32_t cn[DBI_MAX_MODS_PER_PEP + 1];
#pragma unroll
    for (int q = 0; q <= DBI_MAX_MODS_PER_PEP; ++q) cn[q] = binom(n, q);
    const uint32_t total = total_variants(n, kk);
    const bool gated = cnt != total;
    uint64_t o = out;
    for (uint32_t r0 = 0; r0 < total; r0 += 32) {
      const uint32_t r = r0 + l;
      bool pass = false;
      double m = bm;
      uint32_t pat = 0;
      if (r < total) {
        decode_variant(r, n, kk, cn, bm, ws, mt, &m, &pat);
        pass = !gated || r == 0 || (m >= cfg.min_mass && m <= cfg.max_mass);
      }
      const unsigned pm = __ballot_sync(0xffffffffu, pass);
      if (pass) {
        const uint64_t slot = o + __popc(pm & lanemask_lt());
        v_key[slot] = (uint64_t)__double_as_longlong(m) - base_bits;
        v_payload[slot] = (u << 32) | pat;
      }
      o += __popc(pm);
    }
    out += cnt;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(MD_THREADS)
    split_entries_kernel(const uint64_t* __restrict__ skey, const uint64_t* __restrict__ spayload, uint64_t n,
                         uint64_t base_bits, double* __restrict__ e_mass, uint32_t* __restrict__ e_base,
                         uint32_t* __restrict__ e_pat) {
  const uint64_t i = (uint64_t)blockIdx.x * MD_THREADS + threadIdx.x;
  if (i >= n) return;
  e_mass[i] = __longlong_as_double((long long)(skey[i] + base_bits));
  const uint64_t p = spayload[i];
  e_base[i] = (uint32_t)(p >> 32);
  e_pat[i] = (uint32_t)p;
}


// =====================================================================================
// Group path (n_seq <= 32): sort GROUPS, not entries.
//
// All variants of one peptide whose chosen sites have the same class sequence s share one
// mass f_s(base) -- they are one contiguous run of the final index.  So only one record per
// (peptide, sequence) group goes through the radix sort (~10 per peptide instead of ~26
// entries), and the entries are written after the sort, already in place.
//   K5g grp_count : groups and variants per peptide (lane-per-sequence DP)
//   K6g grp_emit  : {key = mass bits - base, payload = peptide << 32 | sequence << 27 | count}
//   (K7 sorts the records)
//   K6x grp_expand: for every sorted group, enumerate the occurrences of its class sequence in
//                   the peptide's site string and write (mass, peptide, pattern) entries.
static_assert(MD_PER_WARP == 32, "grp_emit_kernel keeps one peptide's group count per lane");
constexpr uint32_t kGrpCntBits = 27;
constexpr uint32_t kGrpCntMask = (1u << kGrpCntBits) - 1;

__global__ void __launch_bounds__(MD_THREADS)
    grp_count_kernel(const uint8_t* __restrict__ res, const DevTables* __restrict__ tb, DigestCfg cfg,
                     const double* __restrict__ u_mass, const uint32_t* __restrict__ u_gpos,
                     const uint16_t* __restrict__ u_len, uint64_t n_unique, uint32_t tile0,
                     uint8_t* __restrict__ ng_out, uint32_t* __restrict__ tile_groups,
                     uint32_t* __restrict__ tile_vars, uint32_t* err) {
  __shared__ ModTables mt;
  __shared__ WarpSites wsites[MD_WARPS];
  __shared__ uint32_t wsum_g[MD_WARPS];
  __shared__ uint32_t wsum_v[MD_WARPS];
  load_mod_tables(mt, tb);
  __syncthreads();
  const int w = threadIdx.x >> 5;
  const unsigned l = lane_id();
  WarpSites& ws = wsites[w];
  const uint64_t u0 = (uint64_t)(tile0 + blockIdx.x) * kModTile + (uint64_t)w * MD_PER_WARP;
  uint32_t sum_g = 0, sum_v = 0;
  for (int q = 0; q < MD_PER_WARP; ++q) {
    const uint64_t u = u0 + q;
    if (u >= n_unique) break;
    bool bad = false;
    const uint32_t n = (uint32_t)warp_collect_sites(res, u_gpos[u], u_len[u], mt, ws, &bad);
    if (bad && l == 0) atomicOr(err, kErrModPos);
    uint32_t ng = 1, nv = 1;
    if (n > 0) {
      uint32_t cnt;
      double val;
      bool pass;
      warp_seq_dp(n, u_mass[u], ws, mt, cfg, &cnt, &val, &pass);
      const bool has = pass && cnt > 0;
      if (has && cnt > kGrpCntMask) atomicOr(err, kErrModPos);  // > 2^27 variants in one group
      ng = (uint32_t)__popc(__ballot_sync(0xffffffffu, has));
      nv = has ? cnt : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    }
    if (l == 0) ng_out[u] = (uint8_t)ng;
    sum_g += ng;
    sum_v += nv;
    __syncwarp();
  }
  if (l == 0) { wsum_g[w] = sum_g; wsum_v[w] = sum_v; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t g = 0, v = 0;
    for (int i = 0; i < MD_WARPS; ++i) { g += wsum_g[i]; v += wsum_v[i]; }
    tile_groups[blockIdx.x] = g;
    tile_vars[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(MD_THREADS)
    grp_emit_kernel(const uint8_t* __restrict__ res, const DevTables* __restrict__ tb, DigestCfg cfg,
                    const double* __restrict__ u_mass, const uint32_t* __restrict__ u_gpos,
                    const uint16_t* __restrict__ u_len, uint64_t n_unique, uint32_t tile0,
                    const uint8_t* __restrict__ ng_in, const uint64_t* __restrict__ tile_goffs, uint64_t base_bits,
                    uint64_t* __restrict__ g_key, uint64_t* __restrict__ g_pay) {
  __shared__ ModTables mt;
  __shared__ WarpSites wsites[MD_WARPS];
  __shared__ uint32_t wsum[MD_WARPS];
  load_mod_tables(mt, tb);
  const int w = threadIdx.x >> 5;
  const unsigned l = lane_id();
  WarpSites& ws = wsites[w];
  const uint64_t u0 = (uint64_t)(tile0 + blockIdx.x) * kModTile + (uint64_t)w * MD_PER_WARP;
  // groups of this warp's peptides (MD_PER_WARP == 32: one per lane)
  uint32_t my_ng = (u0 + l < n_unique) ? ng_in[u0 + l] : 0u;
  uint32_t part = my_ng;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (l == 0) wsum[w] = part;
  __syncthreads();
  uint64_t out = tile_goffs[blockIdx.x];
  for (int i = 0; i < w; ++i) out += wsum[i];
  for (int q = 0; q < MD_PER_WARP; ++q) {
    const uint64_t u = u0 + q;
    if (u >= n_unique) break;
    const double bm = u_mass[u];
    const uint32_t ng = __shfl_sync(0xffffffffu, my_ng, q);
    if (ng == 1) {  // only the unmodified peptide (no sites, or every modified mass is gated out)
      if (l == 0) {
        g_key[out] = (uint64_t)__double_as_longlong(bm) - base_bits;
        g_pay[out] = (u << 32) | 1u;
      }
      out += 1;
      continue;
    }
    bool bad = false;
    const uint32_t n = (uint32_t)warp_collect_sites(res, u_gpos[u], u_len[u], mt, ws, &bad);
    uint32_t cnt;
    double val;
    bool pass;
    warp_seq_dp(n, bm, ws, mt, cfg, &cnt, &val, &pass);
    const bool has = pass && cnt > 0;
    const unsigned gm = __ballot_sync(0xffffffffu, has);
    if (has) {
      const uint64_t slot = out + __popc(gm & lanemask_lt());
      g_key[slot] = (uint64_t)__double_as_longlong(val) - base_bits;
      g_pay[slot] = (u << 32) | ((uint64_t)l << kGrpCntBits) | (uint64_t)(cnt & kGrpCntMask);
    }
    out += ng;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(MD_THREADS)
    grp_extract_cnt_kernel(const uint64_t* __restrict__ pay, uint64_t n, uint32_t* __restrict__ cnt) {
  const uint64_t i = (uint64_t)blockIdx.x * MD_THREADS + threadIdx.x;
  if (i < n) cnt[i] = (uint32_t)pay[i] & kGrpCntMask;
}

// Class sequence of heap node v, packed: nibble j = class of the j-th chosen site (site order).
// Returns the sequence length.  Register-only on purpose (no indexed local arrays).
__device__ __forceinline__ int pack_seq(uint32_t v, int C, uint32_t* packed) {
  uint32_t pk = 0;
  int depth = 0;
  while (v > 0) {  // leaf to root: the last class first, so shifting left ends with class 0 in nibble 0
    pk = (pk << 4) | ((v - 1) % (uint32_t)C);
    v = (v - 1) / (uint32_t)C;
    ++depth;
  }
  *packed = pk;
  return depth;
}
__device__ __forceinline__ int seq_class_at(uint32_t packed, int j) { return (int)((packed >> (4 * j)) & 15u); }
__device__ __forceinline__ uint32_t low_bytes_mask(int n_bytes) {
  return n_bytes >= 4 ? 0xffffffffu : ((1u << (8 * n_bytes)) - 1u);
}

// bits strictly above position i
__device__ __forceinline__ uint64_t above(int i) { return (~1ull) << i; }

// Number of occurrences of the class sequence whose FIRST site is i0 (masks cm1..cm3 = sites of
// the 2nd..4th class of the sequence, k = sequence length).
__device__ __forceinline__ uint32_t count_from(int k, int i0, uint64_t cm1, uint64_t cm2, uint64_t cm3) {
  if (k == 1) return 1u;
  uint64_t m1 = cm1 & above(i0);
  if (k == 2) return (uint32_t)__popcll(m1);
  uint32_t c = 0;
  for (; m1; m1 &= m1 - 1) {
    const int i1 = __ffsll((long long)m1) - 1;
    uint64_t m2 = cm2 & above(i1);
    if (k == 3) {
      c += (uint32_t)__popcll(m2);
    } else {
      for (; m2; m2 &= m2 - 1) c += (uint32_t)__popcll(cm3 & above(__ffsll((long long)m2) - 1));
    }
  }
  return c;
}

// Write those occurrences (site order = lexicographic) as entries starting at slot o.
__device__ __forceinline__ uint64_t emit_from(int k, int i0, uint64_t cm1, uint64_t cm2, uint64_t cm3, double mass,
                                              uint32_t b, uint64_t o, double* __restrict__ e_mass,
                                              uint32_t* __restrict__ e_base, uint32_t* __restrict__ e_pat) {
  const uint32_t p0 = (uint32_t)(i0 + 1);
  if (k == 1) {
    e_mass[o] = mass; e_base[o] = b; e_pat[o] = p0;
    return o + 1;
  }
  for (uint64_t m1 = cm1 & above(i0); m1; m1 &= m1 - 1) {
    const int i1 = __ffsll((long long)m1) - 1;
    const uint32_t p1 = p0 | ((uint32_t)(i1 + 1) << 8);
    if (k == 2) { e_mass[o] = mass; e_base[o] = b; e_pat[o] = p1; ++o; continue; }
    for (uint64_t m2 = cm2 & above(i1); m2; m2 &= m2 - 1) {
      const int i2 = __ffsll((long long)m2) - 1;
      const uint32_t p2 = p1 | ((uint32_t)(i2 + 1) << 16);
      if (k == 3) { e_mass[o] = mass; e_base[o] = b; e_pat[o] = p2; ++o; continue; }
      for (uint64_t m3 = cm3 & above(i2); m3; m3 &= m3 - 1) {
        const int i3 = __ffsll((long long)m3) - 1;
        e_mass[o] = mass; e_base[o] = b; e_pat[o] = p2 | ((uint32_t)(i3 + 1) << 24); ++o;
      }
    }
  }
  return o;
}

__global__ void __launch_bounds__(MD_THREADS)
    grp_expand_kernel(const uint8_t* __restrict__ res, const DevTables* __restrict__ tb, DigestCfg cfg,
                      const uint32_t* __restrict__ u_gpos, const uint16_t* __restrict__ u_len,
                      const uint64_t* __restrict__ skey, const uint64_t* __restrict__ spay,
                      const uint64_t* __restrict__ eoff, uint64_t n_groups, uint64_t base_bits,
                      uint32_t small_max, double* __restrict__ e_mass, uint32_t* __restrict__ e_base,
                      uint32_t* __restrict__ e_pat) {
  __shared__ ModTables mt;
  __shared__ WarpSites wsites[MD_WARPS];
  load_mod_tables(mt, tb);
  __syncthreads();
  const int w = threadIdx.x >> 5;
  const unsigned l = lane_id();
  WarpSites& ws = wsites[w];
  // heaviest groups (end of the mass-sorted array) first
  const uint64_t blk = gridDim.x - 1 - blockIdx.x;
  const uint64_t g = blk * MD_THREADS + threadIdx.x;
  const bool valid = g < n_groups;
  const int C = cfg.n_classes;
  uint64_t pay = 0, o = 0;
  double mass = 0;
  if (valid) {
    pay = spay[g];
    o = eoff[g];
    mass = __longlong_as_double((long long)(skey[g] + base_bits));
  }
  const uint32_t b = (uint32_t)(pay >> 32);
  const uint32_t seq = ((uint32_t)pay >> kGrpCntBits) & 31u;
  const uint32_t cnt = (uint32_t)pay & kGrpCntMask;
  uint32_t gp = 0, len = 0;
  if (valid && seq != 0) { gp = u_gpos[b]; len = u_len[b]; }
  uint32_t pk = 0;
  const int k = valid ? pack_seq(seq, C, &pk) : 0;
  // Peptides of up to 64 residues (practically all): bit i of cm_j = residue i is a site whose
  // class is the j-th class of the sequence.  Every path below enumerates with these masks.
  const bool masked = valid && k > 0 && len <= 64;
  uint64_t cm0 = 0, cm1 = 0, cm2 = 0, cm3 = 0;
  if (masked) {
    const int c0 = seq_class_at(pk, 0), c1 = k > 1 ? seq_class_at(pk, 1) : -1;
    const int c2 = k > 2 ? seq_class_at(pk, 2) : -1, c3 = k > 3 ? seq_class_at(pk, 3) : -1;
    for (uint32_t i = 0; i < len; ++i) {
      const uint8_t c = ld_res(res, gp + i);
      if (mt.flags[c] & kFlagDiffMod) {
        const int sc = mt.cls[c];
        const uint64_t bit = 1ull << i;
        if (sc == c0) cm0 |= bit;
        if (sc == c1) cm1 |= bit;
        if (sc == c2) cm2 |= bit;
        if (sc == c3) cm3 |= bit;
      }
    }
  }
  // (1) the unmodified peptide and groups of up to small_max entries: one lane each
  const bool small = valid && (k == 0 || (masked && cnt <= small_max));
  if (small) {
    if (k == 0) {
      e_mass[o] = mass; e_base[o] = b; e_pat[o] = 0;
    } else {
      for (uint64_t m0 = cm0; m0; m0 &= m0 - 1)
        o = emit_from(k, __ffsll((long long)m0) - 1, cm1, cm2, cm3, mass, b, o, e_mass, e_base, e_pat);
    }
  }
  // (2) larger groups: the warp takes one group at a time and splits it by the FIRST matched
  // site -- lane r owns the r-th site of the first class: count its occurrences, warp-scan the
  // counts into offsets, then write them
  unsigned coop = __ballot_sync(0xffffffffu, masked && !small);
  while (coop) {
    const int src = __ffs(coop) - 1;
    coop &= coop - 1;
    const uint32_t gb = __shfl_sync(0xffffffffu, b, src);
    const int gk = __shfl_sync(0xffffffffu, k, src);
    const double gmass = __shfl_sync(0xffffffffu, mass, src);
    uint64_t go = __shfl_sync(0xffffffffu, o, src);
    const uint64_t g0 = __shfl_sync(0xffffffffu, cm0, src), g1 = __shfl_sync(0xffffffffu, cm1, src);
    const uint64_t g2 = __shfl_sync(0xffffffffu, cm2, src), g3 = __shfl_sync(0xffffffffu, cm3, src);
    const int nfirst = __popcll(g0);
    for (int r0 = 0; r0 < nfirst; r0 += 32) {
      const int r = r0 + (int)l;
      const bool has = r < nfirst;
      int i0 = 0;
      if (has) {
        uint64_t m = g0;
        for (int t = 0; t < r; ++t) m &= m - 1;  // drop the r lowest set bits
        i0 = __ffsll((long long)m) - 1;
      }
      const uint32_t c = has ? count_from(gk, i0, g1, g2, g3) : 0u;
      const uint32_t inc = warp_inclusive_sum(c);
      if (has && c) emit_from(gk, i0, g1, g2, g3, gmass, gb, go + (inc - c), e_mass, e_base, e_pat);
      go += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  // (3) peptides longer than 64 residues (rare): the whole warp enumerates one group at a time
  // from the site list in shared memory; the last matched site is searched by all lanes in
  // parallel, the prefix sites by a warp-uniform odometer
  unsigned big = __ballot_sync(0xffffffffu, valid && k > 0 && !masked);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const uint32_t bb = __shfl_sync(0xffffffffu, b, src);
    const uint32_t bseq = __shfl_sync(0xffffffffu, seq, src);
    const uint32_t bgp = __shfl_sync(0xffffffffu, gp, src);
    const uint32_t blen = __shfl_sync(0xffffffffu, len, src);
    const double bmass = __shfl_sync(0xffffffffu, mass, src);
    uint64_t bo = __shfl_sync(0xffffffffu, o, src);
    uint32_t bpk;
    const int bk = pack_seq(bseq, C, &bpk);
    bool bad = false;
    const int n = warp_collect_sites(res, bgp, blen, mt, ws, &bad);
    // odometer over the first bk-1 sites, all lanes in lockstep.  State in registers only:
    // idxp byte L = site ordinal chosen at level L (0xff = none yet), pat = pattern of the prefix.
    uint32_t idxp = 0xffffffffu;
    uint32_t pat = 0;
    int level = 0;
    while (level >= 0) {
      const int want = seq_class_at(bpk, level);
      const int prev = level > 0 ? (int)((idxp >> (8 * (level - 1))) & 0xffu) : -1;  // always set when level > 0
      if (level == bk - 1) {
        // last element: the lanes scan the sites after the prefix for class `want`
        const uint32_t prefix = pat & low_bytes_mask(level);
        for (int j0 = prev + 1; j0 < n; j0 += 32) {
          const int j = j0 + (int)l;
          const bool ok = j < n && (int)mt.cls[ws.res[j]] == want;
          const unsigned om = __ballot_sync(0xffffffffu, ok);
          if (ok) {
            const uint64_t slot = bo + __popc(om & lanemask_lt());
            e_mass[slot] = bmass;
            e_base[slot] = bb;
            e_pat[slot] = prefix | (((uint32_t)ws.pos[j] + 1u) << (8 * level));
          }
          bo += __popc(om);
        }
        --level;
        continue;
      }
      // advance this prefix level to its next site of class `want`
      const uint32_t cur = (idxp >> (8 * level)) & 0xffu;
      int j = (cur != 0xffu ? (int)cur : prev) + 1;
      while (j < n && (int)mt.cls[ws.res[j]] != want) ++j;
      if (j >= n) {
        idxp |= 0xffu << (8 * level);  // exhausted: reset and go up
        --level;
      } else {
        idxp = (idxp & ~(0xffu << (8 * level))) | ((uint32_t)j << (8 * level));
        pat = (pat & low_bytes_mask(level)) | (((uint32_t)ws.pos[j] + 1u) << (8 * level));
        ++level;
        if (level < DBI_MAX_MODS_PER_PEP) idxp |= 0xffu << (8 * level);  // the next level starts fresh
      }
    }
    __syncwarp();
  }
}

}  // namespace

void launch_mod_count(const uint8_t* d_res, const DevTables* d_tb, const DigestCfg& cfg, const double* u_mass,
                      const uint32_t* u_gpos, const uint16_t* u_len, uint64_t n_unique, uint32_t tile0,
                      uint32_t ntiles, uint32_t* counts, uint32_t* tile_counts, uint32_t* d_err, cudaStream_t s) {
  if (n_unique == 0 || ntiles == 0) return;
  DBI_LAUNCH(mod_count_kernel, ntiles, MD_THREADS, 0, s, d_res, d_tb, cfg, u_mass, u_gpos, u_len, n_unique, tile0,
             counts, tile_counts, d_err);
}

void launch_mod_emit(const uint8_t* d_res, const DevTables* d_tb, const DigestCfg& cfg, const double* u_mass,
                     const uint32_t* u_gpos, const uint16_t* u_len, uint64_t n_unique, uint32_t tile0,
                     uint32_t ntiles, const uint32_t* counts, const uint64_t* tile_offs, uint64_t base_bits,
                     uint64_t* v_key, uint64_t* v_payload, cudaStream_t s) {
  if (n_unique == 0 || ntiles == 0) return;
  DBI_LAUNCH(mod_emit_kernel, ntiles, MD_THREADS, 0, s, d_res, d_tb, cfg, u_mass, u_gpos, u_len, n_unique, tile0,
             counts, tile_offs, base_bits, v_key, v_payload);
}

void launch_split_entries(const uint64_t* skey, const uint64_t* spayload, uint64_t n, uint64_t base_bits,
                          double* e_mass, uint32_t* e_base, uint32_t* e_pat, cudaStream_t s) {
  if (n == 0) return;
  const unsigned grid = (unsigned)((n + MD_THREADS - 1) / MD_THREADS);
  DBI_LAUNCH(split_entries_kernel, grid, MD_THREADS, 0, s, skey, spayload, n, base_bits, e_mass,
             e_base, e_pat);
}

void launch_grp_count(const uint8_t* d_res, const DevTables* d_tb, const DigestCfg& cfg, const double* u_mass,
                      const uint32_t* u_gpos, const uint16_t* u_len, uint64_t n_unique, uint32_t tile0,
                      uint32_t ntiles, uint8_t* ng, uint32_t* tile_groups, uint32_t* tile_vars, uint32_t* d_err,
                      cudaStream_t s) {
  if (n_unique == 0 || ntiles == 0) return;
  DBI_LAUNCH(grp_count_kernel, ntiles, MD_THREADS, 0, s, d_res, d_tb, cfg, u_mass, u_gpos, u_len, n_unique, tile0, ng,
             tile_groups, tile_vars, d_err);
}

void launch_grp_emit(const uint8_t* d_res, const DevTables* d_tb, const DigestCfg& cfg, const double* u_mass,
                     const uint32_t* u_gpos, const uint16_t* u_len, uint64_t n_unique, uint32_t tile0,
                     uint32_t ntiles, const uint8_t* ng, const uint64_t* tile_goffs, uint64_t base_bits,
                     uint64_t* g_key, uint64_t* g_pay, cudaStream_t s) {
  if (n_unique == 0 || ntiles == 0) return;
  DBI_LAUNCH(grp_emit_kernel, ntiles, MD_THREADS, 0, s, d_res, d_tb, cfg, u_mass, u_gpos, u_len, n_unique, tile0, ng,
             tile_goffs, base_bits, g_key, g_pay);
}

void launch_grp_extract_cnt(const uint64_t* pay, uint64_t n, uint32_t* cnt, cudaStream_t s) {
  if (n == 0) return;
  DBI_LAUNCH(grp_extract_cnt_kernel, (unsigned)((n + MD_THREADS - 1) / MD_THREADS), MD_THREADS, 0, s, pay, n, cnt);
}

void launch_grp_expand(const uint8_t* d_res, const DevTables* d_tb, const DigestCfg& cfg, const uint32_t* u_gpos,
                       const uint16_t* u_len, const uint64_t* skey, const uint64_t* spay, const uint64_t* eoff,
                       uint64_t n_groups, uint64_t base_bits, double* e_mass, uint32_t* e_base, uint32_t* e_pat,
                       cudaStream_t s) {
  if (n_groups == 0) return;
  static const uint32_t small_max = [] {  // DBI_GX_SMALL: diagnostic override of the lane/warp split
    const char* e = std::getenv("DBI_GX_SMALL");
    return e ? (uint32_t)std::strtoul(e, nullptr, 10) : 16u;
  }();
  DBI_LAUNCH(grp_expand_kernel, (unsigned)((n_groups + MD_THREADS - 1) / MD_THREADS), MD_THREADS, 0, s, d_res, d_tb,
             cfg, u_gpos, u_len, skey, spay, eoff, n_groups, base_bits, small_max, e_mass, e_base, e_pat);
}

}  // namespace dbi
