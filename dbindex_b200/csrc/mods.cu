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

#include "mods_common.cuh"

namespace dbi {
namespace {

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
                    uint64_t id_off, uint64_t* __restrict__ v_key, uint64_t* __restrict__ v_payload) {
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
        v_payload[out] = (id_off + u) << 32;  // peptides are named by global id (id_off = 0 on a single GPU)
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
        v_payload[slot] = ((id_off + u) << 32) | pat;
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
                     uint64_t id_off, uint64_t* v_key, uint64_t* v_payload, cudaStream_t s) {
  if (n_unique == 0 || ntiles == 0) return;
  DBI_LAUNCH(mod_emit_kernel, ntiles, MD_THREADS, 0, s, d_res, d_tb, cfg, u_mass, u_gpos, u_len, n_unique, tile0,
             counts, tile_offs, base_bits, id_off, v_key, v_payload);
}

void launch_split_entries(const uint64_t* skey, const uint64_t* spayload, uint64_t n, uint64_t base_bits,
                          double* e_mass, uint32_t* e_base, uint32_t* e_pat, cudaStream_t s) {
  if (n == 0) return;
  const unsigned grid = (unsigned)((n + MD_THREADS - 1) / MD_THREADS);
  DBI_LAUNCH(split_entries_kernel, grid, MD_THREADS, 0, s, skey, spayload, n, base_bits, e_mass,
             e_base, e_pat);
}

}  // namespace dbi
