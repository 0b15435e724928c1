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
// Two-pass count / scan / emit like the digestion.
#include "kernels.cuh"

namespace dbi {
namespace {

constexpr int MD_THREADS = 256;
constexpr int MD_IPT = kScanTile / MD_THREADS;
constexpr int MD_MAX_SITES = DBI_MAX_MOD_POS + 1;  // positions 0..254

struct ModTables {
  double diff[256];
  uint8_t flags[256];
};

__device__ __forceinline__ void load_mod_tables(ModTables& mt, const DevTables* __restrict__ tb) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    mt.diff[i] = tb->diff[i];
    mt.flags[i] = tb->flags[i];
  }
}

// Collect eligible site positions of the peptide at res[g .. g+len).  Returns the
// number of sites; *bad is set if a site lies beyond DBI_MAX_MOD_POS.
__device__ __forceinline__ int collect_sites(const uint8_t* __restrict__ res, uint32_t g, uint32_t len,
                                             const ModTables& mt, uint8_t* sites, bool* bad) {
  int n = 0;
  for (uint32_t i = 0; i < len; ++i) {
    if (mt.flags[ld_res(res, g + i)] & kFlagDiffMod) {
      if (i > DBI_MAX_MOD_POS) { *bad = true; break; }
      sites[n++] = (uint8_t)i;
    }
  }
  return n;
}

// Enumerate the modified variants (k >= 1) in SPEC order; F(mass, pattern) per variant
// that passes the mass gate.  Returns how many passed.
template <typename F>
__device__ __forceinline__ uint32_t enumerate_variants(const uint8_t* __restrict__ res, uint32_t g, double base_mass,
                                                       const uint8_t* sites, int n, const ModTables& mt,
                                                       const DigestCfg& cfg, F&& f) {
  uint32_t count = 0;
  int idx[DBI_MAX_MODS_PER_PEP];
  const int K = cfg.max_mods;
  for (int k = 1; k <= K && k <= n; ++k) {
    for (int i = 0; i < k; ++i) idx[i] = i;
    while (true) {
      double m = base_mass;
      uint32_t pat = 0;
      for (int i = 0; i < k; ++i) {
        const uint32_t p = sites[idx[i]];
        m = __dadd_rn(m, mt.diff[ld_res(res, g + p)]);
        pat |= (p + 1) << (8 * i);
      }
      if (m >= cfg.min_mass && m <= cfg.max_mass) {
        f(m, pat);
        ++count;
      }
      int i = k - 1;
      while (i >= 0 && idx[i] == n - k + i) --i;
      if (i < 0) break;
      ++idx[i];
      for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
    }
  }
  return count;
}

__global__ void __launch_bounds__(MD_THREADS)
    mod_count_kernel(const uint8_t* __restrict__ res, const DevTables* __restrict__ tb, DigestCfg cfg,
                     const double* __restrict__ u_mass, const uint32_t* __restrict__ u_gpos,
                     const uint16_t* __restrict__ u_len, uint64_t n_unique, uint32_t* __restrict__ counts,
                     uint32_t* __restrict__ tile_counts, uint32_t* err) {
  __shared__ ModTables mt;
  __shared__ uint32_t scratch[MD_THREADS / 32 + 1];
  load_mod_tables(mt, tb);
  __syncthreads();
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t sum = 0;
  uint8_t sites[MD_MAX_SITES];
  for (int k = 0; k < MD_IPT; ++k) {
    const uint64_t u = tile_base + (uint64_t)k * MD_THREADS + threadIdx.x;
    if (u >= n_unique) break;
    bool bad = false;
    const uint32_t g = u_gpos[u];
    const int n = collect_sites(res, g, u_len[u], mt, sites, &bad);
    if (bad) atomicOr(err, kErrModPos);
    const uint32_t c = 1u + enumerate_variants(res, g, u_mass[u], sites, n, mt, cfg, [](double, uint32_t) {});
    counts[u] = c;
    sum += c;
  }
  uint32_t total;
  block_exclusive_sum<uint32_t, MD_THREADS>(sum, scratch, &total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(MD_THREADS)
    mod_emit_kernel(const uint8_t* __restrict__ res, const DevTables* __restrict__ tb, DigestCfg cfg,
                    const double* __restrict__ u_mass, const uint32_t* __restrict__ u_gpos,
                    const uint16_t* __restrict__ u_len, uint64_t n_unique, const uint32_t* __restrict__ counts,
                    const uint64_t* __restrict__ tile_offs, uint64_t base_bits, uint64_t* __restrict__ v_key,
                    uint64_t* __restrict__ v_payload) {
  __shared__ ModTables mt;
  __shared__ uint32_t scratch[MD_THREADS / 32 + 1];
  load_mod_tables(mt, tb);
  __syncthreads();
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t running = tile_offs[blockIdx.x];
  uint8_t sites[MD_MAX_SITES];
  for (int k = 0; k < MD_IPT; ++k) {
    const uint64_t u = tile_base + (uint64_t)k * MD_THREADS + threadIdx.x;
    const bool valid = u < n_unique;
    const uint32_t c = valid ? counts[u] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_sum<uint32_t, MD_THREADS>(c, scratch, &total);
    if (valid) {
      uint64_t o = running + ex;
      const double bm = u_mass[u];
      const uint32_t g = u_gpos[u];
      // k = 0: the unmodified peptide
      v_key[o] = (uint64_t)__double_as_longlong(bm) - base_bits;
      v_payload[o] = u << 32;
      ++o;
      if (c > 1) {
        bool bad = false;
        const int n = collect_sites(res, g, u_len[u], mt, sites, &bad);
        enumerate_variants(res, g, bm, sites, n, mt, cfg, [&](double m, uint32_t pat) {
          v_key[o] = (uint64_t)__double_as_longlong(m) - base_bits;
          v_payload[o] = (u << 32) | pat;
          ++o;
        });
      }
    }
    running += total;
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
                      const uint32_t* u_gpos, const uint16_t* u_len, uint64_t n_unique, uint32_t* counts,
                      uint32_t* tile_counts, uint32_t* d_err, cudaStream_t s) {
  if (n_unique == 0) return;
  const unsigned tiles = (unsigned)((n_unique + kScanTile - 1) / kScanTile);
  DBI_LAUNCH(mod_count_kernel, tiles, MD_THREADS, 0, s, d_res, d_tb, cfg, u_mass, u_gpos, u_len, n_unique, counts,
             tile_counts, d_err);
}

void launch_mod_emit(const uint8_t* d_res, const DevTables* d_tb, const DigestCfg& cfg, const double* u_mass,
                     const uint32_t* u_gpos, const uint16_t* u_len, uint64_t n_unique, const uint32_t* counts,
                     const uint64_t* tile_offs, uint64_t base_bits, uint64_t* v_key, uint64_t* v_payload,
                     cudaStream_t s) {
  if (n_unique == 0) return;
  const unsigned tiles = (unsigned)((n_unique + kScanTile - 1) / kScanTile);
  DBI_LAUNCH(mod_emit_kernel, tiles, MD_THREADS, 0, s, d_res, d_tb, cfg, u_mass, u_gpos, u_len, n_unique, counts,
             tile_offs, base_bits, v_key, v_payload);
}

void launch_split_entries(const uint64_t* skey, const uint64_t* spayload, uint64_t n, uint64_t base_bits,
                          double* e_mass, uint32_t* e_base, uint32_t* e_pat, cudaStream_t s) {
  if (n == 0) return;
  const unsigned grid = (unsigned)((n + MD_THREADS - 1) / MD_THREADS);
  DBI_LAUNCH(split_entries_kernel, grid, MD_THREADS, 0, s, skey, spayload, n, base_bits, e_mass, e_base, e_pat);
}

}  // namespace dbi
