// dedup.cu -- K8: merge equal peptides after the sort.
//
// Restates DBIndexStoreSQLiteByteIndexMerge.getMergedData (Merge:620-719): records
// with the same peptide STRING collapse into one entry that keeps the first
// occurrence's (mass, offset, length) and the protein ids of every occurrence in
// insertion order, duplicates kept (Merge:658-663,678-687).
//
// The records arrive sorted by (mass bits, sequence hash, emission ordinal), so the
// occurrences of one string are adjacent and in insertion order; equal mass does
// NOT imply equal string (permutation isomers), hence the residue comparison.
// Two different strings with equal mass bits AND equal hash would interleave; that case is
// detected here and the caller re-sorts with another seed and a wider hash.  The number of
// distinct-string pairs with equal mass grows with the square of the record count, so the hash
// is 32 bits for small builds and up to 64 bits (two independent 32-bit halves) for large ones.
#include "kernels.cuh"

namespace dbi {
namespace {

constexpr int DD_THREADS = 256;
constexpr int DD_IPT = kScanTile / DD_THREADS;

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

// WIDE: hash is uint64_t[n] (low half = the 32-bit hash, high half = a second, independent one)
template <bool WIDE>
__global__ void __launch_bounds__(DD_THREADS)
    hash_records_kernel(const uint8_t* __restrict__ res, const uint32_t* __restrict__ gpos,
                        const uint16_t* __restrict__ len, uint64_t n, uint32_t seed, void* __restrict__ hash,
                        uint32_t* __restrict__ idx) {
  const uint64_t i = (uint64_t)blockIdx.x * DD_THREADS + threadIdx.x;
  if (i >= n) return;
  const uint32_t g = gpos[i];
  const uint32_t l = len[i];
  uint32_t h = 0x811c9dc5u ^ seed;
  uint32_t h2 = 0x9747b28cu + seed;
  for (uint32_t k = 0; k < l; ++k) {
    const uint32_t c = ld_res(res, g + k);
    h = (h ^ c) * 16777619u;
    if (WIDE) h2 = (h2 + c + 1u) * 0x9e3779b1u ^ (h2 >> 15);
  }
  const uint32_t lo = fmix32(h ^ (l * 0x9e3779b1u));
  if (WIDE) ((uint64_t*)hash)[i] = ((uint64_t)fmix32(h2 ^ l) << 32) | lo;
  else ((uint32_t*)hash)[i] = lo;
  idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(DD_THREADS)
    gather_mass_key_kernel(const uint64_t* __restrict__ mass_bits, const uint32_t* __restrict__ idx, uint64_t n,
                           uint64_t base_bits, uint64_t* __restrict__ key) {
  const uint64_t i = (uint64_t)blockIdx.x * DD_THREADS + threadIdx.x;
  if (i < n) key[i] = mass_bits[idx[i]] - base_bits;
}

template <typename H>
__global__ void __launch_bounds__(DD_THREADS)
    dedup_flags_kernel(const uint8_t* __restrict__ res, const uint64_t* __restrict__ skey,
                       const uint32_t* __restrict__ sidx, const H* __restrict__ hash, H hash_mask,
                       const uint32_t* __restrict__ gpos, const uint16_t* __restrict__ len, uint64_t n,
                       uint8_t* __restrict__ flags, uint32_t* __restrict__ tile_counts, uint32_t* err) {
  __shared__ uint32_t scratch[DD_THREADS / 32 + 1];
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t heads = 0;
#pragma unroll 4
  for (int k = 0; k < DD_IPT; ++k) {
    const uint64_t i = tile_base + (uint64_t)k * DD_THREADS + threadIdx.x;
    if (i >= n) break;
    uint8_t head = 1;
    if (i > 0 && skey[i] == skey[i - 1]) {
      const uint32_t a = sidx[i], b = sidx[i - 1];
      if (((hash[a] ^ hash[b]) & hash_mask) == 0) {  // only the sorted bits of the hash group the run
        const uint32_t la = len[a], lb = len[b];
        bool same = la == lb;
        if (same) {
          const uint32_t ga = gpos[a], gb = gpos[b];
          if (ga != gb)
            for (uint32_t q = 0; q < la; ++q)
              if (ld_res(res, ga + q) != ld_res(res, gb + q)) { same = false; break; }
        }
        if (same) head = 0;
        else atomicOr(err, kErrHashCollision);  // different strings, same (mass, hash)
      }
    }
    flags[i] = head;
    heads += head;
  }
  uint32_t total;
  block_exclusive_sum<uint32_t, DD_THREADS>(heads, scratch, &total);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(DD_THREADS)
    dedup_emit_kernel(const uint64_t* __restrict__ skey, const uint32_t* __restrict__ sidx,
                      const uint8_t* __restrict__ flags, const uint64_t* __restrict__ tile_offs,
                      const uint32_t* __restrict__ gpos, const uint32_t* __restrict__ prot,
                      const uint16_t* __restrict__ len, uint64_t n, uint64_t base_bits, uint64_t n_unique,
                      double* __restrict__ u_mass, uint32_t* __restrict__ u_gpos, uint32_t* __restrict__ u_prot,
                      uint16_t* __restrict__ u_len, uint64_t* __restrict__ u_plo, uint32_t* __restrict__ plist) {
  __shared__ uint32_t scratch[DD_THREADS / 32 + 1];
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t running = tile_offs[blockIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) u_plo[n_unique] = n;
  for (int k = 0; k < DD_IPT; ++k) {
    const uint64_t i = tile_base + (uint64_t)k * DD_THREADS + threadIdx.x;
    const bool valid = i < n;
    const uint32_t head = valid ? flags[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_sum<uint32_t, DD_THREADS>(head, scratch, &total);
    if (valid) {
      const uint32_t r = sidx[i];
      const uint32_t pr = prot[r];
      plist[i] = pr;  // protein ids of all occurrences, insertion order (Merge:678-681)
      if (head) {     // first occurrence supplies mass / offset / length (Merge:684-687)
        const uint64_t u = running + ex;
        u_mass[u] = __longlong_as_double((long long)(skey[i] + base_bits));
        u_gpos[u] = gpos[r];
        u_prot[u] = pr;
        u_len[u] = len[r];
        u_plo[u] = i;
      }
    }
    running += total;
  }
}

}  // namespace

void launch_hash_records(const uint8_t* d_res, const uint32_t* gpos, const uint16_t* len, uint64_t n, uint32_t seed,
                         bool wide, void* hash, uint32_t* idx, cudaStream_t s) {
  if (n == 0) return;
  const unsigned grid = (unsigned)((n + DD_THREADS - 1) / DD_THREADS);
  if (wide) DBI_LAUNCH(hash_records_kernel<true>, grid, DD_THREADS, 0, s, d_res, gpos, len, n, seed, hash, idx);
  else DBI_LAUNCH(hash_records_kernel<false>, grid, DD_THREADS, 0, s, d_res, gpos, len, n, seed, hash, idx);
}

void launch_gather_mass_key(const uint64_t* mass_bits, const uint32_t* idx, uint64_t n, uint64_t base_bits,
                            uint64_t* key, cudaStream_t s) {
  if (n == 0) return;
  const unsigned grid = (unsigned)((n + DD_THREADS - 1) / DD_THREADS);
  DBI_LAUNCH(gather_mass_key_kernel, grid, DD_THREADS, 0, s, mass_bits, idx, n, base_bits, key);
}

void launch_dedup_flags(const uint8_t* d_res, const uint64_t* skey, const uint32_t* sidx, const void* hash,
                        int hash_bits, const uint32_t* gpos, const uint16_t* len, uint64_t n, uint8_t* flags,
                        uint32_t* tile_counts, uint32_t* d_err, cudaStream_t s) {
  if (n == 0) return;
  const unsigned tiles = (unsigned)((n + kScanTile - 1) / kScanTile);
  if (hash_bits > 32) {
    const uint64_t mask = hash_bits >= 64 ? ~0ull : ((1ull << hash_bits) - 1);
    DBI_LAUNCH(dedup_flags_kernel<uint64_t>, tiles, DD_THREADS, 0, s, d_res, skey, sidx, (const uint64_t*)hash, mask,
               gpos, len, n, flags, tile_counts, d_err);
  } else {
    DBI_LAUNCH(dedup_flags_kernel<uint32_t>, tiles, DD_THREADS, 0, s, d_res, skey, sidx, (const uint32_t*)hash,
               0xffffffffu, gpos, len, n, flags, tile_counts, d_err);
  }
}

void launch_dedup_emit(const uint64_t* skey, const uint32_t* sidx, const uint8_t* flags, const uint64_t* tile_offs,
                       const uint32_t* gpos, const uint32_t* prot, const uint16_t* len, uint64_t n,
                       uint64_t base_bits, uint64_t n_unique, double* u_mass, uint32_t* u_gpos, uint32_t* u_prot,
                       uint16_t* u_len, uint64_t* u_plo, uint32_t* plist, cudaStream_t s) {
  if (n == 0) return;
  const unsigned tiles = (unsigned)((n + kScanTile - 1) / kScanTile);
  DBI_LAUNCH(dedup_emit_kernel, tiles, DD_THREADS, 0, s, skey, sidx, flags, tile_offs, gpos, prot, len, n, base_bits,
             n_unique, u_mass, u_gpos, u_prot, u_len, u_plo, plist);
}

}  // namespace dbi
