// mg.cu -- device helpers of the multi-GPU exchange.
//
// The reference already shards its index by mass: DBIndexStoreSQLiteMult splits the mass axis
// into `indexFactor` equal-width buckets, one SQLite file each (DBIndexStoreSQLiteMult.java:55-56,
// 215-217) and answers a query from the buckets its range touches (:333-343).  Here a bucket is
// a GPU, the bucket edges are equal-COUNT splitters taken from a key histogram (the mass density
// is far from uniform), and records travel to their bucket with an NCCL all-to-all.
#include "kernels.cuh"

namespace dbi {
namespace {

constexpr int MG_THREADS = 256;

__global__ void __launch_bounds__(MG_THREADS)
    mg_hist_kernel(const uint64_t* __restrict__ key, uint64_t n, uint64_t sub, int shift,
                   const uint64_t* __restrict__ wpay, uint64_t wmask, unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[kMgBins];
  for (int i = threadIdx.x; i < kMgBins; i += MG_THREADS) sh[i] = 0;
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * MG_THREADS;
  for (uint64_t i = (uint64_t)blockIdx.x * MG_THREADS + threadIdx.x; i < n; i += stride) {
    uint64_t b = (key[i] - sub) >> shift;
    if (b >= kMgBins) b = kMgBins - 1;
    atomicAdd(&sh[b], wpay ? (uint32_t)(wpay[i] & wmask) : 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMgBins; i += MG_THREADS)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

__global__ void __launch_bounds__(MG_THREADS)
    mg_dest_kernel(const uint64_t* __restrict__ key, uint64_t n, uint64_t sub, const uint64_t* __restrict__ thr,
                   int n_thr, uint32_t* __restrict__ dest, uint32_t* __restrict__ idx,
                   unsigned long long* __restrict__ counts) {
  __shared__ uint64_t s_thr[64];
  __shared__ uint32_t s_cnt[64];
  if (threadIdx.x < 64) {
    s_thr[threadIdx.x] = (int)threadIdx.x < n_thr ? thr[threadIdx.x] : ~0ull;
    s_cnt[threadIdx.x] = 0;
  }
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * MG_THREADS + threadIdx.x;
  if (i < n) {
    const uint64_t k = key[i] - sub;
    uint32_t d = 0;
    for (int t = 0; t < n_thr; ++t) d += (s_thr[t] <= k) ? 1u : 0u;  // thresholds ascending
    dest[i] = d;
    idx[i] = (uint32_t)i;
    atomicAdd(&s_cnt[d], 1u);
  }
  __syncthreads();
  if ((int)threadIdx.x <= n_thr && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

template <typename T>
__global__ void __launch_bounds__(MG_THREADS)
    gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, T* __restrict__ dst) {
  const uint64_t i = (uint64_t)blockIdx.x * MG_THREADS + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(MG_THREADS)
    plo_to_counts_kernel(const uint64_t* __restrict__ plo, uint64_t n, uint32_t* __restrict__ cnt) {
  const uint64_t i = (uint64_t)blockIdx.x * MG_THREADS + threadIdx.x;
  if (i < n) cnt[i] = (uint32_t)(plo[i + 1] - plo[i]);
}

constexpr int FS_IPT = kScanTile / MG_THREADS;

__global__ void __launch_bounds__(MG_THREADS)
    tile_sums_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t scratch[MG_THREADS / 32 + 1];
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t sum = 0;
  for (int k = 0; k < FS_IPT; ++k) {
    const uint64_t i = tile_base + (uint64_t)k * MG_THREADS + threadIdx.x;
    if (i < n) sum += in[i];
  }
  uint32_t total;
  block_exclusive_sum<uint32_t, MG_THREADS>(sum, scratch, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(MG_THREADS)
    tile_scan_kernel(const uint32_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ tile_offs,
                     uint64_t ntiles, uint64_t* __restrict__ offs) {
  __shared__ uint32_t scratch[MG_THREADS / 32 + 1];
  const uint64_t tile_base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t running = tile_offs[blockIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) offs[n] = tile_offs[ntiles];
  for (int k = 0; k < FS_IPT; ++k) {
    const uint64_t i = tile_base + (uint64_t)k * MG_THREADS + threadIdx.x;
    const uint32_t v = (i < n) ? in[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_sum<uint32_t, MG_THREADS>(v, scratch, &total);
    if (i < n) offs[i] = running + ex;
    running += total;
  }
}

}  // namespace

void launch_mg_hist(const uint64_t* key, uint64_t n, uint64_t sub, int shift, const uint64_t* wpay, uint64_t wmask,
                    unsigned long long* hist, cudaStream_t s) {
  if (n == 0) return;
  uint64_t g = (n + MG_THREADS * 8 - 1) / (MG_THREADS * 8);
  if (g > (uint64_t)kNumSMsB200 * 8) g = (uint64_t)kNumSMsB200 * 8;
  DBI_LAUNCH(mg_hist_kernel, (unsigned)g, MG_THREADS, 0, s, key, n, sub, shift, wpay, wmask, hist);
}

void launch_mg_dest(const uint64_t* key, uint64_t n, uint64_t sub, const uint64_t* thresholds, int n_thr,
                    uint32_t* dest, uint32_t* idx, unsigned long long* counts, cudaStream_t s) {
  if (n == 0) return;
  DBI_LAUNCH(mg_dest_kernel, (unsigned)((n + MG_THREADS - 1) / MG_THREADS), MG_THREADS, 0, s, key, n, sub, thresholds,
             n_thr, dest, idx, counts);
}

void launch_gather_u64(const uint64_t* src, const uint32_t* idx, uint64_t n, uint64_t* dst, cudaStream_t s) {
  if (n == 0) return;
  DBI_LAUNCH(gather_kernel<uint64_t>, (unsigned)((n + MG_THREADS - 1) / MG_THREADS), MG_THREADS, 0, s, src, idx, n, dst);
}
void launch_gather_u32(const uint32_t* src, const uint32_t* idx, uint64_t n, uint32_t* dst, cudaStream_t s) {
  if (n == 0) return;
  DBI_LAUNCH(gather_kernel<uint32_t>, (unsigned)((n + MG_THREADS - 1) / MG_THREADS), MG_THREADS, 0, s, src, idx, n, dst);
}
void launch_gather_u16(const uint16_t* src, const uint32_t* idx, uint64_t n, uint16_t* dst, cudaStream_t s) {
  if (n == 0) return;
  DBI_LAUNCH(gather_kernel<uint16_t>, (unsigned)((n + MG_THREADS - 1) / MG_THREADS), MG_THREADS, 0, s, src, idx, n, dst);
}

void launch_plo_to_counts(const uint64_t* plo, uint64_t n, uint32_t* cnt, cudaStream_t s) {
  if (n == 0) return;
  DBI_LAUNCH(plo_to_counts_kernel, (unsigned)((n + MG_THREADS - 1) / MG_THREADS), MG_THREADS, 0, s, plo, n, cnt);
}

size_t full_scan_tmp_bytes(uint64_t n) {
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  return (tiles + 1) * 8 + tiles * 4 + 64;
}

void launch_full_scan_u32_to_u64(const uint32_t* in, uint64_t n, uint64_t* offs, void* tmp, cudaStream_t s) {
  if (n == 0) {
    DBI_CUDA(cudaMemsetAsync(offs, 0, 8, s));
    return;
  }
  const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
  uint64_t* tile_offs = (uint64_t*)tmp;
  uint32_t* tile_sums = (uint32_t*)((uint8_t*)tmp + (tiles + 1) * 8);
  DBI_LAUNCH(tile_sums_kernel, (unsigned)tiles, MG_THREADS, 0, s, in, n, tile_sums);
  launch_scan_u32_to_u64(tile_sums, tiles, tile_offs, s);
  DBI_LAUNCH(tile_scan_kernel, (unsigned)tiles, MG_THREADS, 0, s, in, n, tile_offs, tiles, offs);
}

}  // namespace dbi
