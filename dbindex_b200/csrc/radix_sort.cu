// radix_sort.cu -- K7: stable LSD radix sort of (key, value) pairs, 8-bit digits,
// "onesweep" organisation: ONE histogram pass over the keys for all digits, then
// one scatter pass per digit in which every tile resolves its global offsets with
// a decoupled look-back over its predecessors (no separate per-pass scan kernel,
// so a pass costs exactly one read and one write of the pairs).
//
// Replaces the mass-ordered storage of the reference: the per-row
// `Collections.sort(sortedMerged)` of DBIndexStoreSQLiteByteIndexMerge.java:693 and
// the SQLite B-tree on precursor_mass_key (DBIndexStoreSQLiteByte.java:586-587,607).
//
// HBM-bound integer work, no tensor cores.  Algorithmic bytes per pass =
// n * 2 * (sizeof(K) + sizeof(V)); histogram = n * sizeof(K).
#include "radix_sort.cuh"

namespace dbi {

namespace {

constexpr int RS_BITS = 8;
constexpr int RS_RADIX = 1 << RS_BITS;
constexpr int RS_THREADS = 256;  // == RS_RADIX: thread t owns digit t in the tile-level steps
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 16;
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // 4096 pairs per tile
constexpr int RS_MAX_PASSES = 8;

// look-back word: [63:62] flag, [61:58] pass id, [57:0] count
constexpr uint64_t LB_AGG = 1ull << 62;
constexpr uint64_t LB_INC = 2ull << 62;
constexpr uint64_t LB_VAL_MASK = (1ull << 58) - 1;

// ---- histogram of every digit in one read of the keys ----------------------
template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_histogram_kernel(const K* __restrict__ keys, uint64_t n,
                                                                 int begin_bit, int end_bit, int num_passes,
                                                                 unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[RS_MAX_PASSES * RS_RADIX];
  for (int i = threadIdx.x; i < num_passes * RS_RADIX; i += RS_THREADS) sh[i] = 0;
  __syncthreads();
  constexpr int HU = 4;  // independent key loads in flight per thread
  const uint64_t stride = (uint64_t)gridDim.x * RS_THREADS * HU;
  // whole warps iterate together so that match_any sees a full mask
  const uint64_t n_round = (n + 31) & ~31ull;
  // i0 - lane and n_round are multiples of 32, so the loop condition is warp-uniform
  for (uint64_t i0 = (uint64_t)blockIdx.x * RS_THREADS * HU + threadIdx.x; i0 < n_round; i0 += stride) {
    K key[HU];
    bool valid[HU];
#pragma unroll
    for (int u = 0; u < HU; ++u) {
      const uint64_t i = i0 + (uint64_t)u * RS_THREADS;
      valid[u] = i < n;
      key[u] = valid[u] ? keys[i] : K(0);
    }
#pragma unroll
    for (int u = 0; u < HU; ++u) {
      for (int p = 0; p < num_passes; ++p) {
        const int shift = begin_bit + p * RS_BITS;
        const int bits = min(RS_BITS, end_bit - shift);
        const uint32_t d = (uint32_t)(key[u] >> shift) & ((1u << bits) - 1);
        // high digits of neighbouring keys are mostly equal (the input is emitted in near mass
        // order): one atomic for the whole warp then; otherwise plain shared atomics, which
        // rarely conflict on the well-mixed low digits
        const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
        const unsigned vm = __ballot_sync(0xffffffffu, valid[u]);
        if (vm == 0xffffffffu && __all_sync(0xffffffffu, d == d0)) {
          if (lane_id() == 0) atomicAdd(&sh[p * RS_RADIX + d], 32u);
        } else if (valid[u]) {
          atomicAdd(&sh[p * RS_RADIX + d], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < num_passes * RS_RADIX; i += RS_THREADS)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// exclusive scan of each pass's 256 bins -> global digit offsets
__global__ void __launch_bounds__(RS_RADIX) rs_scan_hist_kernel(unsigned long long* __restrict__ hist) {
  __shared__ unsigned long long scratch[RS_RADIX / 32 + 1];
  unsigned long long* h = hist + (size_t)blockIdx.x * RS_RADIX;
  const unsigned long long v = h[threadIdx.x];
  unsigned long long total;
  const unsigned long long ex = block_exclusive_sum<unsigned long long, RS_RADIX>(v, scratch, &total);
  h[threadIdx.x] = ex;
}

// ---- one scatter pass ------------------------------------------------------
template <typename K, typename V>
struct RsSmem {
  union {  // keys are staged and written out first, then the values reuse the space
    K keys[RS_TILE];
    V vals[RS_TILE];
  } st;
  uint8_t sdig[RS_TILE];                // digit of the key in each local slot
  uint32_t whist[RS_WARPS * RS_RADIX];  // per-warp digit counters, then exclusive warp prefixes
  uint64_t outbase[RS_RADIX];           // global position of local slot 0 of each digit run (mod 2^64)
  uint32_t dstart[RS_RADIX];            // first local slot of each digit
  uint32_t scratch[RS_THREADS / 32 + 1];
  uint32_t tile;
};

// SPLIT (last pass of the variant sort): keys leave as key + key_add (the mass bits), the
// 64-bit values as two u32 arrays (vout_hi = base peptide, vout_lo = mod pattern).
template <typename K, typename V, bool SPLIT>
__global__ void __launch_bounds__(RS_THREADS, 4)
    rs_onesweep_kernel(const K* __restrict__ kin, K* __restrict__ kout, const V* __restrict__ vin,
                       V* __restrict__ vout, uint64_t n, int shift, uint32_t mask,
                       const unsigned long long* __restrict__ digit_off, unsigned long long* lookback,
                       uint32_t* tile_counter, uint32_t pass_id, K key_add, uint32_t* __restrict__ vout_hi,
                       uint32_t* __restrict__ vout_lo) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  RsSmem<K, V>& s = *reinterpret_cast<RsSmem<K, V>*>(smem_raw);
  const int t = threadIdx.x;
  const int w = t >> 5;
  const unsigned l = lane_id();

  // tiles are claimed in launch order so that every predecessor of a running tile
  // is itself running or finished (forward progress of the look-back)
  if (t == 0) s.tile = atomicAdd(tile_counter, 1u);
  for (int i = t; i < RS_WARPS * RS_RADIX; i += RS_THREADS) s.whist[i] = 0;
  __syncthreads();
  const uint32_t tile = s.tile;
  const uint64_t base = (uint64_t)tile * RS_TILE;
  const uint32_t tile_n = (uint32_t)min((uint64_t)RS_TILE, n - base);

  // warp-striped load: element e = w*32*IPT + i*32 + lane  (memory order = (w, i, lane))
  K key[RS_IPT];
  uint32_t rank[RS_IPT];
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t e = (uint32_t)w * 32 * RS_IPT + i * 32 + l;
    // out-of-range slots get the all-ones key: largest digit, and being the last
    // elements of the tile they rank after every real key of that digit
    key[i] = (e < tile_n) ? kin[base + e] : (K)~(K)0;
  }

  // rank inside the warp, digit by digit group (stable: item order, then lane order)
  uint32_t* wh = s.whist + w * RS_RADIX;
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
    const unsigned m = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(m) - 1;
    uint32_t old = 0;
    if ((int)l == leader) {
      old = wh[d];
      wh[d] = old + (uint32_t)__popc(m);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[i] = old + (uint32_t)__popc(m & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();

  // thread t owns digit t: exclusive prefix over the warps, tile count, digit starts
  uint32_t sum = 0;
#pragma unroll
  for (int ww = 0; ww < RS_WARPS; ++ww) {
    const uint32_t c = s.whist[ww * RS_RADIX + t];
    s.whist[ww * RS_RADIX + t] = sum;
    sum += c;
  }
  const uint32_t pad = RS_TILE - tile_n;  // all-ones filler keys, all in digit `mask`
  const uint32_t cnt_valid = sum - (((uint32_t)t == mask) ? pad : 0u);
  uint32_t total;
  const uint32_t dstart = block_exclusive_sum<uint32_t, RS_THREADS>(sum, s.scratch, &total);
  s.dstart[t] = dstart;

  // decoupled look-back: number of keys with digit t in all earlier tiles
  const uint64_t tag = (uint64_t)(pass_id & 0xf) << 58;
  volatile unsigned long long* lb = lookback;
  uint64_t excl = 0;
  if (tile == 0) {
    lb[(size_t)tile * RS_RADIX + t] = LB_INC | tag | (uint64_t)cnt_valid;
  } else {
    lb[(size_t)tile * RS_RADIX + t] = LB_AGG | tag | (uint64_t)cnt_valid;
    int64_t j = (int64_t)tile - 1;
    while (true) {
      const uint64_t st = lb[(size_t)j * RS_RADIX + t];
      if ((st & (0xfull << 58)) != tag || (st >> 62) == 0) continue;  // not published yet
      excl += st & LB_VAL_MASK;
      if ((st >> 62) == 2) break;
      --j;
    }
    lb[(size_t)tile * RS_RADIX + t] = LB_INC | tag | (excl + cnt_valid);
  }
  s.outbase[t] = (uint64_t)digit_off[t] + excl - (uint64_t)dstart;
  __syncthreads();

  // local slot of every key; stage the keys in digit order
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
    rank[i] += s.dstart[d] + wh[d];
    s.st.keys[rank[i]] = key[i];
  }
  __syncthreads();
  // coalesced write-out: consecutive local slots of one digit are consecutive in HBM
#pragma unroll 4
  for (uint32_t j = t; j < tile_n; j += RS_THREADS) {
    const K k = s.st.keys[j];
    const uint32_t d = (uint32_t)(k >> shift) & mask;
    s.sdig[j] = (uint8_t)d;
    kout[s.outbase[d] + j] = SPLIT ? (K)(k + key_add) : k;
  }
  __syncthreads();  // every key has left the staging buffer
  // values take the same route through the same buffer
#pragma unroll
  for (int i = 0; i < RS_IPT; ++i) {
    const uint32_t e = (uint32_t)w * 32 * RS_IPT + i * 32 + l;
    if (e < tile_n) s.st.vals[rank[i]] = vin[base + e];
  }
  __syncthreads();
#pragma unroll 4
  for (uint32_t j = t; j < tile_n; j += RS_THREADS) {
    const uint64_t o = s.outbase[s.sdig[j]] + j;
    const V v = s.st.vals[j];
    if (SPLIT) {
      vout_hi[o] = (uint32_t)((uint64_t)v >> 32);
      vout_lo[o] = (uint32_t)v;
    } else {
      vout[o] = v;
    }
  }
}

}  // namespace

size_t radix_sort_tmp_bytes(uint64_t n) {
  const uint64_t tiles = (n + RS_TILE - 1) / RS_TILE;
  // histograms + tile counters + look-back words
  return RS_MAX_PASSES * RS_RADIX * 8 + 256 + (size_t)tiles * RS_RADIX * 8;
}

template <typename K, typename V>
int radix_sort_pairs(K* keys[2], V* vals[2], uint64_t n, int begin_bit, int end_bit, void* tmp,
                     cudaStream_t stream, PassProbe* probe, const SplitOut<K>* split) {
  if (n <= 1 || end_bit <= begin_bit) return 0;
  const int num_passes = (end_bit - begin_bit + RS_BITS - 1) / RS_BITS;
  if (num_passes > RS_MAX_PASSES) throw CudaError{cudaErrorInvalidValue, "radix_sort_pairs: too many passes", __FILE__, __LINE__};
  const uint64_t tiles = (n + RS_TILE - 1) / RS_TILE;
  if (tiles >= (1ull << 32)) throw CudaError{cudaErrorInvalidValue, "radix_sort_pairs: n too large", __FILE__, __LINE__};

  uint8_t* tp = (uint8_t*)tmp;
  unsigned long long* hist = (unsigned long long*)tp;
  uint32_t* counters = (uint32_t*)(tp + RS_MAX_PASSES * RS_RADIX * 8);
  unsigned long long* lookback = (unsigned long long*)(tp + RS_MAX_PASSES * RS_RADIX * 8 + 256);

  DBI_CUDA(cudaMemsetAsync(tp, 0, RS_MAX_PASSES * RS_RADIX * 8 + 256 + (size_t)tiles * RS_RADIX * 8, stream));

  uint64_t hg = (n + RS_THREADS * 4 - 1) / (RS_THREADS * 4);
  if (hg > (uint64_t)kNumSMsB200 * 8) hg = (uint64_t)kNumSMsB200 * 8;
  const int hgrid = (int)hg;
  DBI_LAUNCH((rs_histogram_kernel<K>), hgrid, RS_THREADS, 0, stream, keys[0], n, begin_bit, end_bit, num_passes, hist);
  DBI_LAUNCH(rs_scan_hist_kernel, num_passes, RS_RADIX, 0, stream, hist);

  const size_t smem = sizeof(RsSmem<K, V>);
  DBI_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<K, V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DBI_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<K, V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int cur = 0;
  for (int p = 0; p < num_passes; ++p) {
    const int shift = begin_bit + p * RS_BITS;
    const int bits = (end_bit - shift) < RS_BITS ? (end_bit - shift) : RS_BITS;
    const uint32_t mask = (1u << bits) - 1;
    if (probe) probe->before_pass(p);
    if (split && p == num_passes - 1) {
      // last pass writes the final arrays directly: key + key_add -> split->keys, value halves
      DBI_LAUNCH((rs_onesweep_kernel<K, V, true>), (unsigned)tiles, RS_THREADS, smem, stream, keys[cur], split->keys,
                 vals[cur], vals[cur ^ 1], n, shift, mask, hist + (size_t)p * RS_RADIX, lookback, counters + p,
                 (uint32_t)(p + 1), split->key_add, split->vals_hi, split->vals_lo);
    } else {
      DBI_LAUNCH((rs_onesweep_kernel<K, V, false>), (unsigned)tiles, RS_THREADS, smem, stream, keys[cur],
                 keys[cur ^ 1], vals[cur], vals[cur ^ 1], n, shift, mask, hist + (size_t)p * RS_RADIX, lookback,
                 counters + p, (uint32_t)(p + 1), (K)0, (uint32_t*)nullptr, (uint32_t*)nullptr);
    }
    if (probe) probe->after_pass(p);
    cur ^= 1;
  }
  return cur;
}

template int radix_sort_pairs<uint32_t, uint32_t>(uint32_t* [2], uint32_t* [2], uint64_t, int, int, void*, cudaStream_t, PassProbe*, const SplitOut<uint32_t>*);
template int radix_sort_pairs<uint64_t, uint32_t>(uint64_t* [2], uint32_t* [2], uint64_t, int, int, void*, cudaStream_t, PassProbe*, const SplitOut<uint64_t>*);
template int radix_sort_pairs<uint64_t, uint64_t>(uint64_t* [2], uint64_t* [2], uint64_t, int, int, void*, cudaStream_t, PassProbe*, const SplitOut<uint64_t>*);

}  // namespace dbi
