// radix_sort.cuh -- host interface of the hand-written onesweep LSD radix sort (K7).
#pragma once
#include "common.cuh"

namespace dbi {

// Scratch bytes radix_sort_pairs needs for n pairs.
size_t radix_sort_tmp_bytes(uint64_t n);

// Stable sort of n (key, value) pairs on key bits [begin_bit, end_bit).
// keys[0]/vals[0] hold the input; both double buffers are clobbered.  Returns the
// index (0 or 1) of the buffer pair holding the sorted output.  Instantiated for
// (u32,u32), (u64,u32) and (u64,u64).
// Optional observer called on the host right before / after each scatter pass is
// enqueued (used to bracket the passes with CUDA events).
struct PassProbe {
  virtual void before_pass(int pass) = 0;
  virtual void after_pass(int pass) = 0;
  virtual ~PassProbe() = default;
};

// Optional fused epilogue of the LAST pass: instead of the ping-pong buffers, the sorted
// keys go to `keys` as key + key_add and the (64-bit) values to two u32 arrays.  When used
// (and the sort runs at least one pass) the returned buffer index is meaningless.
template <typename K>
struct SplitOut {
  K* keys;
  K key_add;
  uint32_t* vals_hi;
  uint32_t* vals_lo;
};

template <typename K, typename V>
int radix_sort_pairs(K* keys[2], V* vals[2], uint64_t n, int begin_bit, int end_bit, void* tmp,
                     cudaStream_t stream, PassProbe* probe, const SplitOut<K>* split = nullptr);

}  // namespace dbi
