// Device-wide primitives written for this path: exclusive scan and a stable LSD radix sort of
// (key, value) uint32 pairs.  They implement the "deterministic sort-by-id" step of the sparse
// gradient reduction (BASELINE.json north_star).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

namespace prims {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

constexpr int MAX_RADIX_BITS = 10;     // digit width is chosen per sort (8, 9 or 10 bits) to minimise the passes
constexpr int MAX_RADIX = 1 << MAX_RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

inline int64_t scan_blocks(int64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }
inline int64_t sort_blocks(int64_t n) { return (n + SORT_TILE - 1) / SORT_TILE; }

// temp bytes for an exclusive scan of n elements of size elem
inline size_t scan_temp_bytes(int64_t n, size_t elem) { return (size_t)(scan_blocks(n) + 1) * elem; }
// temp bytes for radix_sort_pairs of n pairs
inline size_t sort_temp_bytes(int64_t n) {
    int64_t cnt = (int64_t)MAX_RADIX * sort_blocks(n);
    return (size_t)cnt * 4 * 2 + scan_temp_bytes(cnt, 4) + 256;
}

// out[i] = sum_{j<i} in[j]; if total != nullptr, *total = sum of all.  in may alias out.
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* temp, uint32_t* total,
                        cudaStream_t st, int64_t* launches);
void exclusive_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, void* temp, uint64_t* total,
                        cudaStream_t st, int64_t* launches);

// Stable LSD radix sort on the low `bits` bits of the keys.  keys[0]/vals[0] hold the input;
// returns the index (0/1) of the buffer pair that holds the sorted output.
int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], int64_t n, int bits, void* temp,
                     cudaStream_t st, int64_t* launches);

// One-sweep variant (decoupled look-back, 8-bit digits): one histogram launch + one launch per digit.  The element count
// may live in device memory (n_dev != nullptr; n_max then bounds it and sizes the grids).
size_t onesweep_temp_bytes(int64_t n_max);
int onesweep_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], int64_t n_max, const uint32_t* n_dev, int bits, void* temp,
                        cudaStream_t st, int64_t* launches);

}  // namespace prims
