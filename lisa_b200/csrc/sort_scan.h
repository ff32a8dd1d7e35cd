// lisa_b200/csrc/sort_scan.h — device primitives of the BVH builder: exclusive scan, LSD radix sort of
// (64-bit key, 32-bit value) pairs, stream compaction of non-negative ints.  Hand-written (no CUB/Thrust).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lisa {

// Workspace sizes are functions of n only; callers allocate once (devmem) and reuse.
size_t scan_temp_bytes(size_t n);
// out[i] = sum_{j<i} in[i]  (in == out allowed); total (optional, device pointer) receives the sum of all elements
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total, void* temp, cudaStream_t st);

size_t radix_sort_temp_bytes(size_t n);
// Stable LSD radix sort, 8 bits per pass over bits [begin_bit, end_bit).  Sorted data ends in (keys_out, vals_out)
// after an odd number of passes and in (keys_in, vals_in) after an even number; returns which: 1 = *_out, 0 = *_in.
int radix_sort_pairs(unsigned long long* keys_in, unsigned long long* keys_out, unsigned int* vals_in, unsigned int* vals_out,
                     size_t n, int begin_bit, int end_bit, void* temp, cudaStream_t st);

size_t compact_temp_bytes(size_t n);
// out = the entries of in that are >= 0, order preserved; *d_count (device) = how many
void compact_nonneg(const int* in, int* out, size_t n, int* d_count, void* temp, cudaStream_t st);

}  // namespace lisa
