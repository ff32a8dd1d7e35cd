// lisa_b200/csrc/devmem.h — process-wide caching allocator for device (and small pinned host) memory.
//
// cudaFree synchronises the device and was measured to take anywhere from 0.6 ms to 540 ms on the B200 boxes
// (driver-side unmapping), cudaMalloc of the 64 MB state arrays several ms each.  A create -> render -> read ->
// destroy cycle through the C ABI would be dominated by that, so every allocation of the library goes through
// this cache: blocks are rounded to size classes, freed blocks are parked (bounded) and reused by the next
// context or BVH build; memory is returned to the driver only when the cache is over its budget or an
// allocation fails.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace lisa {

cudaError_t dev_alloc(void** out, size_t bytes);  // on the current device
void        dev_free(void* p);                    // accepts nullptr
void        dev_cache_trim();                     // give everything cached back to the driver

template <typename T>
inline cudaError_t dev_alloc_t(T** out, size_t count) { return dev_alloc(reinterpret_cast<void**>(out), sizeof(T) * (count ? count : 1)); }

// 256-byte pinned host slots (counters read back every burst)
void* pinned_slot_alloc();
void  pinned_slot_free(void* p);

}  // namespace lisa
