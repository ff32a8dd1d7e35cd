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

// Host -> device copy of a caller's (pageable) array ("pinned, chunked H2D").  cudaMemcpyAsync from
// pageable memory goes through the driver's own small staging buffer at ~11 GB/s (100M triangles: 7.6 GB in 675 ms).
// Above LISA_UPLOAD_CHUNKED_MIN bytes (default 1 GB: below that the one-time set-up of the pinned ring costs more than it
// saves) the copy is pipelined instead: four worker threads memcpy 16 MB
// chunks into their own pinned buffers (8 x 16 MB, allocated once per process) and issue the DMA of each on `st` while
// the engine drains the previous ones; returns when the copy is complete.  Smaller copies are one plain cudaMemcpyAsync
// (asynchronous only for pinned sources, as usual).
cudaError_t upload_async(void* dst, const void* src, size_t bytes, cudaStream_t st);

}  // namespace lisa
