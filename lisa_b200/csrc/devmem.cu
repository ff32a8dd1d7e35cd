// lisa_b200/csrc/devmem.cu — see devmem.h.
#include "devmem.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace lisa {
namespace {

struct Key {
  int    device;
  size_t bytes;
  bool operator<(const Key& o) const { return device != o.device ? device < o.device : bytes < o.bytes; }
};
std::mutex                                      g_mu;
std::multimap<Key, void*>                       g_free;    // parked blocks by (device, class size)
std::unordered_map<void*, Key>                  g_live;    // every block handed out
size_t                                          g_cached = 0;
// parked bytes per process: 64 GB by default (a 100M-triangle build holds ~50 GB at its peak; anything above the budget
// goes back to the driver, and the next build then pays cudaMalloc again: 20-30 ms per multi-GB block); LISA_CACHE_GB overrides
size_t budget() {
  static size_t b = [] {
    const char* e = getenv("LISA_CACHE_GB");
    return (size_t)(e ? std::max(0, atoi(e)) : 64) << 30;
  }();
  return b;
}

size_t size_class(size_t b) {
  if (b < 512) return 512;
  if (b <= (64u << 20)) {  // next power of two
    size_t c = 512;
    while (c < b) c <<= 1;
    return c;
  }
  const size_t g = 64u << 20;  // multiples of 64 MB above that
  return (b + g - 1) / g * g;
}

void release_device_locked(int device) {
  for (auto it = g_free.begin(); it != g_free.end();) {
    if (device < 0 || it->first.device == device) {
      cudaFree(it->second);
      g_cached -= it->first.bytes;
      it = g_free.erase(it);
    } else {
      ++it;
    }
  }
}

std::vector<void*> g_pinned_free;
char*              g_pinned_chunk = nullptr;

}  // namespace

cudaError_t dev_alloc(void** out, size_t bytes) {
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  const Key k{device, size_class(bytes)};
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_free.find(k);
  if (it != g_free.end()) {
    *out = it->second;
    g_cached -= k.bytes;
    g_free.erase(it);
    g_live[*out] = k;
    return cudaSuccess;
  }
  e = cudaMalloc(out, k.bytes);
  if (e != cudaSuccess) {  // out of memory: hand the cache back and retry once
    cudaGetLastError();
    release_device_locked(device);
    e = cudaMalloc(out, k.bytes);
  }
  if (e == cudaSuccess) g_live[*out] = k;
  return e;
}

void dev_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_live.find(p);
  if (it == g_live.end()) { cudaFree(p); return; }  // not ours
  const Key k = it->second;
  g_live.erase(it);
  if (g_cached + k.bytes <= budget()) {
    g_free.emplace(k, p);
    g_cached += k.bytes;
  } else {
    cudaFree(p);
  }
}

void dev_cache_trim() {
  std::lock_guard<std::mutex> lk(g_mu);
  release_device_locked(-1);
}

void* pinned_slot_alloc() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_pinned_free.empty()) {
    char* chunk = nullptr;
    if (cudaMallocHost(&chunk, 256 * 64) != cudaSuccess) return nullptr;
    g_pinned_chunk = chunk;  // never returned: 16 KB per refill
    for (int i = 0; i < 64; i++) g_pinned_free.push_back(chunk + 256 * i);
  }
  void* p = g_pinned_free.back();
  g_pinned_free.pop_back();
  return p;
}
void pinned_slot_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_mu);
  g_pinned_free.push_back(p);
}


// ---- pinned, chunked upload ------------------------------------------------------------------------
namespace {
const size_t kChunk   = 16u << 20;
const int    kWorkers = 4, kPerWorker = 2;  // 8 pinned buffers of 16 MB per process
struct UploadRing {
  std::mutex  mu;  // one upload at a time per process (the ring is shared)
  char*       buf[kWorkers * kPerWorker] = {};  // portable pinned memory: usable from any device
  bool        ok = false, tried = false;
} g_ring;
}  // namespace

cudaError_t upload_async(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  size_t min_bytes = 1u << 30;  // measured on B200: 0.76 GB 70 ms plain vs 100-145 ms chunked (pinned ring set-up), 3.6 GB x 2: 675 vs 384 ms
  if (const char* e = getenv("LISA_UPLOAD_CHUNKED_MIN")) min_bytes = (size_t)strtoull(e, nullptr, 10);
  if (bytes == 0) return cudaSuccess;
  if (bytes < min_bytes) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  std::lock_guard<std::mutex> lk(g_ring.mu);
  if (!g_ring.tried) {
    g_ring.tried = true;
    g_ring.ok = true;
    for (int k = 0; k < kWorkers * kPerWorker && g_ring.ok; k++)
      g_ring.ok = cudaHostAlloc((void**)&g_ring.buf[k], kChunk, cudaHostAllocPortable) == cudaSuccess;
    if (!g_ring.ok) (void)cudaGetLastError();  // no pinned memory to be had: plain copies from now on
  }
  if (!g_ring.ok) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  int device = 0;
  cudaGetDevice(&device);
  // an event can only be recorded on a stream of the device it was created on, and the ring serves every device of the
  // process (render_multi creates one context per GPU): the "buffer drained" events are made per call, on `st`'s device
  struct Events {
    cudaEvent_t e[kWorkers * kPerWorker] = {};
    ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
  } done;
  for (cudaEvent_t& x : done.e) {
    const cudaError_t ce = cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
    if (ce != cudaSuccess) return ce;
  }
  const char*  s = static_cast<const char*>(src);
  char*        d = static_cast<char*>(dst);
  const size_t chunks = (bytes + kChunk - 1) / kChunk;
  cudaError_t  err[kWorkers];
  // worker w stages chunks w, w + kWorkers, ... through its own two pinned buffers and issues their DMA itself: the
  // destinations are disjoint, so the order in which the chunks reach the stream does not matter
  auto work = [&](int w) {
    err[w] = cudaSetDevice(device);
    size_t mine = 0;
    for (size_t k = (size_t)w; k < chunks && err[w] == cudaSuccess; k += kWorkers, mine++) {
      const int    r   = w * kPerWorker + (int)(mine % kPerWorker);
      const size_t off = k * kChunk, n = std::min(kChunk, bytes - off);
      if (mine >= (size_t)kPerWorker) err[w] = cudaEventSynchronize(done.e[r]);  // its previous DMA has drained
      if (err[w] != cudaSuccess) break;
      memcpy(g_ring.buf[r], s + off, n);
      err[w] = cudaMemcpyAsync(d + off, g_ring.buf[r], n, cudaMemcpyHostToDevice, st);
      if (err[w] == cudaSuccess) err[w] = cudaEventRecord(done.e[r], st);
    }
  };
  std::thread th[kWorkers];
  for (int w = 1; w < kWorkers; w++) th[w] = std::thread(work, w);
  work(0);
  for (int w = 1; w < kWorkers; w++) th[w].join();
  for (int w = 0; w < kWorkers; w++) if (err[w] != cudaSuccess) return err[w];
  // the ring is reused by the next upload (possibly on another stream or device): wait until the DMA engine has
  // drained it (which also makes it safe to destroy this call's events).  On return the copy is complete.
  return cudaStreamSynchronize(st);
}

}  // namespace lisa
