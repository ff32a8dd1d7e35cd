// lisa_b200/csrc/devmem.cu — see devmem.h.
#include "devmem.h"

#include <algorithm>
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
const size_t                                    kBudget = 24ull << 30;  // parked bytes per process

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
  if (g_cached + k.bytes <= kBudget) {
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

}  // namespace lisa
