// lisa_b200/csrc/sort_scan.cu — see sort_scan.h.
//
// scan:     two-level: 4096-element tiles (256 threads x 16) reduced to tile sums, tile sums scanned by one CTA
//           (recursively tiled when there are more than 4096 of them), then a second sweep adds the offsets.
// sort:     per 8-bit pass: k_rs_hist (per-tile digit histogram, layout [digit][tile]), exclusive scan of the
//           histogram = global base of every (digit, tile), k_rs_scatter (stable: a warp owns a contiguous chunk of
//           the tile, ranks equal digits with __match_any_sync + running per-warp counters in shared memory, the tile
//           is sorted by digit in shared memory and written out in runs).
//           Algorithmic bytes per key and pass: 8 (hist) + 12 read + 12 written.
// compact:  per-CTA ballot counts -> scan -> scatter.
#include "sort_scan.h"

namespace lisa {

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (unsigned)o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the CTA; returns the exclusive prefix, *total = CTA sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* smem /*>= 33*/, uint32_t* total) {
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const uint32_t inc = warp_incl_scan(v);
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane < nw ? smem[lane] : 0;
    const uint32_t xi = warp_incl_scan(x);
    smem[lane] = xi - x;
    if (lane == 31) smem[32] = xi;
  }
  __syncthreads();
  const uint32_t r = smem[w] + inc - v;
  if (total) *total = smem[32];
  __syncthreads();
  return r;
}

__global__ void k_scan_tile_sums(const uint32_t* in, size_t n, uint32_t* tile_sums) {
  __shared__ uint32_t sm[33];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint32_t tot;
  block_excl_scan(s, sm, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// scans one tile; tile_offsets == nullptr means a single-tile problem
__global__ void k_scan_tiles(const uint32_t* in, uint32_t* out, size_t n, const uint32_t* tile_offsets,
                             uint32_t* total) {
  __shared__ uint32_t sm[33];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;  // blocked: thread owns 16 consecutive
  uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? in[base + k] : 0u; s += v[k]; }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, sm, &tot) + (tile_offsets ? tile_offsets[blockIdx.x] : 0u);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
  if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) *total = ex;
}

static size_t n_tiles(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

size_t scan_temp_bytes(size_t n) {
  size_t bytes = 0;
  for (size_t m = n_tiles(n); ; m = n_tiles(m)) { bytes += (m + 64) * sizeof(uint32_t); if (m <= 1) break; }
  return bytes + 256;
}

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total, void* temp, cudaStream_t st) {
  if (n == 0) { if (total) cudaMemsetAsync(total, 0, sizeof(uint32_t), st); return; }
  const size_t tiles = n_tiles(n);
  if (tiles == 1) { k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr, total); return; }
  uint32_t* sums = (uint32_t*)temp;
  k_scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums);
  exclusive_scan_u32(sums, sums, tiles, nullptr, sums + tiles + 64, st);  // in place, recursive
  k_scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, n, sums, total);
}

// ---- radix sort ---------------------------------------------------------------------------------------
// A CTA owns a tile of 4096 consecutive keys (256 threads x 16), a warp a contiguous chunk of 512 of them.
//   k_rs_hist     digit histogram of every tile, layout [digit][tile] (so that ONE exclusive scan of the whole array is
//                 the global base of every (digit, tile) bucket)
//   k_rs_scatter  ranks the keys of the tile stably (per warp: __match_any_sync + running per-warp digit counters), sorts
//                 the tile by digit INTO SHARED MEMORY, and writes it out position by position: a warp's 32 stores then
//                 fall into a few long runs (one per digit: ~16 keys = 128 B on uniform digits) instead of 32 scattered
//                 sectors.  Algorithmic bytes per key and pass: 8 (hist) + 12 read + 12 written.
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)  // 4096 keys per CTA; a warp owns 512 consecutive keys
#define RS_WARPS (RS_THREADS / 32)

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned long long* __restrict__ keys, size_t n, int shift, uint32_t* hist,
                                                        uint32_t ntiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const size_t i = base + (size_t)k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

struct RsSmem {
  unsigned long long key[RS_TILE];
  unsigned int       val[RS_TILE];
  uint32_t           cnt[RS_WARPS][256];  // running count of each digit inside each warp's chunk -> start of (warp, digit) in the digit's run
  uint32_t           dstart[256];         // start of the digit's run in the sorted tile
  uint32_t           gdelta[256];         // global index of the digit's run minus dstart (mod 2^32)
  uint32_t           scan[33];
};

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals,
                                                           size_t n, int shift, const uint32_t* __restrict__ hist_scanned, uint32_t ntiles,
                                                           unsigned long long* keys_out, unsigned int* vals_out) {
  extern __shared__ unsigned char rs_raw[];
  RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_raw);
  for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&sm.cnt[0][0])[k] = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t   tile0 = (size_t)blockIdx.x * RS_TILE;
  const size_t   chunk = tile0 + (size_t)w * (RS_ITEMS * 32);
  unsigned long long key[RS_ITEMS];
  uint32_t           off[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const size_t i = chunk + (size_t)r * 32 + lane;
    const bool   ok = i < n;
    key[r] = ok ? keys[i] : ~0ull;
    const unsigned d = ok ? ((unsigned)(key[r] >> shift) & 255u) : 256u + lane;  // invalid lanes match nobody
    const unsigned same = __match_any_sync(0xffffffffu, d);
    const unsigned rank = __popc(same & ((1u << lane) - 1u));
    uint32_t       before = 0;
    if (ok) before = sm.cnt[w][d];
    __syncwarp();
    if (ok && rank == 0) sm.cnt[w][d] = before + __popc(same);
    __syncwarp();
    off[r] = before + rank;
  }
  __syncthreads();
  // thread d owns digit d: exclusive prefix of its counts over the warps, then the digit's place in the sorted tile
  {
    const unsigned d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ww++) { const uint32_t c = sm.cnt[ww][d]; sm.cnt[ww][d] = run; run += c; }
    uint32_t tot;
    const uint32_t start = block_excl_scan(run, sm.scan, &tot);
    sm.dstart[d] = start;
    sm.gdelta[d] = hist_scanned[(size_t)d * ntiles + blockIdx.x] - start;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const size_t i = chunk + (size_t)r * 32 + lane;
    if (i < n) {
      const unsigned d = (unsigned)(key[r] >> shift) & 255u;
      const uint32_t pos = sm.dstart[d] + sm.cnt[w][d] + off[r];
      sm.key[pos] = key[r];
      sm.val[pos] = vals[i];
    }
  }
  __syncthreads();
  const uint32_t n_tile = (uint32_t)(n - tile0 < (size_t)RS_TILE ? n - tile0 : (size_t)RS_TILE);
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    const uint32_t pos = (uint32_t)k * RS_THREADS + threadIdx.x;
    if (pos < n_tile) {
      const unsigned long long kk = sm.key[pos];
      const size_t dst = (size_t)(uint32_t)(sm.gdelta[(unsigned)(kk >> shift) & 255u] + pos);
      keys_out[dst] = kk;
      vals_out[dst] = sm.val[pos];
    }
  }
}

static size_t rs_tiles(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

size_t radix_sort_temp_bytes(size_t n) {
  const size_t h = 256 * rs_tiles(n > 0 ? n : 1);
  return h * sizeof(uint32_t) + 256 + scan_temp_bytes(h);
}

int radix_sort_pairs(unsigned long long* keys_in, unsigned long long* keys_out, unsigned int* vals_in, unsigned int* vals_out,
                     size_t n, int begin_bit, int end_bit, void* temp, cudaStream_t st) {
  if (n == 0) return 0;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured[dev & 63]) {  // 58 KB of shared memory per CTA: above the 48 KB a kernel gets without asking
    cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem));
    configured[dev & 63] = true;
  }
  const uint32_t tiles = (uint32_t)rs_tiles(n);
  const size_t   h = 256 * (size_t)tiles;
  uint32_t*      hist = (uint32_t*)temp;
  void*          scan_tmp = (char*)temp + ((h * sizeof(uint32_t) + 255) / 256) * 256;
  int            flip = 0;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    unsigned long long* ki = flip ? keys_out : keys_in;
    unsigned long long* ko = flip ? keys_in : keys_out;
    unsigned int*       vi = flip ? vals_out : vals_in;
    unsigned int*       vo = flip ? vals_in : vals_out;
    k_rs_hist<<<tiles, RS_THREADS, 0, st>>>(ki, n, shift, hist, tiles);
    exclusive_scan_u32(hist, hist, h, nullptr, scan_tmp, st);
    k_rs_scatter<<<tiles, RS_THREADS, sizeof(RsSmem), st>>>(ki, vi, n, shift, hist, tiles, ko, vo);
    flip ^= 1;
  }
  return flip;
}

// ---- compaction --------------------------------------------------------------------------------------
#define CP_THREADS 256
#define CP_ITEMS 8
#define CP_TILE (CP_THREADS * CP_ITEMS)

__global__ void k_cp_count(const int* __restrict__ in, size_t n, uint32_t* counts) {
  __shared__ uint32_t sm[33];
  const size_t base = (size_t)blockIdx.x * CP_TILE + (size_t)threadIdx.x * CP_ITEMS;
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; k++) c += (base + k < n && in[base + k] >= 0) ? 1u : 0u;
  uint32_t tot;
  block_excl_scan(c, sm, &tot);
  if (threadIdx.x == 0) counts[blockIdx.x] = tot;
}
__global__ void k_cp_scatter(const int* __restrict__ in, size_t n, const uint32_t* __restrict__ offsets, int* out) {
  __shared__ uint32_t sm[33];
  const size_t base = (size_t)blockIdx.x * CP_TILE + (size_t)threadIdx.x * CP_ITEMS;
  int      v[CP_ITEMS];
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < CP_ITEMS; k++) { v[k] = base + k < n ? in[base + k] : -1; c += v[k] >= 0 ? 1u : 0u; }
  uint32_t pos = block_excl_scan(c, sm, nullptr) + offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < CP_ITEMS; k++) if (v[k] >= 0) out[pos++] = v[k];
}

static size_t cp_tiles(size_t n) { return (n + CP_TILE - 1) / CP_TILE; }
size_t compact_temp_bytes(size_t n) { const size_t t = cp_tiles(n > 0 ? n : 1); return (t + 64) * sizeof(uint32_t) + 256 + scan_temp_bytes(t); }

void compact_nonneg(const int* in, int* out, size_t n, int* d_count, void* temp, cudaStream_t st) {
  if (n == 0) { cudaMemsetAsync(d_count, 0, sizeof(int), st); return; }
  const size_t t = cp_tiles(n);
  uint32_t*    counts = (uint32_t*)temp;
  void*        scan_tmp = (char*)temp + (((t + 64) * sizeof(uint32_t) + 255) / 256) * 256;
  k_cp_count<<<(unsigned)t, CP_THREADS, 0, st>>>(in, n, counts);
  exclusive_scan_u32(counts, counts, t, (uint32_t*)d_count, scan_tmp, st);
  k_cp_scatter<<<(unsigned)t, CP_THREADS, 0, st>>>(in, n, counts, out);
}

}  // namespace lisa
