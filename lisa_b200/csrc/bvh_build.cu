// lisa_b200/csrc/bvh_build.cu — GPU BVH builder (replaces optixAccelBuild, src/LiSA/src/optix_wrapper.cc:132-145).
//
// Pipeline (all on the device, one stream):
//   1. k_centroid_bounds   scene bounds of triangle centroids (block reduce + ordered-int atomics)
//   2. k_morton            63-bit Morton key per triangle; bit 63 = "triangle emits", so one sort also
//                          partitions the soup into non-emitters | emitters
//   3. radix sort          (key, triangle id) pairs, own LSD sort, 8 bits per pass over the key bits that matter
//                          (emitter bit + ceil(log2 T / 3) + 6 bits per axis; sort_scan.cu)
//   4a. PLOC (default)     parallel locally-ordered clustering over the Morton order (Meister & Bittner 2018):
//                          ONE kernel per round (k_ploc_round: nearest neighbour by merged surface area within +-radius,
//                          mutual pairs become a node, survivors compacted in order by a single-pass look-back scan),
//                          k_ploc_tail finishes the last <= 512 clusters in one CTA; ~log n rounds.
//                          Near-SAH quality: large wall triangles stay near the root, small ones cluster first.
//   4b. LBVH (fast build)  k_karras hierarchy per partition (Karras 2012), one thread per internal node,
//   5.                     then k_refit: leaf boxes + bottom-up box refit with arrival counters
//   6a. k_emit_binary      binary traversal nodes (ablation path)          -- or --
//   6b. k_collapse8        level-by-level collapse into compressed 8-wide nodes (Ylitie et al. 2017):
//                          greedy largest-area child opening, octant-affinity slot assignment,
//                          8-bit box quantisation, leaf triangles made contiguous per node
//   7. k_pack_triangles    soup -> (tri_v, tri_n) float4 arrays in final leaf order
// Algorithmic bytes per triangle: bounds 36 + keys 36 + 12, sort 32 per pass, leaf records 36 + 4 + 96,
// PLOC ~2.7 rounds-worth of (64 read + 64 written) + 44 per node, collapse ~100, pack 72 + 72 + 8 read and 100 written.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "build.h"
#include "devmem.h"
#include "sort_scan.h"
#include "common.cuh"

namespace lisa {

#define CK(x)                                                              \
  do {                                                                     \
    cudaError_t e_ = (x);                                                  \
    if (e_ != cudaSuccess) { snprintf(err, errlen, "%s: %s", #x, cudaGetErrorString(e_)); return -2; } \
  } while (0)

__device__ __forceinline__ int   f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct BoundsAcc { int lo[3], hi[3]; unsigned int n_emit; };

__global__ void k_init_bounds(BoundsAcc* b) {
  for (int a = 0; a < 3; a++) { b->lo[a] = f2ord(FLT_MAX); b->hi[a] = f2ord(-FLT_MAX); }
  b->n_emit = 0;
}

// centre of primitive t: the centre of its bounding box (the triangle's own, or the reference's when the primitives are
// references).  Box centres rather than vertex centroids: measured on B200, the same PLOC gives trees that cost 1.4 % less
// on Cornell C1 (1541 -> 1562 Msamples/s), 5 % fewer node visits and 6 % fewer triangle tests per ray on the C3 knot
// (1551 -> 1650), and the same work within 1 % on the C4 soups (profiles/r02_split_bench.jsonl).
__device__ __forceinline__ float3 prim_centre(const float* __restrict__ verts, int t, const float4* __restrict__ ref_lo,
                                              const float4* __restrict__ ref_hi) {
  if (ref_lo) {
    const float4 l = ref_lo[t], h = ref_hi[t];
    return f3(0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z));
  }
  const float* p = verts + 9ll * t;
  return f3(0.5f * (fminf(p[0], fminf(p[3], p[6])) + fmaxf(p[0], fmaxf(p[3], p[6]))), 0.5f * (fminf(p[1], fminf(p[4], p[7])) + fmaxf(p[1], fmaxf(p[4], p[7]))),
            0.5f * (fminf(p[2], fminf(p[5], p[8])) + fmaxf(p[2], fmaxf(p[5], p[8]))));
}

__global__ void k_centroid_bounds(const float* __restrict__ verts, int ntris, BoundsAcc* acc, const float4* __restrict__ ref_lo,
                                  const float4* __restrict__ ref_hi) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntris; t += gridDim.x * blockDim.x) {
    const float3 cc = prim_centre(verts, t, ref_lo, ref_hi);
    const float  c3[3] = {cc.x, cc.y, cc.z};
    for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], c3[a]); hi[a] = fmaxf(hi[a], c3[a]); }
  }
  for (int a = 0; a < 3; a++) {
    for (int o = 16; o; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&acc->lo[a], f2ord(lo[a])); atomicMax(&acc->hi[a], f2ord(hi[a])); }
  }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

__global__ void k_morton(const float* __restrict__ verts, const int* __restrict__ mat_idx,
                         const unsigned char* __restrict__ mat_emit, int num_mats, int ntris, const BoundsAcc* acc,
                         unsigned long long* keys, unsigned int* ids, unsigned int* n_emit, const int* __restrict__ ref_tri,
                         const float4* __restrict__ ref_lo, const float4* __restrict__ ref_hi) {
  int          t    = blockIdx.x * blockDim.x + threadIdx.x;
  bool         emit = false;
  if (t < ntris) {
    const float3 cc = prim_centre(verts, t, ref_lo, ref_hi);
    const float  c3[3] = {cc.x, cc.y, cc.z};
    unsigned long long q[3];
    for (int a = 0; a < 3; a++) {
      float lo = ord2f(acc->lo[a]), hi = ord2f(acc->hi[a]);
      float c  = c3[a];
      float e  = hi - lo;
      float x  = e > 0.0f ? (c - lo) / e : 0.0f;
      x        = fminf(fmaxf(x, 0.0f), 1.0f);
      q[a]     = (unsigned long long)fminf(x * 2097152.0f, 2097151.0f);
    }
    int m = mat_idx[ref_tri ? ref_tri[t] : t];
    emit  = (m >= 0 && m < num_mats) ? (mat_emit[m] != 0) : false;
    unsigned long long k = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
    keys[t] = k | (emit ? (1ull << 63) : 0ull);
    ids[t]  = (unsigned int)t;
  }
  unsigned int m = __ballot_sync(0xffffffffu, emit);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_emit, __popc(m));
}

// ---- Karras 2012 -------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
  return __clzll((long long)(a ^ b));
}

// Node ids inside one partition with n leaves: internal 0..n-2, leaf j -> (n-1)+j.  Arrays are the partition's slices.
__global__ void k_karras(const unsigned long long* __restrict__ keys, int n, int2* child, int* parent, int* cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d     = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin  = delta(keys, n, i, i - d);
  int lmax  = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j     = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s     = 0;
  int t     = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  int left  = (lo == gamma) ? (n - 1) + gamma : gamma;
  int right = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
  child[i]  = make_int2(left, right);
  cnt[i]    = hi - lo + 1;
  parent[left]  = i;
  parent[right] = i;
  if (i == 0) parent[0] = -1;
}

// box of primitive t: the triangle's, or the reference's
__device__ __forceinline__ void prim_box(const float* __restrict__ verts, unsigned int t, const float4* __restrict__ ref_lo,
                                         const float4* __restrict__ ref_hi, float3& lo, float3& hi) {
  if (ref_lo) { lo = f3(ref_lo[t]); hi = f3(ref_hi[t]); return; }
  const float* p = verts + 9ll * t;
  lo = f3(fminf(p[0], fminf(p[3], p[6])), fminf(p[1], fminf(p[4], p[7])), fminf(p[2], fminf(p[5], p[8])));
  hi = f3(fmaxf(p[0], fmaxf(p[3], p[6])), fmaxf(p[1], fmaxf(p[4], p[7])), fmaxf(p[2], fmaxf(p[5], p[8])));
}

// Leaf boxes, then bottom-up refit.  box_lo/box_hi: 2n-1 entries per partition.
__global__ void k_refit(const float* __restrict__ verts, const unsigned int* __restrict__ ids, int n,
                        const int2* __restrict__ child, const int* __restrict__ parent, float4* box_lo, float4* box_hi,
                        int* flags, const float4* __restrict__ ref_lo, const float4* __restrict__ ref_hi) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float3 lo, hi;
  prim_box(verts, ids[j], ref_lo, ref_hi, lo, hi);
  int          me = (n - 1) + j;
  box_lo[2 * (me)] = make_float4(lo.x, lo.y, lo.z, 0.0f);
  box_hi[2 * (me)] = make_float4(hi.x, hi.y, hi.z, 0.0f);
  if (n == 1) return;
  int cur = parent[me];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(&flags[cur], 1) == 0) return;  // first arrival: the sibling will continue
    int2   c  = child[cur];
    float4 l0 = __ldcg(box_lo + 2 * c.x), h0 = __ldcg(box_hi + 2 * c.x), l1 = __ldcg(box_lo + 2 * c.y), h1 = __ldcg(box_hi + 2 * c.y);
    box_lo[2 * (cur)] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.0f);
    box_hi[2 * (cur)] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.0f);
    cur = parent[cur];
  }
}

// ---- PLOC ------------------------------------------------------------------------------------------
// Clusters are kept as RECORDS in position (Morton) order — two float4 per cluster: (box lo | node id), (box hi | number
// of triangles below) — so that a round reads its neighbourhood with coalesced 16-byte loads instead of gathering boxes
// through an index array.  One kernel per round (k_ploc_round) does what took five launches: a CTA stages its tile of
// 256 clusters plus a halo of 2r on either side in shared memory, finds every cluster's nearest neighbour within +-r
// (for the tile AND a halo of r, so that mutual pairs across tile borders are seen identically by both tiles), merges
// mutual pairs (the lower position creates the node and keeps the place, the higher one is dropped) and compacts the
// survivors in order: the output offset of a tile is the sum over the tiles before it, obtained by a single-pass scan
// with decoupled look-back (tiles take tickets, publish their aggregate, and look back over their predecessors'
// aggregates / inclusive prefixes).  The same scan carries the number of merges, which numbers the new nodes
// deterministically (same input, same node ids).  Once 512 or fewer clusters remain, ONE CTA finishes all remaining
// rounds in shared memory (k_ploc_tail).
#define PLOC_TILE 256
#define PLOC_RMAX 32
#define PLOC_TAIL 512

__global__ void k_leaf_records(const float* __restrict__ verts, const unsigned int* __restrict__ ids, int n, float4* box_lo,
                               float4* box_hi, float4* rec, const float4* __restrict__ ref_lo, const float4* __restrict__ ref_hi) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float3 l3, h3;
  prim_box(verts, ids[j], ref_lo, ref_hi, l3, h3);
  const float4 lo = make_float4(l3.x, l3.y, l3.z, 0.0f), hi = make_float4(h3.x, h3.y, h3.z, 0.0f);
  box_lo[2 * ((n - 1) + j)] = lo;
  box_hi[2 * ((n - 1) + j)] = hi;
  rec[2ll * j]     = make_float4(lo.x, lo.y, lo.z, __int_as_float((n - 1) + j));
  rec[2ll * j + 1] = make_float4(hi.x, hi.y, hi.z, __int_as_float(1));
}

__device__ __forceinline__ float merged_half_area(const float4& l0, const float4& h0, const float4& l1, const float4& h1) {
  float dx = fmaxf(h0.x, h1.x) - fminf(l0.x, l1.x), dy = fmaxf(h0.y, h1.y) - fminf(l0.y, l1.y),
        dz = fmaxf(h0.z, h1.z) - fminf(l0.z, l1.z);
  return dx * dy + dy * dz + dz * dx;
}

// Nearest neighbour of the cluster at position p among positions [p-r, p+r] of [0, m) by merged surface area; slo / shi
// hold the records of positions lo0, lo0 + 1, ...  Ties are broken by a key that is SYMMETRIC in the pair — (distance in
// the order, parity of the lower position, lower position) — so candidate pairs are totally ordered, the best pair overall
// is always mutual (progress every round), and a run of coincident boxes (equal areas everywhere) pairs up (0,1), (2,3),
// ... and halves per round: a balanced tree, where "smallest position wins" made every cluster point at the start of
// the run and merged ONE pair per round into a chain.
__device__ __forceinline__ int ploc_nearest(const float4* slo, const float4* shi, int lo0, int p, int m, int r) {
  const float4 l = slo[p - lo0], h = shi[p - lo0];
  float    best = FLT_MAX;
  unsigned bkey = 0xffffffffu;
  int      bj = -1;
  const int j0 = max(0, p - r), j1 = min(m - 1, p + r);
  for (int j = j0; j <= j1; j++) {
    if (j == p) continue;
    const float    a = merged_half_area(l, h, slo[j - lo0], shi[j - lo0]);
    const unsigned key = ((unsigned)abs(j - p) << 1) | ((unsigned)min(p, j) & 1u);  // same distance + parity: the lower j comes first in the scan
    if (a < best || (a == best && key < bkey)) { best = a; bkey = key; bj = j; }
  }
  return bj;
}

// creates the node that merges the clusters (la, ha) and (lb, hb) — a at the lower position — and returns its record
// (inputs by value: callers pass the record they are about to overwrite)
__device__ __forceinline__ void ploc_make_node(int id, const float4 la, const float4 ha, const float4 lb, const float4 hb,
                                               float4* node_lo, float4* node_hi, int2* child, int* cnt, float4& rlo, float4& rhi) {
  const int  c = __float_as_int(ha.w) + __float_as_int(hb.w);
  const int2 ch = make_int2(__float_as_int(la.w), __float_as_int(lb.w));
  const float4 mlo = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.0f);
  const float4 mhi = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.0f);
  node_lo[2 * (id)] = mlo;
  node_hi[2 * (id)] = mhi;
  child[id]   = ch;
  cnt[id]     = c;
  rlo = make_float4(mlo.x, mlo.y, mlo.z, __int_as_float(id));
  rhi = make_float4(mhi.x, mhi.y, mhi.z, __int_as_float(c));
}

// tile state of the look-back scan: bits 0..29 survivors, 30..59 merges, 60..61 status (1 aggregate, 2 inclusive prefix), 62..63 round tag
#define PLOC_ST_AGG (1ull << 60)
#define PLOC_ST_PFX (2ull << 60)
#define PLOC_VAL_MASK ((1ull << 60) - 1ull)
__device__ __forceinline__ unsigned long long ploc_ld(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
__device__ __forceinline__ void ploc_st(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }

// One thread searches the neighbourhoods of PLOC_K consecutive clusters at once: the window of candidates
// [p0 - r, p0 + PLOC_K - 1 + r] is read from shared memory ONCE and every candidate is tried against the (up to) four
// clusters it is a neighbour of — a quarter of the shared-memory traffic of one cluster per thread, which is what bounded
// the search.  A CTA's 256 threads cover exactly the tile plus its halo of r on either side: tile = 1024 - 2 r positions.
#define PLOC_K 4
#define PLOC_SPAN (PLOC_TILE * PLOC_K)  // positions whose nearest neighbour a CTA computes: tile + 2 r
__host__ __device__ inline int ploc_tile_size(int r) { return PLOC_SPAN - 2 * r; }
// Shared-memory slot of the record staged for window index i: thread t reads indices 4 t + c, so the records are kept in
// four planes by i mod 4 — the lanes of a warp then read CONSECUTIVE 16-byte slots (no bank conflicts), where the plain
// layout had them 64 bytes apart (4-way conflicts: 70 % of the kernel's shared-memory wavefronts, ncu).
#define PLOC_PLANE ((PLOC_SPAN + 2 * PLOC_RMAX) / 4 + 1)
__device__ __forceinline__ int ploc_slot(int i) { return (i & 3) * PLOC_PLANE + (i >> 2); }

// RT > 0: the radius is the compile-time constant RT and the candidate loop is unrolled completely — distances, range checks and
// the tie-break keys of all 4 x (2 RT + 4) pairs become constants (up to the CTA-uniform parity of the window's start), and
// positions outside [0, m) are staged as boxes of infinite extent that no cluster ever prefers.  RT == 0: any radius.
template <int RT>
__global__ void __launch_bounds__(PLOC_TILE) k_ploc_round(const float4* __restrict__ in, int m, int r_arg, float4* __restrict__ out,
                                                          float4* node_lo, float4* node_hi, int2* child, int* cnt, int node_base,
                                                          unsigned long long* state, unsigned int* ticket, int* m_out, unsigned tag) {
  __shared__ float4             slo[4 * PLOC_PLANE], shi[4 * PLOC_PLANE];  // positions [base - 2r, base + tile + 2r), by ploc_slot()
  __shared__ int                snn[PLOC_SPAN];                                                  // positions [base - r, base + tile + r)
  __shared__ unsigned           s_tile, s_warp[PLOC_TILE / 32];
  __shared__ unsigned long long s_excl;
  const int t = threadIdx.x;
  const int r = RT > 0 ? RT : r_arg;
  if (t == 0) s_tile = atomicAdd(ticket, 1u);  // tiles are numbered in the order they START: a tile only ever waits for earlier ones
  __syncthreads();
  const int tsize = ploc_tile_size(r);
  const int tile = (int)s_tile, ntiles = (m + tsize - 1) / tsize;
  const int base = tile * tsize, lo0 = base - 2 * r;
  for (int k = t; k < PLOC_SPAN + 2 * r; k += PLOC_TILE) {
    const int p = lo0 + k, sl = ploc_slot(k);
    if (p >= 0 && p < m) { slo[sl] = in[2ll * p]; shi[sl] = in[2ll * p + 1]; }
    else {  // beyond the ends of the array: merging with this costs an infinite area
      slo[sl] = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.0f);
      shi[sl] = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
    }
  }
  __syncthreads();
  {  // nearest neighbours of positions p0 .. p0 + 3 (the tile and its halo of r: PLOC_SPAN positions, four per thread)
    const int p0 = base - r + PLOC_K * t;
    float4   l[PLOC_K], h[PLOC_K];
    float    best[PLOC_K];
    unsigned bkey[PLOC_K];
    int      bj[PLOC_K];
#pragma unroll
    for (int k = 0; k < PLOC_K; k++) {
      const int q = ploc_slot(min(max(p0 + k, 0), m - 1) - lo0);  // clamped: out-of-range positions are computed and discarded
      l[k] = slo[q]; h[k] = shi[q];
      best[k] = FLT_MAX; bkey[k] = 0xffffffffu; bj[k] = -1;
    }
    if (RT > 0) {
      const unsigned pb = (unsigned)p0 & 1u;  // parity of the window's first cluster (the same for the whole CTA: p0 = base - r + 4 t)
      const int      w0 = p0 - RT - lo0;      // shared-memory index of the first candidate
#pragma unroll
      for (int o = -RT; o <= PLOC_K - 1 + RT; o++) {  // candidate p0 + o
        const int    sl = ploc_slot(w0 + o + RT);  // w0 = 4 t: plane (o + RT) & 3, slot t + ((o + RT) >> 2)
        const float4 cl = slo[sl], ch = shi[sl];
#pragma unroll
        for (int k = 0; k < PLOC_K; k++) {
          const int d = o - k < 0 ? k - o : o - k;  // compile-time
          if (d == 0 || d > RT) continue;
          const float    a = merged_half_area(l[k], h[k], cl, ch);
          const unsigned key = ((unsigned)d << 1) | ((((unsigned)(o < k ? o : k)) & 1u) ^ pb);  // parity of min(p, j) = p0 + min(k, o)
          if (a < best[k] || (a == best[k] && key < bkey[k])) { best[k] = a; bkey[k] = key; bj[k] = p0 + o; }
        }
      }
    } else {
      const int j0 = max(0, p0 - r), j1 = min(m - 1, p0 + PLOC_K - 1 + r);
      for (int j = j0; j <= j1; j++) {
        const float4 cl = slo[ploc_slot(j - lo0)], ch = shi[ploc_slot(j - lo0)];
#pragma unroll
        for (int k = 0; k < PLOC_K; k++) {
          const int p = p0 + k, d = abs(j - p);
          const float    a = merged_half_area(l[k], h[k], cl, ch);
          const unsigned key = ((unsigned)d << 1) | ((unsigned)min(p, j) & 1u);
          // candidates of p: the positions within r of it, except p itself; best by (area, key), see ploc_nearest
          if (d != 0 && d <= r && (a < best[k] || (a == best[k] && key < bkey[k]))) { best[k] = a; bkey[k] = key; bj[k] = j; }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < PLOC_K; k++) {
      const int p = p0 + k;
      snn[PLOC_K * t + k] = (p >= 0 && p < m) ? bj[k] : -1;
    }
  }
  __syncthreads();
  // decisions for the tile's own positions: thread t owns base + 4 t .. base + 4 t + 3
  int      jn[PLOC_K];
  unsigned vmask = 0, omask = 0;
#pragma unroll
  for (int k = 0; k < PLOC_K; k++) {
    const int p = base + PLOC_K * t + k;
    jn[k] = -1;
    if (PLOC_K * t + k < tsize && p < m) {
      const int j = snn[p - (base - r)];
      const bool mutual = j >= 0 && snn[j - (base - r)] == p;
      jn[k] = j;
      if (mutual && p < j) omask |= 1u << k;
      if (!(mutual && p > j)) vmask |= 1u << k;
    }
  }
  // CTA-wide exclusive scan of (survivors, merges) per thread, packed 16 | 16
  const unsigned lane = t & 31, w = t >> 5;
  const unsigned v = (unsigned)__popc(vmask) | ((unsigned)__popc(omask) << 16);
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += x; }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  unsigned wbase = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < PLOC_TILE / 32; k++) { const unsigned x = s_warp[k]; if ((unsigned)k < w) wbase += x; tot += x; }
  const unsigned excl = wbase + inc - v;
  // decoupled look-back over the tiles before this one (warp 0)
  if (t < 32) {
    const unsigned long long TAG = (unsigned long long)(tag & 3u) << 62;
    const unsigned long long agg = (unsigned long long)(tot & 0xffffu) | ((unsigned long long)(tot >> 16) << 30);
    unsigned long long       before = 0;
    if (tile == 0) {
      if (t == 0) ploc_st(&state[0], TAG | PLOC_ST_PFX | agg);
    } else {
      if (t == 0) ploc_st(&state[tile], TAG | PLOC_ST_AGG | agg);
      int look = tile - 1;
      while (true) {
        const int          idx = look - t;
        unsigned long long sv = TAG | PLOC_ST_PFX;  // in front of tile 0: an inclusive prefix of zero
        if (idx >= 0) {
          do { sv = ploc_ld(&state[idx]); } while ((sv >> 62) != (unsigned long long)(tag & 3u) || ((sv >> 60) & 3ull) == 0ull);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, ((sv >> 60) & 3ull) == 2ull);
        const int      first = pm ? __ffs(pm) - 1 : 31;  // nearest predecessor that already knows its inclusive prefix
        unsigned long long val = t <= first ? (sv & PLOC_VAL_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        before += val;
        if (pm) break;
        look -= 32;
      }
      if (t == 0) ploc_st(&state[tile], TAG | PLOC_ST_PFX | (before + agg));
    }
    if (t == 0) s_excl = before;
  }
  __syncthreads();
  const unsigned valid_before = (unsigned)(s_excl & 0x3fffffffull), merges_before = (unsigned)((s_excl >> 30) & 0x3fffffffull);
  long long q = (long long)valid_before + (excl & 0xffffu);
  int       id = node_base + (int)(merges_before + (excl >> 16));
#pragma unroll
  for (int k = 0; k < PLOC_K; k++) {
    if (!(vmask >> k & 1u)) continue;
    const int p = base + PLOC_K * t + k;
    float4 rlo = slo[ploc_slot(p - lo0)], rhi = shi[ploc_slot(p - lo0)];
    if (omask >> k & 1u)
      ploc_make_node(id++, rlo, rhi, slo[ploc_slot(jn[k] - lo0)], shi[ploc_slot(jn[k] - lo0)], node_lo, node_hi, child, cnt, rlo, rhi);
    out[2 * q]     = rlo;
    out[2 * q + 1] = rhi;
    q++;
  }
  if (tile == ntiles - 1 && t == 0) {  // the holder of the last ticket started last: every ticket of this round is taken
    *m_out  = (int)(valid_before + (tot & 0xffffu));
    *ticket = 0u;
  }
}

// all remaining rounds for m <= PLOC_TAIL clusters, in one CTA; *root_out = node id of the last cluster standing
__global__ void __launch_bounds__(PLOC_TAIL) k_ploc_tail(const float4* __restrict__ in, int m, int r, float4* node_lo, float4* node_hi,
                                                         int2* child, int* cnt, int node_base, int* root_out) {
  __shared__ float4   slo[2][PLOC_TAIL], shi[2][PLOC_TAIL];
  __shared__ int      snn[PLOC_TAIL];
  __shared__ unsigned s_warp[PLOC_TAIL / 32];
  const int      t = threadIdx.x;
  const unsigned lane = t & 31, w = t >> 5;
  int cur = 0;
  if (t < m) { slo[0][t] = in[2 * t]; shi[0][t] = in[2 * t + 1]; }
  __syncthreads();
  while (m > 1) {
    if (t < m) snn[t] = ploc_nearest(slo[cur], shi[cur], 0, t, m, r);
    __syncthreads();
    int  j = -1;
    bool valid = false, owner = false;
    if (t < m) {
      j = snn[t];
      const bool mutual = j >= 0 && snn[j] == t;
      owner = mutual && t < j;
      valid = !(mutual && t > j);
    }
    const unsigned v = (valid ? 1u : 0u) | (owner ? 0x10000u : 0u);
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += x; }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    unsigned wbase = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < PLOC_TAIL / 32; k++) { const unsigned x = s_warp[k]; if ((unsigned)k < w) wbase += x; tot += x; }
    const unsigned excl = wbase + inc - v;
    if (valid) {
      const unsigned q = excl & 0xffffu;
      float4 rlo = slo[cur][t], rhi = shi[cur][t];
      if (owner) ploc_make_node(node_base + (int)(excl >> 16), rlo, rhi, slo[cur][j], shi[cur][j], node_lo, node_hi, child, cnt, rlo, rhi);
      slo[cur ^ 1][q] = rlo;
      shi[cur ^ 1][q] = rhi;
    }
    __syncthreads();
    if ((tot >> 16) == 0u) break;  // no mutual pair (NaN boxes): the host reports it
    m = (int)(tot & 0xffffu);
    node_base += (int)(tot >> 16);
    cur ^= 1;
  }
  if (t == 0) { root_out[0] = __float_as_int(slo[cur][0].w); root_out[1] = m; }
}

// ---- SAH restructuring of the binary hierarchy: tree rotations ------------------------------------------------------
// (Kensler 2008, "Tree Rotations for Improving Bounding Volume Hierarchies".)  At an internal node X = (L, R) a rotation swaps
// one child with a grandchild on the other side — L <-> RL, L <-> RR, R <-> LL or R <-> LR — which changes exactly one
// bounding box, that of the child whose pair is re-formed.  With the surface-area heuristic (a node's cost is proportional
// to its area) the best of the four is the one that shrinks that box's area most; it is applied when it shrinks it at all.
// The pass is bottom-up, one thread per leaf climbing through arrival counters like a refit: the thread that arrives
// second at X owns X's whole subtree (every node below has been processed, nobody else is inside), so rotations need no
// locks.  Triangle counts and boxes of the re-formed node are updated in place; parents are recomputed per pass.
__global__ void k_parents(const int2* __restrict__ child, int n, int root, int* parent) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int2 c = child[i];
  parent[c.x] = i;
  parent[c.y] = i;
  if (i == root) parent[i] = -1;
}

__device__ __forceinline__ float union_half_area(const float4& l0, const float4& h0, const float4& l1, const float4& h1) {
  return merged_half_area(l0, h0, l1, h1);
}

__global__ void k_rotate(int n, int2* child, const int* __restrict__ parent, float4* box_lo, float4* box_hi, int* cnt, int* flags,
                         unsigned int* n_rotations) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int cur = parent[(n - 1) + j];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(&flags[cur], 1) == 0) return;  // first arrival: the sibling's thread will continue
    __threadfence();
    const int2 c = __ldcg(&child[cur]);
    const int  L = c.x, R = c.y;
    const float4 lL = __ldcg(box_lo + 2 * L), hL = __ldcg(box_hi + 2 * L), lR = __ldcg(box_lo + 2 * R), hR = __ldcg(box_hi + 2 * R);
    float best = 0.0f;  // area saved
    int   which = -1;
    int2  cR = make_int2(-1, -1), cL = make_int2(-1, -1);
    float4 lRL, hRL, lRR, hRR, lLL, hLL, lLR, hLR;
    if (R < n - 1) {  // R is internal: L <-> RL (R' = L + RR) or L <-> RR (R' = RL + L)
      cR = __ldcg(&child[R]);
      lRL = __ldcg(box_lo + 2 * cR.x); hRL = __ldcg(box_hi + 2 * cR.x); lRR = __ldcg(box_lo + 2 * cR.y); hRR = __ldcg(box_hi + 2 * cR.y);
      const float aR = union_half_area(lR, hR, lR, hR);
      const float s0 = aR - union_half_area(lL, hL, lRR, hRR), s1 = aR - union_half_area(lRL, hRL, lL, hL);
      if (s0 > best) { best = s0; which = 0; }
      if (s1 > best) { best = s1; which = 1; }
    }
    if (L < n - 1) {  // L is internal: R <-> LL (L' = R + LR) or R <-> LR (L' = LL + R)
      cL = __ldcg(&child[L]);
      lLL = __ldcg(box_lo + 2 * cL.x); hLL = __ldcg(box_hi + 2 * cL.x); lLR = __ldcg(box_lo + 2 * cL.y); hLR = __ldcg(box_hi + 2 * cL.y);
      const float aL = union_half_area(lL, hL, lL, hL);
      const float s2 = aL - union_half_area(lR, hR, lLR, hLR), s3 = aL - union_half_area(lLL, hLL, lR, hR);
      if (s2 > best) { best = s2; which = 2; }
      if (s3 > best) { best = s3; which = 3; }
    }
    auto count = [&](int node) { return node >= n - 1 ? 1 : __ldcg(&cnt[node]); };
    auto remake = [&](int node, int a, int b, const float4& la, const float4& ha, const float4& lb, const float4& hb) {
      child[node]  = make_int2(a, b);
      box_lo[2 * (node)] = make_float4(fminf(la.x, lb.x), fminf(la.y, lb.y), fminf(la.z, lb.z), 0.0f);
      box_hi[2 * (node)] = make_float4(fmaxf(ha.x, hb.x), fmaxf(ha.y, hb.y), fmaxf(ha.z, hb.z), 0.0f);
      cnt[node]    = count(a) + count(b);
    };
    if (which == 0) { remake(R, L, cR.y, lL, hL, lRR, hRR); child[cur] = make_int2(cR.x, R); }
    else if (which == 1) { remake(R, cR.x, L, lRL, hRL, lL, hL); child[cur] = make_int2(cR.y, R); }
    else if (which == 2) { remake(L, R, cL.y, lR, hR, lLR, hLR); child[cur] = make_int2(L, cL.x); }
    else if (which == 3) { remake(L, cL.x, R, lLL, hLL, lR, hR); child[cur] = make_int2(L, cL.y); }
    if (which >= 0) atomicAdd(n_rotations, 1u);
    cur = parent[cur];
  }
}

// ---- binary traversal nodes -----------------------------------------------------------------------
// One node per internal Karras node; tri_base = first final triangle index of the partition.
__global__ void k_emit_binary(int n, const int2* __restrict__ child, const float4* __restrict__ box_lo,
                              const float4* __restrict__ box_hi, int node_base, int tri_base, float4* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max(n - 1, 1)) return;
  float4* o = out + 4ll * (node_base + i);
  if (n == 1) {  // single triangle: child 0 = the leaf, child 1 = empty box
    float4 l = box_lo[2 * (0)], h = box_hi[2 * (0)];
    o[0] = make_float4(l.x, h.x, l.y, h.y);
    o[1] = make_float4(FLT_MAX, -FLT_MAX, FLT_MAX, -FLT_MAX);
    o[2] = make_float4(l.z, h.z, FLT_MAX, -FLT_MAX);
    o[3] = make_float4(__int_as_float(~tri_base), __int_as_float(~tri_base), 0.0f, 0.0f);
    return;
  }
  int2   c  = child[i];
  float4 l0 = box_lo[2 * (c.x)], h0 = box_hi[2 * (c.x)], l1 = box_lo[2 * (c.y)], h1 = box_hi[2 * (c.y)];
  o[0] = make_float4(l0.x, h0.x, l0.y, h0.y);
  o[1] = make_float4(l1.x, h1.x, l1.y, h1.y);
  o[2] = make_float4(l0.z, h0.z, l1.z, h1.z);
  int e0 = c.x >= n - 1 ? ~(tri_base + (c.x - (n - 1))) : node_base + c.x;
  int e1 = c.y >= n - 1 ? ~(tri_base + (c.y - (n - 1))) : node_base + c.y;
  o[3] = make_float4(__int_as_float(e0), __int_as_float(e1), 0.0f, 0.0f);
}

// ---- collapse to compressed 8-wide nodes --------------------------------------------------------
#ifndef LEAF_MAX
#define LEAF_MAX 3
#endif
struct WorkItem { int node2; int out; };  // binary node to expand -> index of the wide node to write

__device__ __forceinline__ float half_area(const float4& lo, const float4& hi) {
  float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ int node_count(int node, int n, const int* __restrict__ cnt) {
  return node >= n - 1 ? 1 : cnt[node];
}
// leaves (sorted positions) of a subtree with at most LEAF_MAX triangles
__device__ __forceinline__ int gather_leaves(int node, int n, const int2* __restrict__ child, int* out) {
  int stack[2 * LEAF_MAX], sp = 0, k = 0;
  stack[sp++] = node;
  while (sp) {
    int c = stack[--sp];
    if (c >= n - 1) out[k++] = c - (n - 1);
    else { int2 cc = child[c]; stack[sp++] = cc.y; stack[sp++] = cc.x; }
  }
  return k;
}

// counters: [0] next free wide node, [1] next free final triangle slot, [2] items in the output queue
__global__ void k_collapse8(const WorkItem* __restrict__ in, int n_in, WorkItem* out_q, int n, const int2* __restrict__ child,
                            const int* __restrict__ range, const float4* __restrict__ box_lo,
                            const float4* __restrict__ box_hi, int* counters, float4* nodes,
                            unsigned int* final_to_sorted, int sorted_base, float leaf_sah, int root2,
                            unsigned long long* sah_acc) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  // surface-area estimate of the node visits per ray (BuildOutput::sah_nodes_per_ray): every wide node adds area / root area,
  // in 2^-20 fixed point so that the sum does not depend on the order of the atomics; one atomic per warp
  unsigned long long sah_q = 0ull;
  if (w < n_in) {
    const int   node2 = n == 1 ? 0 : in[w].node2, r2 = n == 1 ? 0 : root2;
    const float ar = half_area(box_lo[2 * (r2)], box_hi[2 * (r2)]), an = half_area(box_lo[2 * (node2)], box_hi[2 * (node2)]);
    sah_q = ar > 0.0f ? (unsigned long long)(fminf(an / ar, 1.0f) * 1048576.0f) : 1048576ull;
  }
  for (int o = 16; o; o >>= 1) sah_q += __shfl_xor_sync(0xffffffffu, sah_q, o);
  if ((threadIdx.x & 31) == 0 && sah_q) atomicAdd(sah_acc, sah_q);
  if (w >= n_in) return;
  const WorkItem item = in[w];
  // the children gathered so far, with everything the steps below need of them read ONCE (the kernel waits on these
  // gathers: 10.6 of its stall cycles per instruction were global loads when every step fetched the boxes again)
  int    ch[8], ccnt[8];
  float4 clo[8], chi[8];
  bool   copen[8];  // a subtree of 2..LEAF_MAX triangles that is better NOT made one leaf (below)
  int    nc = 0;
  float4 plo, phi;
  // Leaves by the surface-area heuristic (leaf_sah >= 0): a subtree of 2..LEAF_MAX triangles costs, as ONE leaf, a test of
  // all its triangles whenever its box is hit — count x area(box) — and, opened into its two halves, count_a x area(a) +
  // count_b x area(b) plus one more child slot in the wide node (leaf_sah x area(box), in units of a triangle test).
  // Neighbours in a mesh share their box (a quad's two halves: 2A against 2A + slot) and stay one leaf; unrelated
  // triangles of a soup (merged box ~2.5x each half) are kept apart: 3x fewer triangle tests per ray on the C4 soups.
  // everything about a child is fetched in ONE round of independent loads: its count, its box and (internal nodes) its two
  // children, which the opening step below and the rule above would otherwise fetch in a second, dependent round
  int2 cch[8];
  auto add_child = [&](int k, int node) {
    const bool internal = node < n - 1;
    const int  cnt = internal ? range[node] : 1;
    const int2 c2  = internal ? child[node] : make_int2(node, node);
    ch[k] = node; ccnt[k] = cnt; cch[k] = c2; clo[k] = box_lo[2 * (node)]; chi[k] = box_hi[2 * (node)];
    copen[k] = false;
    if (leaf_sah >= 0.0f && cnt >= 2 && cnt <= LEAF_MAX) {
      // a subtree of <= 3 triangles: a child that is not a leaf of the binary tree holds all the others
      const float am = half_area(clo[k], chi[k]);
      const float ca = (float)(c2.x >= n - 1 ? 1 : cnt - 1) * half_area(box_lo[2 * (c2.x)], box_hi[2 * (c2.x)]);
      const float cb = (float)(c2.y >= n - 1 ? 1 : cnt - 1) * half_area(box_lo[2 * (c2.y)], box_hi[2 * (c2.y)]);
      copen[k] = ca + cb + leaf_sah * am < (float)cnt * am;
    }
  };
  if (n == 1) {  // degenerate partition: one triangle, no binary internal node
    add_child(nc++, 0);
    plo = clo[0]; phi = chi[0];
  } else {
    int2 c = child[item.node2];
    add_child(nc++, c.x); add_child(nc++, c.y);
    plo = box_lo[2 * (item.node2)]; phi = box_hi[2 * (item.node2)];
    while (nc < 8) {
      int   best = -1;
      float ba   = -1.0f;
      for (int k = 0; k < nc; k++) {
        if (ccnt[k] <= LEAF_MAX && !copen[k]) continue;  // stays a leaf
        float a = half_area(clo[k], chi[k]);
        if (a > ba) { ba = a; best = k; }
      }
      if (best < 0) break;
      const int2 c2 = cch[best];
      add_child(best, c2.x);
      add_child(nc++, c2.y);
    }
  }
  // octant-affinity slot assignment (greedy): slot bit 4/2/1 set = child lies on the +x/+y/+z side
  const float3 pc = f3((plo.x + phi.x) * 0.5f, (plo.y + phi.y) * 0.5f, (plo.z + phi.z) * 0.5f);
  float3 rel[8];
  for (int k = 0; k < nc; k++) {
    const float4 l = clo[k], h = chi[k];
    rel[k] = f3((l.x + h.x) * 0.5f - pc.x, (l.y + h.y) * 0.5f - pc.y, (l.z + h.z) * 0.5f - pc.z);
  }
  int slot_k[8];  // slot -> index into ch / ccnt / clo / chi, -1 = empty
  for (int s = 0; s < 8; s++) slot_k[s] = -1;
  unsigned int child_done = 0, slot_done = 0;
  for (int it = 0; it < nc; it++) {
    float bc = -FLT_MAX;
    int   bk = -1, bs = -1;
    for (int k = 0; k < nc; k++) {
      if (child_done >> k & 1) continue;
      for (int s = 0; s < 8; s++) {
        if (slot_done >> s & 1) continue;
        float c = ((s & 4) ? rel[k].x : -rel[k].x) + ((s & 2) ? rel[k].y : -rel[k].y) + ((s & 1) ? rel[k].z : -rel[k].z);
        if (c > bc) { bc = c; bk = k; bs = s; }
      }
    }
    slot_k[bs] = bk;
    child_done |= 1u << bk;
    slot_done |= 1u << bs;
  }
  // quantisation frame
  int   ex[3];
  float scale[3];
  const float plo3[3] = {plo.x, plo.y, plo.z}, phi3[3] = {phi.x, phi.y, phi.z};
  for (int a = 0; a < 3; a++) {
    float ext = (phi3[a] - plo3[a]) * (1.0f / 255.0f);
    int   e   = -126;
    if (ext > 0.0f) {
      frexpf(ext, &e);  // ext = m * 2^e with m in [0.5, 1) => 2^e >= ext
      e = max(e, -126);
    }
    e = min(e, 126);
    ex[a]    = e;
    scale[a] = __int_as_float((127 - e) << 23);  // 2^-e, exact
  }
  // counts
  int n_inner = 0, n_tri = 0;
  for (int s = 0; s < 8; s++) {
    if (slot_k[s] < 0) continue;
    const int cnt = ccnt[slot_k[s]];
    if (cnt <= LEAF_MAX) n_tri += cnt; else n_inner++;
  }
  int child_base = n_inner ? atomicAdd(&counters[0], n_inner) : 0;
  int tri_base   = n_tri ? atomicAdd(&counters[1], n_tri) : 0;
  int q_base     = n_inner ? atomicAdd(&counters[2], n_inner) : 0;
  unsigned int imask = 0;
  unsigned int meta[8], qlo[3][8], qhi[3][8];
  int k_inner = 0, tri_off = 0;
  for (int s = 0; s < 8; s++) {
    meta[s] = 0;
    for (int a = 0; a < 3; a++) { qlo[a][s] = 255; qhi[a][s] = 0; }  // empty: inverted box, never hit
    if (slot_k[s] < 0) continue;
    const int    c = ch[slot_k[s]];
    const float4 l = clo[slot_k[s]], h = chi[slot_k[s]];
    const float l3[3] = {l.x, l.y, l.z}, h3[3] = {h.x, h.y, h.z};
    for (int a = 0; a < 3; a++) {
      float step = __int_as_float((ex[a] + 127) << 23);  // 2^e, exact
      int   ql = (int)floorf((l3[a] - plo3[a]) * scale[a]);
      int   qh = (int)ceilf((h3[a] - plo3[a]) * scale[a]);
      ql = max(0, min(255, ql)); qh = max(0, min(255, qh));
      // conservative against the exact value of p + q*2^e (double holds it exactly)
      while (ql > 0 && (double)plo3[a] + (double)ql * (double)step > (double)l3[a]) ql--;
      while (qh < 255 && (double)plo3[a] + (double)qh * (double)step < (double)h3[a]) qh++;
      qlo[a][s] = (unsigned)ql; qhi[a][s] = (unsigned)qh;
    }
    const int cnt = ccnt[slot_k[s]];
    if (cnt <= LEAF_MAX) {
      meta[s]   = (((1u << cnt) - 1u) << 5) | (unsigned)tri_off;
      int leaves[LEAF_MAX];
      gather_leaves(c, n, child, leaves);
      for (int k = 0; k < cnt; k++) final_to_sorted[tri_base + tri_off + k] = (unsigned)(sorted_base + leaves[k]);
      tri_off += cnt;
    } else {
      meta[s] = (1u << 5) | (24u + (unsigned)s);
      imask |= 1u << s;
      out_q[q_base + k_inner] = WorkItem{c, child_base + k_inner};
      k_inner++;
    }
  }
  auto pack4 = [](const unsigned int* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
  float4* o = nodes + 5ll * item.out;
  unsigned int eim = (unsigned)(ex[0] + 127) | ((unsigned)(ex[1] + 127) << 8) | ((unsigned)(ex[2] + 127) << 16) | (imask << 24);
  o[0] = make_float4(plo.x, plo.y, plo.z, __uint_as_float(eim));
  o[1] = make_float4(__uint_as_float((unsigned)child_base), __uint_as_float((unsigned)tri_base),
                     __uint_as_float(pack4(meta)), __uint_as_float(pack4(meta + 4)));
  o[2] = make_float4(__uint_as_float(pack4(qlo[0])), __uint_as_float(pack4(qlo[0] + 4)), __uint_as_float(pack4(qlo[1])),
                     __uint_as_float(pack4(qlo[1] + 4)));
  o[3] = make_float4(__uint_as_float(pack4(qlo[2])), __uint_as_float(pack4(qlo[2] + 4)), __uint_as_float(pack4(qhi[0])),
                     __uint_as_float(pack4(qhi[0] + 4)));
  o[4] = make_float4(__uint_as_float(pack4(qhi[1])), __uint_as_float(pack4(qhi[1] + 4)), __uint_as_float(pack4(qhi[2])),
                     __uint_as_float(pack4(qhi[2] + 4)));
}

__global__ void k_iota_sorted(unsigned int* final_to_sorted, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) final_to_sorted[i] = (unsigned)i;
}

// final order -> packed float4 arrays
__global__ void k_pack_triangles(const float* __restrict__ verts, const float* __restrict__ normals,
                                 const int* __restrict__ mat_idx, const unsigned int* __restrict__ sorted_ids,
                                 const unsigned int* __restrict__ final_to_sorted, int ntris, float4* tri_v,
                                 float4* tri_n, int* final_to_orig, const int* __restrict__ ref_tri) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= ntris) return;
  unsigned int t = sorted_ids[final_to_sorted[f]];
  if (ref_tri) t = (unsigned int)ref_tri[t];  // a reference: the packed arrays get a full copy of its triangle
  const float* p = verts + 9ll * t;
  const float* q = normals + 9ll * t;
  tri_v[3ll * f + 0] = make_float4(p[0], p[1], p[2], __int_as_float(mat_idx[t]));
  tri_v[3ll * f + 1] = make_float4(p[3], p[4], p[5], 0.0f);
  tri_v[3ll * f + 2] = make_float4(p[6], p[7], p[8], 0.0f);
  tri_n[3ll * f + 0] = make_float4(q[0], q[1], q[2], __int_as_float((int)t));
  tri_n[3ll * f + 1] = make_float4(q[3], q[4], q[5], 0.0f);
  tri_n[3ll * f + 2] = make_float4(q[6], q[7], q[8], 0.0f);
  final_to_orig[f] = (int)t;
}

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

int build_bvh(const BuildInput& in, BuildOutput* out, cudaStream_t st, char* err, size_t errlen) {
  const int T = in.num_tris;
  const bool dbg = getenv("LISA_DEBUG_TIMING") != nullptr;
  // stage boundaries: a CUDA event each (BuildOutput::stage_ms, always), an NVTX range each (visible to nsys / ncu --nvtx when
  // a tool is attached, free otherwise), and under LISA_DEBUG_TIMING a synchronised wall-clock line on stderr
  struct StageEvents {
    cudaEvent_t e[8] = {};
    int         n = 0;
    const char* name[8] = {};
    ~StageEvents() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
  } sev;
  bool range_open = false;
  auto tick = [&](const char* what) {  // closes the stage `what` (nullptr: the start of the build)
    static std::chrono::steady_clock::time_point last;
    if (range_open) { nvtxRangePop(); range_open = false; }
    if (sev.n < 8 && cudaEventCreate(&sev.e[sev.n]) == cudaSuccess) { cudaEventRecord(sev.e[sev.n], st); sev.name[sev.n] = what; sev.n++; }
    if (!dbg) return;
    cudaStreamSynchronize(st);
    auto now = std::chrono::steady_clock::now();
    if (what) fprintf(stderr, "  build %-10s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  };
  auto stage = [&](const char* name) { nvtxRangePushA(name); range_open = true; };
  memset(out, 0, sizeof(*out));
  tick(nullptr);
  out->root_other = out->root_emit = -1;
  out->num_tris = T;
  for (int k = 0; k < 3; k++) { out->box_other[k] = out->box_emit[k] = FLT_MAX; out->box_other[3 + k] = out->box_emit[3 + k] = -FLT_MAX; }
  if (T == 0) return 0;
  stage("lisa: bvh morton + sort");

  BoundsAcc*          d_acc;
  unsigned long long *d_keys, *d_keys2;
  unsigned int *      d_ids, *d_ids2, *d_final_to_sorted;
  CK(dev_alloc((void**)&d_acc, sizeof(BoundsAcc)));
  CK(dev_alloc((void**)&d_keys, sizeof(unsigned long long) * T));
  CK(dev_alloc((void**)&d_keys2, sizeof(unsigned long long) * T));
  CK(dev_alloc((void**)&d_ids, sizeof(unsigned int) * T));
  CK(dev_alloc((void**)&d_ids2, sizeof(unsigned int) * T));
  CK(dev_alloc((void**)&d_final_to_sorted, sizeof(unsigned int) * T));

  k_init_bounds<<<1, 1, 0, st>>>(d_acc);
  k_centroid_bounds<<<min(cdiv(T, 256), 148 * 8), 256, 0, st>>>(in.d_verts, T, d_acc, in.d_ref_lo, in.d_ref_hi);
  k_morton<<<cdiv(T, 256), 256, 0, st>>>(in.d_verts, in.d_mat_idx, in.d_mat_emit, in.num_mats, T, d_acc, d_keys, d_ids,
                                         &d_acc->n_emit, in.d_ref_tri, in.d_ref_lo, in.d_ref_hi);
  const size_t tmp_bytes = radix_sort_temp_bytes((size_t)T);
  void* d_tmp;
  CK(dev_alloc((void**)&d_tmp, tmp_bytes));
  // PLOC only needs the Morton ORDER to be a good one: its window search looks +-radius positions around every cluster.
  // The sort therefore covers the emitter bit and the top ceil(log2(T) / 3) + 6 bits per axis — 64 cells per axis finer
  // than the mean triangle spacing, so a region 2^18 times denser than average still resolves — and leaves triangles that
  // agree in all of those in input order (stable sort): 6 passes instead of 8 at 100M triangles, 4 at 1k.  The Karras
  // hierarchy needs fully sorted keys; LISA_MORTON_BITS=21 sorts everything.
  int bits_axis = 21;
  if (!in.lbvh) {
    int lg = 0;
    while ((1ll << lg) < (long long)T) lg++;
    bits_axis = std::min(21, (lg + 2) / 3 + 6);
    if (const char* e = getenv("LISA_MORTON_BITS")) bits_axis = std::max(1, std::min(21, atoi(e)));
  }
  const int sort_passes = (3 * bits_axis + 1 + 7) / 8, begin_bit = 64 - 8 * sort_passes;
  if (radix_sort_pairs(d_keys, d_keys2, d_ids, d_ids2, (size_t)T, begin_bit, 64, d_tmp, st) == 0) {  // even number of passes: ends in *_in
    std::swap(d_keys, d_keys2);
    std::swap(d_ids, d_ids2);
  }
  tick("morton+sort");
  stage("lisa: bvh hierarchy");
  BoundsAcc h_acc;
  CK(cudaMemcpyAsync(&h_acc, d_acc, sizeof(h_acc), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int n_emit = (int)h_acc.n_emit, n_other = T - n_emit;
  out->num_emit_tris = n_emit;

  // hierarchy storage for both partitions: partition p uses node slice [base_p, base_p + 2 n_p - 1)
  int2*   d_child;
  int *   d_range, *d_parent = nullptr, *d_flags = nullptr;  // d_range: triangles below each internal node
  float4 *d_box, *d_lo, *d_hi;  // boxes of the binary nodes, interleaved (lo, hi): one 32-byte sector per box; d_lo[2 i], d_hi[2 i]
  const size_t NN = 2ull * T + 2;
  CK(dev_alloc((void**)&d_child, sizeof(int2) * NN));
  CK(dev_alloc((void**)&d_range, sizeof(int) * NN));
  CK(dev_alloc((void**)&d_box, sizeof(float4) * 2 * NN));
  d_lo = d_box; d_hi = d_box + 1;
  if (in.lbvh) {
    CK(dev_alloc((void**)&d_parent, sizeof(int) * NN));
    CK(dev_alloc((void**)&d_flags, sizeof(int) * NN));
    CK(cudaMemsetAsync(d_flags, 0, sizeof(int) * NN, st));
  }

  struct Part { int n, sorted_base, slice, root; } parts[2] = {{n_other, 0, 0, 0}, {n_emit, n_other, 2 * n_other + 1, 0}};
  float4*             d_rec[2] = {nullptr, nullptr};  // PLOC cluster records, ping-pong
  unsigned long long* d_tile_state = nullptr;
  int*                d_ploc_ctl = nullptr;           // [0] ticket, [1] clusters after the round, [2] root id, [3] clusters left by the tail
  const int radius = std::max(1, std::min(in.ploc_radius, PLOC_RMAX));
  const size_t max_tiles = ((size_t)T + ploc_tile_size(radius) - 1) / ploc_tile_size(radius) + 1;
  if (!in.lbvh) {
    CK(dev_alloc((void**)&d_rec[0], sizeof(float4) * 2 * (size_t)T));
    CK(dev_alloc((void**)&d_rec[1], sizeof(float4) * 2 * (size_t)T));
    CK(dev_alloc((void**)&d_tile_state, sizeof(unsigned long long) * max_tiles));
    CK(dev_alloc((void**)&d_ploc_ctl, sizeof(int) * 4));
    CK(cudaMemsetAsync(d_ploc_ctl, 0, sizeof(int) * 4, st));
  }
  for (int p = 0; p < 2; p++) {
    Part& P = parts[p];
    if (P.n == 0) continue;
    if (in.lbvh) {
      if (P.n > 1)
        k_karras<<<cdiv(P.n - 1, 256), 256, 0, st>>>(d_keys2 + P.sorted_base, P.n, d_child + P.slice, d_parent + P.slice,
                                                     d_range + P.slice);
      k_refit<<<cdiv(P.n, 256), 256, 0, st>>>(in.d_verts, d_ids2 + P.sorted_base, P.n, d_child + P.slice,
                                              d_parent + P.slice, d_lo + 2 * P.slice, d_hi + 2 * P.slice, d_flags + P.slice, in.d_ref_lo, in.d_ref_hi);
      P.root = 0;
    } else {
      k_leaf_records<<<cdiv(P.n, 256), 256, 0, st>>>(in.d_verts, d_ids2 + P.sorted_base, P.n, d_lo + 2 * P.slice, d_hi + 2 * P.slice, d_rec[0],
                                                     in.d_ref_lo, in.d_ref_hi);
      int m = P.n, cur = 0, rounds = 0;
      // stale tile states of the other partition must not be taken for this one's
      CK(cudaMemsetAsync(d_tile_state, 0, sizeof(unsigned long long) * (((size_t)P.n + ploc_tile_size(radius) - 1) / ploc_tile_size(radius)), st));
      while (m > PLOC_TAIL) {
        // the default radius (4) and the former one (16) have their own, fully unrolled instances
        auto round_kernel = radius == 4 ? k_ploc_round<4> : radius == 16 ? k_ploc_round<16> : k_ploc_round<0>;
        round_kernel<<<cdiv(m, ploc_tile_size(radius)), PLOC_TILE, 0, st>>>(d_rec[cur], m, radius, d_rec[cur ^ 1], d_lo + 2 * P.slice, d_hi + 2 * P.slice,
                                                                           d_child + P.slice, d_range + P.slice, P.n - m, d_tile_state,
                                                                           (unsigned int*)d_ploc_ctl, d_ploc_ctl + 1, (unsigned)rounds);
        int m2 = 0;
        CK(cudaMemcpyAsync(&m2, d_ploc_ctl + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (m2 >= m || m2 < 1) { snprintf(err, errlen, "PLOC made no progress (%d -> %d clusters)", m, m2); return -5; }
        m = m2;
        cur ^= 1;
        rounds++;
      }
      k_ploc_tail<<<1, PLOC_TAIL, 0, st>>>(d_rec[cur], m, radius, d_lo + 2 * P.slice, d_hi + 2 * P.slice, d_child + P.slice, d_range + P.slice,
                                          P.n - m, d_ploc_ctl + 2);
      int tail[2] = {0, 0};
      CK(cudaMemcpyAsync(tail, d_ploc_ctl + 2, sizeof(tail), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (tail[1] != 1) { snprintf(err, errlen, "PLOC made no progress (%d clusters left)", tail[1]); return -5; }
      if (dbg) fprintf(stderr, "  ploc partition %d: %d leaves, %d rounds + tail from %d clusters\n", p, P.n, rounds, m);
      P.root = P.n > 1 ? tail[0] : 0;
    }
  }

  tick(in.lbvh ? "lbvh" : "ploc");
  if (in.rotate_passes > 0) {
    stage("lisa: bvh rotations");
    int *d_par = nullptr, *d_arr = nullptr;
    unsigned int* d_nrot = nullptr;
    CK(dev_alloc((void**)&d_par, sizeof(int) * NN));
    CK(dev_alloc((void**)&d_arr, sizeof(int) * NN));
    CK(dev_alloc((void**)&d_nrot, sizeof(unsigned int)));
    CK(cudaMemsetAsync(d_nrot, 0, sizeof(unsigned int), st));
    for (int pass = 0; pass < in.rotate_passes; pass++)
      for (int p = 0; p < 2; p++) {
        const Part& P = parts[p];
        if (P.n < 3) continue;
        k_parents<<<cdiv(P.n - 1, 256), 256, 0, st>>>(d_child + P.slice, P.n, P.root, d_par + P.slice);
        CK(cudaMemsetAsync(d_arr + P.slice, 0, sizeof(int) * (size_t)(P.n - 1), st));
        k_rotate<<<cdiv(P.n, 256), 256, 0, st>>>(P.n, d_child + P.slice, d_par + P.slice, d_lo + 2 * P.slice, d_hi + 2 * P.slice,
                                                 d_range + P.slice, d_arr + P.slice, d_nrot);
      }
    if (dbg) {
      unsigned int nrot = 0;
      CK(cudaMemcpyAsync(&nrot, d_nrot, sizeof(nrot), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      fprintf(stderr, "  tree rotations: %u in %d pass(es)\n", nrot, in.rotate_passes);
    }
    dev_free(d_par); dev_free(d_arr); dev_free(d_nrot);
    tick("rotations");
  }
  stage("lisa: bvh collapse");
  for (int p = 0; p < 2; p++) {  // root bounds of each partition (node 0 of its slice)
    float* dst = p == 0 ? out->box_other : out->box_emit;
    for (int k = 0; k < 3; k++) { dst[k] = FLT_MAX; dst[3 + k] = -FLT_MAX; }
    if (parts[p].n == 0) continue;
    float4 lo, hi;
    CK(cudaMemcpyAsync(&lo, d_lo + 2 * (parts[p].slice + parts[p].root), sizeof(float4), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&hi, d_hi + 2 * (parts[p].slice + parts[p].root), sizeof(float4), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    dst[0] = lo.x; dst[1] = lo.y; dst[2] = lo.z; dst[3] = hi.x; dst[4] = hi.y; dst[5] = hi.z;
  }

  float4* d_nodes = nullptr;
  int     total_nodes = 0;
  if (!in.wide) {
    int nn[2] = {n_other ? max(n_other - 1, 1) : 0, n_emit ? max(n_emit - 1, 1) : 0};
    total_nodes = nn[0] + nn[1];
    CK(dev_alloc((void**)&d_nodes, sizeof(float4) * 4 * (size_t)max(total_nodes, 1)));
    int base = 0;
    for (int p = 0; p < 2; p++) {
      const Part& P = parts[p];
      if (P.n == 0) continue;
      k_emit_binary<<<cdiv(nn[p], 256), 256, 0, st>>>(P.n, d_child + P.slice, d_lo + 2 * P.slice, d_hi + 2 * P.slice, base,
                                                      P.sorted_base, d_nodes);
      (p == 0 ? out->root_other : out->root_emit) = base + P.root;
      (p == 0 ? out->nodes_other : out->nodes_emit) = nn[p];
      base += nn[p];
    }
    k_iota_sorted<<<cdiv(T, 256), 256, 0, st>>>(d_final_to_sorted, T);
  } else {
    // every wide node has >= 2 children except degenerate roots, so #wide nodes <= #binary internal nodes + 2
    const size_t max_nodes = (size_t)T + 4;
    CK(dev_alloc((void**)&d_nodes, sizeof(float4) * 5 * max_nodes));
    WorkItem *d_q[2];
    int*      d_counters;
    CK(dev_alloc((void**)&d_q[0], sizeof(WorkItem) * max_nodes));
    CK(dev_alloc((void**)&d_q[1], sizeof(WorkItem) * max_nodes));
    CK(dev_alloc((void**)&d_counters, sizeof(int) * 4 + sizeof(unsigned long long)));
    unsigned long long* d_sah = reinterpret_cast<unsigned long long*>(d_counters + 4);
    CK(cudaMemsetAsync(d_sah, 0, sizeof(unsigned long long), st));
    int node_next = 0, tri_next = 0;
    for (int p = 0; p < 2; p++) {
      const Part& P = parts[p];
      if (P.n == 0) continue;
      const int root = node_next++;
      (p == 0 ? out->root_other : out->root_emit) = root;
      WorkItem h_item = {P.root, root};
      CK(cudaMemcpyAsync(d_q[0], &h_item, sizeof(h_item), cudaMemcpyHostToDevice, st));
      int n_in = 1, cur = 0;
      const int nodes_before = node_next - 1;
      while (n_in > 0) {
        int h_cnt[4] = {node_next, tri_next, 0, 0};
        CK(cudaMemcpyAsync(d_counters, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, st));
        k_collapse8<<<cdiv(n_in, 128), 128, 0, st>>>(d_q[cur], n_in, d_q[cur ^ 1], P.n, d_child + P.slice, d_range + P.slice,
                                                     d_lo + 2 * P.slice, d_hi + 2 * P.slice, d_counters, d_nodes,
                                                     d_final_to_sorted, P.sorted_base, in.leaf_sah, P.root, d_sah);
        CK(cudaMemcpyAsync(h_cnt, d_counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        node_next = h_cnt[0]; tri_next = h_cnt[1]; n_in = h_cnt[2];
        cur ^= 1;
      }
      (p == 0 ? out->nodes_other : out->nodes_emit) = node_next - nodes_before;
      if (tri_next != P.sorted_base + P.n) {
        snprintf(err, errlen, "bvh collapse: partition %d placed %d of %d triangles", p, tri_next - P.sorted_base, P.n);
        return -5;
      }
    }
    total_nodes = node_next;
    unsigned long long h_sah = 0ull;
    CK(cudaMemcpyAsync(&h_sah, d_sah, sizeof(h_sah), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out->sah_nodes_per_ray = (float)((double)h_sah / 1048576.0);
    dev_free(d_q[0]); dev_free(d_q[1]); dev_free(d_counters);
    // the node array was sized for the worst case (one wide node per triangle: 8 GB at 100M triangles); keep what is used
    // (typically T / 7 nodes) and hand the rest back
    if ((size_t)total_nodes * 4 < max_nodes) {
      float4* d_fit = nullptr;
      CK(dev_alloc((void**)&d_fit, sizeof(float4) * 5 * (size_t)std::max(total_nodes, 1)));
      CK(cudaMemcpyAsync(d_fit, d_nodes, sizeof(float4) * 5 * (size_t)total_nodes, cudaMemcpyDeviceToDevice, st));
      CK(cudaStreamSynchronize(st));
      dev_free(d_nodes);
      d_nodes = d_fit;
    }
  }

  tick("collapse");
  stage("lisa: bvh pack");
  float4 *d_tri_v, *d_tri_n;
  int*    d_final_to_orig;
  CK(dev_alloc((void**)&d_tri_v, sizeof(float4) * 3 * (size_t)T));
  CK(dev_alloc((void**)&d_tri_n, sizeof(float4) * 3 * (size_t)T));
  CK(dev_alloc((void**)&d_final_to_orig, sizeof(int) * (size_t)T));
  k_pack_triangles<<<cdiv(T, 256), 256, 0, st>>>(in.d_verts, in.d_normals, in.d_mat_idx, d_ids2, d_final_to_sorted, T,
                                                 d_tri_v, d_tri_n, d_final_to_orig, in.d_ref_tri);
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  tick("pack");
  for (int k = 1; k < sev.n; k++) {  // the stream is synchronised: every event has completed
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, sev.e[k - 1], sev.e[k]);
    const char* w = sev.name[k];
    if (!w) continue;
    if (!strcmp(w, "morton+sort")) out->stage_ms[0] = ms;
    else if (!strcmp(w, "ploc") || !strcmp(w, "lbvh")) out->stage_ms[1] = ms;
    else if (!strcmp(w, "rotations")) out->stage_ms[1] += ms;
    else if (!strcmp(w, "collapse")) out->stage_ms[2] = ms;
    else if (!strcmp(w, "pack")) out->stage_ms[3] = ms;
  }

  dev_free(d_acc); dev_free(d_keys); dev_free(d_keys2); dev_free(d_ids); dev_free(d_ids2); dev_free(d_final_to_sorted);
  dev_free(d_rec[0]); dev_free(d_rec[1]); dev_free(d_tile_state); dev_free(d_ploc_ctl);
  dev_free(d_tmp); dev_free(d_child); dev_free(d_range); dev_free(d_parent); dev_free(d_flags); dev_free(d_box);

  out->d_nodes = d_nodes;
  out->num_nodes = total_nodes;
  out->node_bytes = (size_t)total_nodes * (in.wide ? 80 : 64);
  out->d_tri_v = d_tri_v;
  out->d_tri_n = d_tri_n;
  out->d_final_to_orig = d_final_to_orig;
  return 0;
}

}  // namespace lisa
