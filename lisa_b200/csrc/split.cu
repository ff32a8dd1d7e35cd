// lisa_b200/csrc/split.cu — triangle splitting for the BVH builder (opt-in: LISA_FLAG_SPLIT_TRIANGLES).
//
// A long or large triangle has a bounding box that is mostly empty, and every ray through that box pays a node visit or a
// triangle test for nothing.  The remedy (Ernst & Greiner 2007, "Early Split Clipping") is to hand the builder several
// REFERENCES to such a triangle, each with the box of the part of the triangle inside one cell of a grid laid over the
// triangle's own bounding box: the cells measure at most L = 4 * (largest scene extent) * T^(-1/3) (four times the mean
// triangle spacing of the scene; only triangles whose box is mostly empty are split at all, see split_grid), the triangle is clipped against each cell (Sutherland-Hodgman, polygon of at most 9 vertices), and
// cells it does not reach produce nothing.  A triangle no longer than L keeps its single reference with its own box.
//
//   k_split_count   references per triangle (grid cells with a non-degenerate clipped polygon; at most SPLIT_MAX_CELLS)
//   exclusive scan  (sort_scan.cu) -> first reference of every triangle, total
//   k_split_emit    the same clipping again, writing (triangle id, box lo, box hi) per reference
// The total is kept under `budget` x T by doubling alpha until it fits (a few count + scan rounds).  The builder then
// works on references wherever it worked on triangles (bvh_build.cu: ref_tri / ref_lo / ref_hi): Morton keys from the box
// centres, leaf boxes from the reference boxes; the packed triangle arrays hold one full copy of the triangle per
// reference (the intersection test is on the whole triangle, so a hit is a hit whichever reference led to it) and
// final_to_orig maps every copy back to the caller's triangle.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "build.h"
#include "devmem.h"
#include "sort_scan.h"
#include "common.cuh"

namespace lisa {

#define SPLIT_MAX_CELLS 64

struct SplitBounds { int lo[3], hi[3]; };
__device__ __forceinline__ int   s_f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__host__ __device__ inline float s_ord2f(int i) { i = i >= 0 ? i : i ^ 0x7fffffff; float f; memcpy(&f, &i, 4); return f; }

__global__ void k_split_init(SplitBounds* b) {
  for (int a = 0; a < 3; a++) { b->lo[a] = s_f2ord(FLT_MAX); b->hi[a] = s_f2ord(-FLT_MAX); }
}
__global__ void k_split_bounds(const float* __restrict__ verts, int ntris, SplitBounds* acc) {
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntris; t += gridDim.x * blockDim.x) {
    const float* p = verts + 9ll * t;
    for (int a = 0; a < 3; a++) {
      lo[a] = fminf(lo[a], fminf(p[a], fminf(p[3 + a], p[6 + a])));
      hi[a] = fmaxf(hi[a], fmaxf(p[a], fmaxf(p[3 + a], p[6 + a])));
    }
  }
  for (int a = 0; a < 3; a++) {
    for (int o = 16; o; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&acc->lo[a], s_f2ord(lo[a])); atomicMax(&acc->hi[a], s_f2ord(hi[a])); }
  }
}

// clips polygon (n vertices) against the half space  x[axis] >= s (keep_ge) or <= s; returns the new vertex count
__device__ int clip_plane(const float3* in, int n, float3* out, int axis, float s, bool keep_ge) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    const float3 a = in[i], b = in[(i + 1 == n) ? 0 : i + 1];
    const float  fa = (axis == 0 ? a.x : axis == 1 ? a.y : a.z) - s, fb = (axis == 0 ? b.x : axis == 1 ? b.y : b.z) - s;
    const bool   ia = keep_ge ? fa >= 0.0f : fa <= 0.0f, ib = keep_ge ? fb >= 0.0f : fb <= 0.0f;
    if (ia) out[m++] = a;
    if (ia != ib) {  // the edge crosses the plane
      const float t = fa / (fa - fb);
      float3 c = f3(fmaf(t, b.x - a.x, a.x), fmaf(t, b.y - a.y, a.y), fmaf(t, b.z - a.z, a.z));
      if (axis == 0) c.x = s; else if (axis == 1) c.y = s; else c.z = s;  // exactly on the plane
      out[m++] = c;
    }
  }
  return m;
}

// The grid of a triangle: cells of at most L per axis over its bounding box, at most SPLIT_MAX_CELLS in all.
struct SplitGrid { float lo[3], hi[3], w[3]; int n[3]; };
__device__ __forceinline__ SplitGrid split_grid(const float* p, float L, float empty, int max_cells) {
  SplitGrid g;
  float ext[3];
  for (int a = 0; a < 3; a++) {
    const float l = fminf(p[a], fminf(p[3 + a], p[6 + a])), h = fmaxf(p[a], fmaxf(p[3 + a], p[6 + a]));
    g.lo[a] = l; g.hi[a] = h; ext[a] = h - l;
    g.n[a] = (int)fminf(fmaxf(ceilf(ext[a] / L), 1.0f), (float)max_cells);
  }
  // How empty is the box?  Half its surface area over the triangle's area projected on the three axis planes: 2 for an
  // axis-aligned right triangle (a wall: nothing to gain, and measured: splitting the Cornell walls costs 37 %), unbounded for
  // a diagonal sliver.  Only boxes emptier than `empty` are split.
  {
    const float e1[3] = {p[3] - p[0], p[4] - p[1], p[5] - p[2]}, e2[3] = {p[6] - p[0], p[7] - p[1], p[8] - p[2]};
    const float proj = 0.5f * (fabsf(e1[1] * e2[2] - e1[2] * e2[1]) + fabsf(e1[2] * e2[0] - e1[0] * e2[2]) + fabsf(e1[0] * e2[1] - e1[1] * e2[0]));
    const float half = ext[0] * ext[1] + ext[1] * ext[2] + ext[2] * ext[0];
    if (!(half > empty * proj)) g.n[0] = g.n[1] = g.n[2] = 1;  // (NaN from inf * 0: not split either)
  }
  while (g.n[0] * g.n[1] * g.n[2] > max_cells) {  // too many cells: coarsen the axis with the smallest cells
    int   a = -1;
    float best = FLT_MAX;
    for (int k = 0; k < 3; k++)
      if (g.n[k] > 1 && ext[k] / g.n[k] < best) { best = ext[k] / g.n[k]; a = k; }
    g.n[a] = (g.n[a] + 1) / 2;
  }
  for (int a = 0; a < 3; a++) g.w[a] = ext[a] / (float)g.n[a];
  return g;
}

// Box of the part of triangle p inside cell (ix, iy, iz); false when the triangle does not reach the cell (or only touches it).
__device__ bool split_cell_box(const float* p, const SplitGrid& g, int ix, int iy, int iz, float3& blo, float3& bhi) {
  float3 A[10], B[10];
  A[0] = f3(p[0], p[1], p[2]); A[1] = f3(p[3], p[4], p[5]); A[2] = f3(p[6], p[7], p[8]);
  int       n = 3;
  const int idx[3] = {ix, iy, iz};
  float     clo[3], chi[3];
  for (int a = 0; a < 3 && n >= 3; a++) {
    clo[a] = g.lo[a] + g.w[a] * (float)idx[a];
    chi[a] = idx[a] + 1 == g.n[a] ? FLT_MAX : g.lo[a] + g.w[a] * (float)(idx[a] + 1);  // the last cell is open towards +inf (rounding)
    if (idx[a] > 0) { n = clip_plane(A, n, B, a, clo[a], true); } else { for (int k = 0; k < n; k++) B[k] = A[k]; }
    if (n < 3) break;
    if (idx[a] + 1 < g.n[a]) { n = clip_plane(B, n, A, a, chi[a], false); } else { for (int k = 0; k < n; k++) A[k] = B[k]; }
  }
  if (n < 3) return false;
  blo = A[0]; bhi = A[0];
  for (int k = 1; k < n; k++) { blo = fmin3(blo, A[k]); bhi = fmax3(bhi, A[k]); }
  // a polygon squeezed into a line or a point (the triangle only touches the cell): the neighbouring cell covers it
  const int flat = (bhi.x <= blo.x) + (bhi.y <= blo.y) + (bhi.z <= blo.z);
  if (flat >= 2) return false;
  // conservative against the rounding of the interpolated vertices (their error scales with the triangle's coordinates, not
  // the cell's), then cut back to the triangle's own box: a reference never reaches outside its triangle's box, and an
  // axis along which the triangle is flat stays exact
  const float3 tlo = f3(g.lo[0], g.lo[1], g.lo[2]), thi = f3(g.hi[0], g.hi[1], g.hi[2]);
  const float3 pad = f3(1e-6f * fmaxf(fabsf(tlo.x), fabsf(thi.x)), 1e-6f * fmaxf(fabsf(tlo.y), fabsf(thi.y)),
                        1e-6f * fmaxf(fabsf(tlo.z), fabsf(thi.z)));
  blo = fmax3(blo - pad, tlo); bhi = fmin3(bhi + pad, thi);
  return true;
}

template <bool EMIT>
__global__ void k_split(const float* __restrict__ verts, int ntris, float L, float empty, int max_cells, uint32_t* counts, const uint32_t* __restrict__ first,
                        int* ref_tri, float4* ref_lo, float4* ref_hi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntris) return;
  const float* p = verts + 9ll * t;
  const SplitGrid g = split_grid(p, L, empty, max_cells);
  uint32_t k = 0;
  const uint32_t base = EMIT ? first[t] : 0u;
  if (g.n[0] * g.n[1] * g.n[2] == 1) {  // short enough: the triangle itself, with its own box
    if (EMIT) {
      ref_tri[base] = t;
      ref_lo[base] = make_float4(g.lo[0], g.lo[1], g.lo[2], 0.0f);
      ref_hi[base] = make_float4(g.hi[0], g.hi[1], g.hi[2], 0.0f);
    }
    k = 1;
  } else {
    for (int iz = 0; iz < g.n[2]; iz++)
      for (int iy = 0; iy < g.n[1]; iy++)
        for (int ix = 0; ix < g.n[0]; ix++) {
          float3 lo, hi;
          if (!split_cell_box(p, g, ix, iy, iz, lo, hi)) continue;
          if (EMIT) {
            ref_tri[base + k] = t;
            ref_lo[base + k] = make_float4(lo.x, lo.y, lo.z, 0.0f);
            ref_hi[base + k] = make_float4(hi.x, hi.y, hi.z, 0.0f);
          }
          k++;
        }
    if (k == 0) {  // cannot happen for a proper triangle; a degenerate one keeps its own box
      if (EMIT) {
        ref_tri[base] = t;
        ref_lo[base] = make_float4(g.lo[0], g.lo[1], g.lo[2], 0.0f);
        ref_hi[base] = make_float4(g.hi[0], g.hi[1], g.hi[2], 0.0f);
      }
      k = 1;
    }
  }
  if (!EMIT) counts[t] = k;
}

int split_triangles(const float* d_verts, int T, float budget, SplitOutput* out, cudaStream_t st, char* err, size_t errlen) {
  out->d_ref_tri = nullptr; out->d_ref_lo = out->d_ref_hi = nullptr; out->num_refs = T; out->cell = 0.0f;
  if (T == 0) return 0;
#define SCK(x)                                                                                                    \
  do {                                                                                                             \
    cudaError_t e_ = (x);                                                                                          \
    if (e_ != cudaSuccess) { snprintf(err, errlen, "%s: %s", #x, cudaGetErrorString(e_)); return -2; }             \
  } while (0)
  struct Temps {  // freed on every exit path
    SplitBounds* b = nullptr;
    uint32_t *   cnt = nullptr, *first = nullptr, *total = nullptr;
    void*        tmp = nullptr;
    ~Temps() { dev_free(b); dev_free(cnt); dev_free(first); dev_free(total); dev_free(tmp); }
  } d;
  if ((double)budget * T + 64.0 > 1.0e9) { snprintf(err, errlen, "triangle splitting: %d triangles x %.2f references exceed the builder's 2^30 primitives", T, budget); return -1; }
  // a triangle gives at most max_cells references: keep the 32-bit total of the scan exact whatever the soup
  const int max_cells = (int)std::min<long long>(SPLIT_MAX_CELLS, std::max<long long>(1, (1ll << 32) / T - 1));
  SCK(dev_alloc((void**)&d.b, sizeof(SplitBounds)));
  SCK(dev_alloc((void**)&d.cnt, sizeof(uint32_t) * (size_t)T));
  SCK(dev_alloc((void**)&d.first, sizeof(uint32_t) * (size_t)T));
  SCK(dev_alloc((void**)&d.total, sizeof(uint32_t)));
  SCK(dev_alloc(&d.tmp, scan_temp_bytes((size_t)T)));
  k_split_init<<<1, 1, 0, st>>>(d.b);
  k_split_bounds<<<std::min((T + 255) / 256, 148 * 8), 256, 0, st>>>(d_verts, T, d.b);
  SplitBounds hb;
  SCK(cudaMemcpyAsync(&hb, d.b, sizeof(hb), cudaMemcpyDeviceToHost, st));
  SCK(cudaStreamSynchronize(st));
  float maxext = 0.0f;
  for (int a = 0; a < 3; a++) maxext = std::max(maxext, s_ord2f(hb.hi[a]) - s_ord2f(hb.lo[a]));
  float L = 4.0f * maxext * powf((float)T, -1.0f / 3.0f);  // four times the mean triangle spacing
  float empty = 3.0f;
  if (const char* e = getenv("LISA_SPLIT_EMPTY")) empty = std::max(0.0f, (float)atof(e));
  if (const char* e = getenv("LISA_SPLIT_CELL")) L = std::max(1e-30f, (float)atof(e)) * maxext;
  uint32_t total = 0;
  for (int it = 0; it < 13; it++) {
    if (it == 12) empty = INFINITY;  // still over budget with cells 280x larger: one reference per triangle
    k_split<false><<<(T + 127) / 128, 128, 0, st>>>(d_verts, T, L, empty, max_cells, d.cnt, nullptr, nullptr, nullptr, nullptr);
    exclusive_scan_u32(d.cnt, d.first, (size_t)T, d.total, d.tmp, st);
    SCK(cudaMemcpyAsync(&total, d.total, sizeof(total), cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    if ((double)total <= (double)budget * T + 64.0) break;
    L *= 1.6f;  // over budget: coarser cells
  }
  if (total < (uint32_t)T) { snprintf(err, errlen, "triangle splitting lost triangles (%u references for %d)", total, T); return -5; }
  SCK(dev_alloc((void**)&out->d_ref_tri, sizeof(int) * (size_t)total));
  SCK(dev_alloc((void**)&out->d_ref_lo, sizeof(float4) * (size_t)total));
  SCK(dev_alloc((void**)&out->d_ref_hi, sizeof(float4) * (size_t)total));
  k_split<true><<<(T + 127) / 128, 128, 0, st>>>(d_verts, T, L, empty, max_cells, nullptr, d.first, out->d_ref_tri, out->d_ref_lo, out->d_ref_hi);
  SCK(cudaStreamSynchronize(st));
  SCK(cudaGetLastError());
  out->num_refs = (int)total;
  out->cell = L;
  return 0;
}

void split_free(SplitOutput* s) {
  dev_free(s->d_ref_tri); dev_free(s->d_ref_lo); dev_free(s->d_ref_hi);
  s->d_ref_tri = nullptr; s->d_ref_lo = s->d_ref_hi = nullptr;
}

}  // namespace lisa
