// lisa_b200/csrc/estimator.cuh — device code shared by the three schedules of the estimator (sched_pool.cuh,
// sched_path.cuh, sched_wavefront.cuh): flag words, chain-state loads/stores, warp helpers, the radiance arithmetic
// with its contraction spelled out, the emitter bounds / cone tests behind the exact culling of shadow tries, shading
// normal, camera ray and chain seed (shader.cu:141-152), and the reference queries used by the diagnostics.
#pragma once
#include <cstdio>
#include <cstdlib>

// The interchangeable BSDF (seam B4): chosen at compile time, like `#include "bsdfs/lambertian.cu"` in shader.cu:4
#ifndef LISA_BSDF_HEADER
#define LISA_BSDF_HEADER "bsdf/lambertian.cuh"
#endif
#include LISA_BSDF_HEADER
#include "traverse.cuh"
#include "estimator.h"

#ifndef LISA_STATE_NO_L1
#define LISA_STATE_NO_L1 0
#endif

namespace lisa {

// flags word (DState::c .w)
#define F_BOUNCE_MASK 0x000000ffu
#define F_STICKY 0x00000100u  // RayState::hit carried across bounces of one sample (Q1)
#define F_NEW 0x00000200u     // previous sample ended: regenerate a camera ray
#define F_DEFER 0x00008000u   // light sampling of this bounce continues in the NEXT iteration: k_extend skips the chain
#define F_LIGHT_SHIFT 16      // material id of the last light found (RayState::material)

#define FULL 0xffffffffu
#define SHADOW_BATCH 32
#define RING_STRIDE 16
#define R_CNTJ 0   // [p] length of the job queue of pass p
#define R_CURJ 4   // [p] fetch cursor
#define R_CNTC 8   // [p] length of the candidate queue of pass p
#define R_CURC 12  // [p] fetch cursor
#define F_TRIES_SHIFT 10
#define F_TRIES_MASK (0x1fu << F_TRIES_SHIFT)


// Chain state is streamed (read once and written once per stage): evict-first loads/stores keep it from displacing
// the BVH and the triangles in L1/L2.
__device__ __forceinline__ float4 ld_state(const float4* p) {
#if LISA_STATE_NO_L1
  float4 v;  // do not allocate the line in L1 at all: the L1 is for BVH nodes and triangles
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldcs(p);
#endif
}
__device__ __forceinline__ void   st_state(float4* p, const float4& v) { __stcs(p, v); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

enum { ST_RADIANCE = 0, ST_SHADOW = 1, ST_SAMPLES = 2, ST_NULLDIR = 3, ST_CHAINS_DONE = 4, ST_NODES = 5, ST_TRIS = 6, ST_JOBS = 7, ST_CULLED = 8 };

__device__ __forceinline__ void warp_add(unsigned long long* p, uint32_t v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  if (lane_id() == 0 && v) atomicAdd(p, (unsigned long long)v);
}

// Radiance arithmetic with the contraction spelled out, so that every kernel that accumulates a sample (both
// pipelines) rounds identically: color += e * atten;  color += (e * brdf) * atten;  sum += color.
__device__ __forceinline__ float3 add_emission(const float3& color, const float3& e, const float3& atten) {
  return f3(__fmaf_rn(e.x, atten.x, color.x), __fmaf_rn(e.y, atten.y, color.y), __fmaf_rn(e.z, atten.z, color.z));
}
__device__ __forceinline__ float3 add_light(const float3& color, const float3& e, float brdf, const float3& atten) {
  return f3(__fmaf_rn(__fmul_rn(e.x, brdf), atten.x, color.x), __fmaf_rn(__fmul_rn(e.y, brdf), atten.y, color.y),
            __fmaf_rn(__fmul_rn(e.z, brdf), atten.z, color.z));
}
__device__ __forceinline__ float3 add_sample(const float3& sum, const float3& color) {
  return f3(__fadd_rn(sum.x, color.x), __fadd_rn(sum.y, color.y), __fadd_rn(sum.z, color.z));
}

// warp-aggregated append of `id` for the lanes of the CURRENT convergent group that have push == true
__device__ __forceinline__ void queue_push(bool push, int id, int* q, unsigned int* count) {
  const unsigned am = __activemask();
  const unsigned m  = __ballot_sync(am, push);
  if (!m) return;
  const int leader = __ffs(m) - 1;
  unsigned  base = 0;
  if ((int)lane_id() == leader) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(am, base, leader);
  if (push) q[base + __popc(m & lanemask_lt())] = id;
}

// slab test of the ray against the (padded) bounds of all emitters
__device__ __forceinline__ bool hits_emitter_bounds(const DScene& sc, const float3& o, const float3& d, float tmin, float tmax) {
  const float3 id = safe_rcp_dir(d);
  const float  ax = (sc.emit_lo.x - o.x) * id.x, bx = (sc.emit_hi.x - o.x) * id.x;
  const float  ay = (sc.emit_lo.y - o.y) * id.y, by = (sc.emit_hi.y - o.y) * id.y;
  const float  az = (sc.emit_lo.z - o.z) * id.z, bz = (sc.emit_hi.z - o.z) * id.z;
  const float  tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
  const float  tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
  return tn * 0.999999f <= tf * 1.000001f;
}

template <bool WIDE, bool ANY>
__device__ __forceinline__ bool trace_one(const DScene& sc, int root, const float3& o, const float3& d, float tmin, float tmax,
                                          Hit& h, Stack& stack, uint32_t& nn, uint32_t& nt) {
  return trace_steps<WIDE, ANY>(sc.bvh, sc.tri_v, root, o, d, tmin, tmax, h, stack, nn, nt);
}

// closest hit over emitters and non-emitters (trace_radiance)
template <bool WIDE>
__device__ __forceinline__ Hit closest_hit(const DScene& sc, const float3& o, const float3& d, float tmin, float tmax,
                                           Stack& stack, uint32_t& nn, uint32_t& nt) {
  Hit he, ho;
  he.prim = -1; he.t = tmax;
  if (hits_emitter_bounds(sc, o, d, tmin, tmax)) trace_one<WIDE, false>(sc, sc.root_emit, o, d, tmin, tmax, he, stack, nn, nt);
  trace_one<WIDE, false>(sc, sc.root_other, o, d, tmin, he.t, ho, stack, nn, nt);
  return ho.prim >= 0 ? ho : he;
}

// Shadow query (trace_occlusion + the two occlusion programs, shader.cu:53-74,172-184).
// Returns 0 miss, 1 the deciding hit is an emitter (light = its material), 2 it is not.
template <bool WIDE>
__device__ __forceinline__ int shadow_query(const DScene& sc, const float3& o, const float3& d, float tmin, float tmax,
                                            int& light, Stack& stack, uint32_t& nn, uint32_t& nt) {
  Hit he, ho;
  he.prim = -1; he.t = tmax;
  const bool near_light = hits_emitter_bounds(sc, o, d, tmin, tmax);
  if (sc.shadow_first_found) {
    if (near_light && trace_one<WIDE, true>(sc, sc.root_emit, o, d, tmin, tmax, he, stack, nn, nt)) {
      light = __float_as_int(__ldg(sc.tri_v + 3 * he.prim).w);
      return 1;
    }
    return trace_one<WIDE, true>(sc, sc.root_other, o, d, tmin, tmax, ho, stack, nn, nt) ? 2 : 0;
  }
  if (near_light) trace_one<WIDE, false>(sc, sc.root_emit, o, d, tmin, tmax, he, stack, nn, nt);  // closest emitter
  if (trace_one<WIDE, true>(sc, sc.root_other, o, d, tmin, he.t, ho, stack, nn, nt)) return 2;  // any occluder in front
  if (he.prim < 0) return 0;
  light = __float_as_int(__ldg(sc.tri_v + 3 * he.prim).w);
  return 1;
}

__device__ __forceinline__ float3 shading_normal(const DScene& sc, const Hit& h) {
  // barycentric_normal (maths.cu:33-57) with the barycentrics of the intersection test
  const float4 n0 = __ldg(sc.tri_n + 3 * h.prim), n1 = __ldg(sc.tri_n + 3 * h.prim + 1), n2 = __ldg(sc.tri_n + 3 * h.prim + 2);
  const float  w0 = 1.0f - h.u - h.v;
  return normalize(madd(madd(w0 * f3(n0), h.u, f3(n1)), h.v, f3(n2)));
}

// camera ray of pixel (x, y) for the chain's next sample (shader.cu:149-152)
__device__ __forceinline__ float3 camera_ray_xy(const DCamera& cam, uint32_t x, uint32_t y, uint32_t& seed) {
  const float jx = rng(seed), jy = rng(seed);
  const float dx = (2.0f * (float)x + jx) / (float)cam.width - 1.0f;
  const float dy = (2.0f * (float)y + jy) / (float)cam.height - 1.0f;
  return normalize(madd(madd(cam.W, dy, cam.V), dx, cam.U));
}
__device__ __forceinline__ float3 camera_ray(const DCamera& cam, uint32_t p, uint32_t& seed) {
  const uint32_t x = p % cam.width, y = p / cam.width;
  return camera_ray_xy(cam, x, y, seed);
}

__device__ __forceinline__ uint32_t chain_seed(const DCamera& cam, uint32_t p, uint32_t subframe) {
  const uint32_t x = p % cam.width, y = p / cam.width;
  // shader.cu:141 — the pixel index is formed in float
  return tea16((uint32_t)((float)y * (float)cam.width + (float)x), subframe);
}

template <bool WIDE>
struct TravState;
template <>
struct TravState<true> : WideState {};
template <>
struct TravState<false> : BinState {};

// cone (axis, cos half-angle) around the sphere that bounds all emitters, seen from P
__device__ __forceinline__ void emitter_cone(const DScene& sc, const float3& P, float3& axis, float& cosa) {
  const float3 v  = sc.emit_c - P;
  const float  d2 = dot(v, v);
  if (sc.emit_r2 < 0.0f) { axis = f3(0, 0, 0); cosa = 2.0f; }                 // no emitters: nothing passes
  else if (!sc.cull || d2 <= sc.emit_r2 * 1.01f) { axis = f3(0, 0, 0); cosa = -2.0f; }  // inside the sphere / culling off
  else {
    axis = v * rsqrtf(d2);
    cosa = sqrtf(fmaxf(1.0f - sc.emit_r2 / d2, 0.0f)) - 1e-4f;
  }
}

}  // namespace lisa
