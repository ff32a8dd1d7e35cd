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

// Per-lane counters of node visits and triangle tests (lisa_stats::nodes_visited / triangles_tested; one IADD each in the
// traversal quantum, ~1 % of the issue slots).  A production build can drop them: make EXTRA_NVFLAGS=-DLISA_COUNT_TRAVERSAL=0
// (the two statistics then read 0; ray and sample counters are kept, they live in the management section).
#ifndef LISA_COUNT_TRAVERSAL
#define LISA_COUNT_TRAVERSAL 1
#endif
#if LISA_COUNT_TRAVERSAL
#define LISA_COUNT(x) ((x)++)
#else
#define LISA_COUNT(x) ((void)0)
#endif

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

#ifndef LISA_NORMAL_FROM_P
#define LISA_NORMAL_FROM_P 0
#endif
__device__ __forceinline__ float3 shading_normal(const DScene& sc, const Hit& h, const float3& P) {
  const float4 n0 = __ldg(sc.tri_n + 3 * h.prim), n1 = __ldg(sc.tri_n + 3 * h.prim + 1), n2 = __ldg(sc.tri_n + 3 * h.prim + 2);
#if LISA_NORMAL_FROM_P
  // barycentric_normal exactly as maths.cu:33-57: barycentrics recomputed from the hit point by edge dot products
  const float3 v1 = f3(__ldg(sc.tri_v + 3 * h.prim)), v2 = f3(__ldg(sc.tri_v + 3 * h.prim + 1)), v3 = f3(__ldg(sc.tri_v + 3 * h.prim + 2));
  const float3 e1 = v2 - v1, e2 = v3 - v1, i = P - v1;
  const float  d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2), d20 = dot(i, e1), d21 = dot(i, e2);
  const float  denom = d00 * d11 - d01 * d01;
  const float  w = (d00 * d21 - d01 * d20) / denom, v = (d11 * d20 - d01 * d21) / denom, u = 1.0f - v - w;
  return normalize(madd(madd(u * f3(n0), v, f3(n1)), w, f3(n2)));
#else
  // barycentric_normal (maths.cu:33-57) with the barycentrics of the intersection test
  const float  w0 = 1.0f - h.u - h.v;
  return normalize(madd(madd(w0 * f3(n0), h.u, f3(n1)), h.v, f3(n2)));
#endif
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

// component (i mod 3) of v, by selects (a runtime index into a local array would live in local memory)
__device__ __forceinline__ float pick3(const float3& v, int i) {
  return (i == 0 || i == 3) ? v.x : ((i == 1 || i == 4) ? v.y : v.z);
}

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

// ------------------------------------------------------------------------------------------------
// One event of a chain — the management stages the persistent kernels (sched_path.cuh, sched_pool.cuh) share.
// Called by ALL 32 lanes of a warp together (it votes); a lane takes part with has_event == true.
//   in   ev        the chain's finished ray: kind, closest hit (t, u, v, prim; prim < 0 = miss), occluded (shadow rays)
//   io   c         the chain's registers (origin of the finished ray -> hit point, radiance direction, attenuation,
//                  radiance, normal, material, LCG, flags, tries, BRDF of the shadow ray that was in flight)
//   out            what the chain needs next: end of the sample, a radiance ray along c.d from c.o, or a shadow ray
//                  along w from c.o
// Stages: (1) material dispatch of a radiance hit (__closesthit__radiance / __miss__radiance, shader.cu:189-194,211-253);
// (2) retiring a shadow ray into RayState::hit (shader.cu:172-184, Q1); (3) shoot_ray_to_light (shader.cu:196-209):
// every try the job has left is evaluated at once, warp-cooperatively — the warp walks its jobs, lane i jumps the job's
// LCG ahead by 3i draws (lcg_a / lcg_c: x -> A^(3k) x + C_(3k)), draws try i's direction and tests it against the cone
// around the emitter bounds; the job's owner confirms the few cone hits against the emitter box in try order with the
// reference's own expression; tries that cannot change RayState::hit are thereby resolved without traversal, exactly;
// (4) the light term and the BSDF bounce (shader.cu:251-252).
struct ChainRegs {
  float3   o, d, atten, color, N;
  uint32_t seed, flags, tries;
  int      mid;
  float    brdf_w;
};
struct ChainEvent {
  bool  shadow_ray, occluded;
  float t, u, v;
  int   prim;
};
struct ChainNext {
  bool   end_sample, start_rad, start_shd;
  float3 w;
};
struct EventCounters { uint32_t jobs, shadow, culled; };

// FLAT: the scene's emitter bounds are flat (sc.emit_flat >= 0): the kernels are instantiated for both cases, so that each
// carries only its own filter loop.
template <bool FLAT>
__device__ __forceinline__ ChainNext chain_event(const DScene& sc, const Tile& t, const uint32_t* lcg_a, const uint32_t* lcg_c,
                                                 float4* jb /* this warp's 96 x float4 job buffer */, bool has_event,
                                                 const ChainEvent& ev, ChainRegs& c, EventCounters& cnt) {
  const unsigned lane = lane_id();
  ChainNext nx;
  nx.end_sample = nx.start_rad = nx.start_shd = false;
  nx.w = f3(0, 0, 0);
  bool finish = false, trying = false;
  if (has_event) {
    if (!ev.shadow_ray) {
      // ---- (1) material dispatch of the finished radiance ray
      if (ev.prim < 0) {
        nx.end_sample = true;  // __miss__radiance (background 0, optix_wrapper.cc:354) or a null direction (Q7)
      } else {
        c.mid = __float_as_int(__ldg(sc.tri_v + 3 * ev.prim).w);
        const DMaterial m = load_material(sc.mats, c.mid);
        if (m.emit()) {  // shader.cu:216-218
          c.color = add_emission(c.color, m.emission(), c.atten);
          nx.end_sample = true;
        } else {
          const float3 P = madd(c.o, ev.t, c.d);  // shader.cu:221
          Hit h;
          h.t = ev.t; h.u = ev.u; h.v = ev.v; h.prim = ev.prim;
          c.N = shading_normal(sc, h, P);
          if (m.alpha() < 1.0f) {  // dielectric, shader.cu:226-246
            float  cosI = dot(c.d, c.N), eta;
            float3 Nn;
            if (cosI < 0.0f) { cosI = -cosI; eta = 1.0f / m.ior(); Nn = c.N; }
            else { c.atten = c.atten * m.diffuse(); eta = m.ior(); Nn = -c.N; }
            float3 nd;
            if (eta == 1.0f) nd = c.d;
            else if (rnd(c.seed) <= bsdf::BTDF(cosI, eta)) nd = reflect(c.d, Nn);
            else nd = refract(cosI, c.d, Nn, eta);
            const uint32_t bounce = (c.flags & F_BOUNCE_MASK) + 1;
            if (bounce >= t.bounces) nx.end_sample = true;
            else {
              c.flags = (c.flags & ~F_BOUNCE_MASK) | bounce;
              c.o = P; c.d = nd;
              nx.start_rad = true;
            }
          } else {  // opaque, shader.cu:248-253
            c.atten = c.atten * m.diffuse();
            c.o = P;
            c.tries = 0;
            trying = true;
            cnt.jobs++;
          }
        }
      }
    } else {
      // ---- (2) retire the finished shadow ray into RayState::hit
      c.tries++;
      if (ev.occluded) {               // a non-emitter decides: RayState::hit keeps its value (Q1)
      } else if (ev.prim >= 0) {       // __closesthit__occlusion on an emitter
        const int light = __float_as_int(__ldg(sc.tri_v + 3 * ev.prim).w);
        c.flags = (c.flags & 0x0000ffffu) | F_STICKY | ((uint32_t)light << F_LIGHT_SHIFT);
      } else c.flags &= ~F_STICKY;     // __miss__occlusion
      if ((c.flags & F_STICKY) || c.tries == LISA_SHADOW_TRIES) finish = true;
      else trying = true;
    }
  }
  // ---- (3) shoot_ray_to_light
  // Per job, the filter that decides which tries CAN reach an emitter (everything else is resolved by consuming its draws):
  //   flat bounds (a quad light, sc.emit_flat = k): the try's direction w crosses the slab |X_k - plane| <= hh for
  //     t in [(|dy| - hh) / u, (|dy| + hh) / u] with u = +-w_k > 0 towards the plane, dy = plane - P_k; it can reach the
  //     rectangle [A0, B0] x [A2, B2] (relative to P, on the other two axes) only if, on each axis,
  //     |dy| w_i + hh |w_i| >= A_i u  and  |dy| w_i - hh |w_i| <= B_i u (hh |w_i| <= hh goes into the slack).  No division, no normalisation (homogeneous in w),
  //     and as tight as the box test itself: ~1 in 4 of the tries the bounding-sphere cone lets through.
  //   otherwise: the cone around the sphere that bounds all emitters (q|q| >= cos|cos| v.v).
  // Both are conservative with respect to hits_emitter_bounds(), which confirms the survivors in try order.
  // The job's row of the warp's job buffer is written right here (three float4: N | LCG state, the filter's four
  // parameters — flat: A0, B0, A2, B2; cone: axis.xyz, cos|cos| — and dy, the threshold for the component towards the plane
  // (-inf = every try passes), tries left, slack), so that none of it stays live in registers; rows of lanes that stop
  // trying below are never written.
  bool      nothing = false;                           // no try of this job can reach an emitter
  const int fk = sc.emit_flat;
  if (trying) {
    float4   fpar = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float    f_ady = 0.0f, f_sl = 0.0f, f_thr = 0.0f, f_dy = 0.0f;
    float    Nk = c.N.x, N0 = c.N.y, N2 = c.N.z;  // the normal as the loop wants it: permuted (k, k+1, k+2) for flat bounds
    if (sc.emit_r2 < 0.0f) nothing = true;             // no emitters at all
    else if (FLAT) {
      // components along the flat axis (k) and along the two axes of the rectangle (k + 1, k + 2 mod 3); fk is uniform
      float Pk, P0, P2, l0, h0, l2, h2;
      if (fk == 0)      { Pk = c.o.x; P0 = c.o.y; P2 = c.o.z; Nk = c.N.x; N0 = c.N.y; N2 = c.N.z; l0 = sc.emit_lo.y; h0 = sc.emit_hi.y; l2 = sc.emit_lo.z; h2 = sc.emit_hi.z; }
      else if (fk == 1) { Pk = c.o.y; P0 = c.o.z; P2 = c.o.x; Nk = c.N.y; N0 = c.N.z; N2 = c.N.x; l0 = sc.emit_lo.z; h0 = sc.emit_hi.z; l2 = sc.emit_lo.x; h2 = sc.emit_hi.x; }
      else              { Pk = c.o.z; P0 = c.o.x; P2 = c.o.y; Nk = c.N.z; N0 = c.N.x; N2 = c.N.y; l0 = sc.emit_lo.x; h0 = sc.emit_hi.x; l2 = sc.emit_lo.y; h2 = sc.emit_hi.y; }
      const float dy = sc.emit_plane - Pk;
      float A0 = l0 - P0, B0 = h0 - P0, A2 = l2 - P2, B2 = h2 - P2;
      f_ady = fabsf(dy);
      f_dy = dy;
      const float sum = fabsf(A0) + fabsf(B0) + fabsf(A2) + fabsf(B2) + f_ady;
      const float mrg = 4e-6f * sum;  // hits_emitter_bounds accepts with a relative slack of 1e-6 on its slab distances
      A0 -= mrg; B0 += mrg; A2 -= mrg; B2 += mrg;
      fpar = make_float4(A0, B0, A2, B2);
      f_sl = 1e-6f * sum + sc.emit_hh;  // the draws below are within 2^-23 of the true ones and the test is linear in them; + the slab's half thickness
      f_thr = -2.4e-7f;
      if (!sc.cull || f_ady <= sc.emit_hh) { f_thr = -INFINITY; f_sl = INFINITY; }  // culling off, or P inside the slab of the light's plane: every try passes
      else {
        // every try lies in the hemisphere of N: if the whole light is below that horizon no try can reach it
        const float top = fmaxf(N0 * A0, N0 * B0) + fmaxf(N2 * A2, N2 * B2) + Nk * dy + sc.emit_hh * fabsf(Nk);
        if (top < -1e-4f * sum) nothing = true;
      }
    } else {
      float3 cone_axis;
      float  cone_cos;
      emitter_cone(sc, c.o, cone_axis, cone_cos);
      // every try lies in the hemisphere of N: if the whole cone is below that horizon no try can be a candidate
      if (cone_cos > -1.0f && cone_cos <= 1.0f &&
          dot(c.N, cone_axis) < -sqrtf(fmaxf(1.0f - cone_cos * cone_cos, 0.0f)) - 1e-3f) cone_cos = 2.0f;
      if (cone_cos > 1.0f) nothing = true;
      fpar = make_float4(cone_axis.x, cone_axis.y, cone_axis.z, cone_cos * fabsf(cone_cos));
    }
    if (!nothing && !(c.flags & F_STICKY)) {  // the lanes that stay in `trying` below
      jb[3 * lane]     = make_float4(Nk, N0, N2, __uint_as_float(c.seed));
      jb[3 * lane + 1] = fpar;
      jb[3 * lane + 2] = make_float4(f_dy, f_thr, __uint_as_float(LISA_SHADOW_TRIES - c.tries), f_sl);
    }
  }
  if (trying && (c.flags & F_STICKY)) {  // RayState::hit is true (Q1), only at the first try of a bounce: a real ray
    nx.w = shoot_ray_hemisphere(c.N, c.seed);
    nx.start_shd = true; trying = false;
  } else if (trying && nothing) {  // no try can reach an emitter: consume the draws of all that are left
    const uint32_t k = LISA_SHADOW_TRIES - c.tries;
    c.seed = lcg_a[k] * c.seed + lcg_c[k];
    cnt.shadow += k; cnt.culled += k;
    c.tries = LISA_SHADOW_TRIES;
    finish = true; trying = false;
  }
  const unsigned jobs = __ballot_sync(FULL, trying);
  if (jobs) {
    __syncwarp();
    // The filter only has to be conservative (its hits are confirmed below with the reference's own expression), so a
    // draw is taken from the LCG state shifted left by 8 bits (t = s << 8 obeys t' = A t + (C << 8), and holds the 24
    // bits rnd() uses in its top 24): 2 + (u >> 1) 2^-22 is one IMAD.HI, minus 3 gives rng() to within 2^-23 —
    // 3 instructions on the FMA pipe per draw instead of 6, 4 of them on the busier ALU pipe.
    const uint32_t my_a = lcg_a[lane], my_c8 = lcg_c[lane] << 8;
    unsigned cone_mask = 0;
    if (FLAT) {
      // The three draws of a try are x, y, z of its direction, in that order (maths.cu:10-15); the loop wants them as
      // (k, k + 1, k + 2): it steps the LCG by the matching number of draws — multipliers and increments chosen once, per
      // kernel-uniform fk — and the job's normal is stored in the same order.  No select in the loop, and no branch: the
      // four margins are folded with min.
      const uint32_t A1 = 1664525u, A2 = 1664525u * 1664525u, A3 = 1664525u * 1664525u * 1664525u;
      const uint32_t C1 = 1013904223u << 8, C2 = (1664525u * 1013904223u + 1013904223u) << 8,
                     C3 = ((1664525u * 1664525u) * 1013904223u + 1664525u * 1013904223u + 1013904223u) << 8;
      const uint32_t mk = fk == 0 ? A1 : (fk == 1 ? A2 : A3), ck = fk == 0 ? C1 : (fk == 1 ? C2 : C3);
      const uint32_t m0 = fk == 0 ? A2 : (fk == 1 ? A3 : A1), c0i = fk == 0 ? C2 : (fk == 1 ? C3 : C1);
      const uint32_t m2 = fk == 0 ? A3 : (fk == 1 ? A1 : A2), c2i = fk == 0 ? C3 : (fk == 1 ? C1 : C2);
      for (unsigned rem = jobs; rem; rem &= rem - 1u) {
        const int      j  = __ffs(rem) - 1;
        const float4   r0 = jb[3 * j], r1 = jb[3 * j + 1], r2 = jb[3 * j + 2];
        const uint32_t left = __float_as_uint(r2.z);
        const uint32_t t0 = my_a * (__float_as_uint(r0.w) << 8) + my_c8;  // shifted LCG state before try (tries_j + lane)
        const float    vk = __uint_as_float(__umulhi(mk * t0 + ck, 0x00800000u) + 0x40000000u) - 3.0f;
        const float    v0 = __uint_as_float(__umulhi(m0 * t0 + c0i, 0x00800000u) + 0x40000000u) - 3.0f;
        const float    v2 = __uint_as_float(__umulhi(m2 * t0 + c2i, 0x00800000u) + 0x40000000u) - 3.0f;
        // w = +-v with the sign of v.N (shoot_ray_hemisphere); where that sign is within rounding of zero the try is kept
        const float    sN = fmaf(v2, r0.z, fmaf(v0, r0.y, vk * r0.x));
        const uint32_t sb = __float_as_uint(sN) & 0x80000000u;
        // r2.x = dy: its sign turns w_k into the component towards the plane, its magnitude is |dy|
        const float    u  = __uint_as_float(__float_as_uint(vk) ^ sb ^ (__float_as_uint(r2.x) & 0x80000000u));
        const float    w0 = __uint_as_float(__float_as_uint(v0) ^ sb), w2 = __uint_as_float(__float_as_uint(v2) ^ sb);
        // |w_i| <= 1 (the draws lie in [-1, 1]), so hh |w_i| <= hh: the half thickness is part of the job's slack
        const float    c0 = fabsf(r2.x) * w0, c2 = fabsf(r2.x) * w2;
        const float    g1 = fmaf(-r1.x, u, c0);   // |dy| w0 - A0 u  >= -(slack + hh)
        const float    g2 = fmaf(r1.y, u, -c0);   // B0 u - |dy| w0  >= -(slack + hh)
        const float    g3 = fmaf(-r1.z, u, c2);
        const float    g4 = fmaf(r1.w, u, -c2);
        const float    g  = fminf(fminf(g1, g2), fminf(g3, g4));
        // "every try passes" (culling off, P inside the light's slab) is in the job's data: threshold -inf, slack +inf
        const bool     keep = ((g >= -r2.w) & (u > r2.y)) | (fabsf(sN) < 4e-6f);
        const unsigned m = __ballot_sync(FULL, keep && lane < left);
        if ((int)lane == j) cone_mask = m;
      }
    } else {
      for (unsigned rem = jobs; rem; rem &= rem - 1u) {
        const int      j  = __ffs(rem) - 1;
        const float4   r0 = jb[3 * j], r1 = jb[3 * j + 1];
        const uint32_t left = __float_as_uint(jb[3 * j + 2].z);
        const uint32_t t0 = my_a * (__float_as_uint(r0.w) << 8) + my_c8;  // shifted LCG state before try (tries_j + lane)
        const uint32_t t1 = 1664525u * t0 + (1013904223u << 8);
        const uint32_t t2 = (1664525u * 1664525u) * t0 + ((1664525u * 1013904223u + 1013904223u) << 8);
        const uint32_t t3 = (1664525u * 1664525u * 1664525u) * t0 + (((1664525u * 1664525u) * 1013904223u + 1664525u * 1013904223u + 1013904223u) << 8);
        const float    a = __uint_as_float(__umulhi(t1, 0x00800000u) + 0x40000000u) - 3.0f;
        const float    b = __uint_as_float(__umulhi(t2, 0x00800000u) + 0x40000000u) - 3.0f;
        const float    cc = __uint_as_float(__umulhi(t3, 0x00800000u) + 0x40000000u) - 3.0f;
        // w = +-v/|v| with the sign of v.N (shoot_ray_hemisphere); w.A >= cos  <=>  q|q| >= cos|cos| * v.v with
        // q = +-v.A, no normalisation needed.  Where the sign of v.N is within rounding of zero the try is kept; so are
        // very short v (the draws' 2^-23 becomes a direction error above the 1e-4 the cone's cosine is widened by).
        const float    sN = fmaf(cc, r0.z, fmaf(b, r0.y, a * r0.x));
        const float    qA = fmaf(cc, r1.z, fmaf(b, r1.y, a * r1.x));
        const float    vv = fmaf(cc, cc, fmaf(b, b, a * a));
        const float    q  = __uint_as_float(__float_as_uint(qA) ^ (__float_as_uint(sN) & 0x80000000u));
        const bool     in_cone = q * fabsf(q) >= r1.w * vv || fabsf(sN) < 4e-6f || vv < 1e-4f;
        const unsigned m = __ballot_sync(FULL, in_cone && lane < left);
        if ((int)lane == j) cone_mask = m;
      }
    }
    __syncwarp();
    // the survivors are confirmed against the emitter box, in try order, by the job's owner, with the reference's own
    // expression for the direction (with flat bounds the first survivor almost always is one: ~1.1 passes of this loop
    // per management section instead of 4).  Measured twice (k_path: 1143 vs 1152; k_pool: 1279 vs 1351 Msamples/s):
    // compacting the survivors of all jobs into one list and confirming one per lane in a single pass LOSES to this loop.
    if (trying) {
      int first = -1;
      while (cone_mask) {
        const int b = __ffs(cone_mask) - 1;
        cone_mask &= cone_mask - 1u;
        uint32_t     sd = lcg_a[b] * c.seed + lcg_c[b];
        const float3 wb = shoot_ray_hemisphere(c.N, sd);
        if (!sc.cull || hits_emitter_bounds(sc, c.o, wb, LISA_TMIN, LISA_TMAX)) { first = b; nx.w = wb; c.seed = sd; break; }
      }
      const uint32_t consumed = first >= 0 ? (uint32_t)first : LISA_SHADOW_TRIES - c.tries;
      cnt.shadow += consumed; cnt.culled += consumed;
      c.tries += consumed;
      if (first >= 0) nx.start_shd = true;
      else { c.seed = lcg_a[consumed] * c.seed + lcg_c[consumed]; finish = true; }
    }
  }
  // ---- (4) end of the opaque branch (shader.cu:251-252): light term, BSDF bounce
  if (finish) {
    const MatRef m{sc.mats, c.mid};
    if (c.flags & F_STICKY) {  // emission of the last light found (Q1) * BRDF(N, w) * attenuation
      const DMaterial lm = load_material(sc.mats, (int)(c.flags >> F_LIGHT_SHIFT));
      c.color = add_light(c.color, lm.emission(), c.brdf_w, c.atten);
    }
    const float3   nd = bsdf::bounce(c.d, c.N, c.seed, m);  // also after the last bounce: it consumes RNG
    const uint32_t bounce = (c.flags & F_BOUNCE_MASK) + 1;
    if (bounce >= t.bounces) nx.end_sample = true;
    else {
      c.flags = (c.flags & ~F_BOUNCE_MASK) | bounce;
      c.d = nd;
      nx.start_rad = true;
    }
  }
  return nx;
}

// the warp's LCG jump tables: x -> A^(3k) x + C_(3k), k = threadIdx.x < 32 (call before a __syncthreads())
__device__ __forceinline__ void fill_lcg_tables(uint32_t* lcg_a, uint32_t* lcg_c) {
  if (threadIdx.x < 32) {
    uint32_t a = 1u, c = 0u;
    for (unsigned k = 0; k < 3 * threadIdx.x; k++) { c = 1664525u * c + 1013904223u; a *= 1664525u; }
    lcg_a[threadIdx.x] = a; lcg_c[threadIdx.x] = c;
  }
}

}  // namespace lisa
