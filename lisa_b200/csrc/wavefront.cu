// lisa_b200/csrc/wavefront.cu — wavefront path tracing kernels (sm_100a).
//
// Replaces the reference's single OptiX megakernel launch (src/LiSA/src/shader.cu, 5 programs, one thread per pixel
// looping samples x bounces x <=30 shadow tries) by three stages that run once per "iteration" (= one radiance bounce
// of every live chain) over SoA chain state in HBM (DState, wavefront.cuh).  A chain is one (pixel, subframe)
// sample sequence with its own LCG stream.
//
//   k_extend  persistent CTAs; a LANE fetches a chain, regenerates a camera ray when the previous sample ended
//             (__raygen__rg, shader.cu:141-152), traverses the radiance ray (closest hit, trace_radiance
//             shader.cu:77-98) one work quantum per loop iteration, and the warp runs the material dispatch of
//             __closesthit__radiance / __miss__radiance (shader.cu:189-194, 211-246) for its finished lanes together:
//             miss and emitter end the sample, a dielectric produces the next direction in place, an opaque hit
//             stores P, N, attenuation and is appended to the job queue (or, when RayState::hit is already true, to
//             the candidate queue) by warp-ballot + prefix-popcount compaction (one atomicAdd per warp and queue).
//   k_tries   shoot_ray_to_light (shader.cu:196-209) without traversal: one WARP per job, one LANE per try (LCG
//             jump-ahead); tries that provably cannot change RayState::hit are resolved here; jobs whose 30 tries are
//             all of that kind get their BSDF bounce (lambertian.cu:7-13) and are written back; the others go to the
//             candidate queue at their first try that can reach an emitter.
//   k_rays    persistent CTAs over the candidate queue: traces that try (closest emitter, then any occluder in front
//             of it), retires it into RayState::hit, finishes lit jobs, continues the remaining tries of the others.
//
// Every chain owns its LCG stream, so the order in which chains are processed never changes a result: images are
// bit-reproducible run to run and independent of queue order, tile size and launch configuration.
#include <cstdio>

// The interchangeable BSDF (seam B4): chosen at compile time, like `#include "bsdfs/lambertian.cu"` in shader.cu:4
#ifndef LISA_BSDF_HEADER
#define LISA_BSDF_HEADER "bsdf/lambertian.cuh"
#endif
#include LISA_BSDF_HEADER
#include "traverse.cuh"
#include "wavefront.cuh"

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

// ------------------------------------------------------------------------------------------------
__global__ void k_init_chains(DState s, DCamera cam, Tile t) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    for (int k = 0; k < 3 * RING_STRIDE + 4; k++) s.ring[k] = 0;
    s.stats[ST_CHAINS_DONE] = 0;
  }
  if (i >= t.n_chains) return;
  const uint32_t p = t.pix0 + i % t.npix, f = t.f0 + i / t.npix;
  st_state(&s.a[i], make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(chain_seed(cam, p, f))));
  st_state(&s.c[i], make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(F_NEW)));
  st_state(&s.sum[i], make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(0u)));
}

template <bool WIDE>
struct TravState;
template <>
struct TravState<true> : WideState {};
template <>
struct TravState<false> : BinState {};

// ------------------------------------------------------------------------------------------------
// k_extend: persistent state machine over ALL chains of the tile.  A lane fetches a chain (warp batches of
// consecutive ids, so the state loads coalesce), regenerates a camera ray if the previous sample ended,
// traverses it one quantum per loop iteration (phase 0 closest emitter, phase 1 closest other triangle in
// front of it), and when enough lanes have finished the warp runs the material dispatch of
// __closesthit__radiance / __miss__radiance for them together and fetches new chains.
// k_extend / k_rays run 128-thread CTAs, at least 6 per SM (<= 80 registers): measured on B200 against the natural 94
// registers (5 CTAs): 6 -> +16 %, 7 (72 regs, spills) -> +14 %, 8 (64 regs) -> +12 %.  The kernels are latency bound
// (long-scoreboard stalls on chain state), so resident warps count more than a few spilled registers.
// k_rays (last pass): tries a lane may run per management section while the lanes in flight wait.  Measured on B200:
// 30 (run to completion) 826, 8 -> 836, 4 -> 831 Msamples/s.
#ifndef LISA_INLINE_TRIES
#define LISA_INLINE_TRIES 8
#endif
#ifndef LISA_VOTE_SECTIONS
#define LISA_VOTE_SECTIONS 0
#endif
#ifndef LISA_VOTE_WN
#define LISA_VOTE_WN 1
#endif
#ifndef LISA_VOTE_WT
#define LISA_VOTE_WT 1
#endif
#ifndef LISA_MIN_BLOCKS
#define LISA_MIN_BLOCKS 6
#endif
template <bool WIDE>
__global__ void __launch_bounds__(128, LISA_MIN_BLOCKS) k_extend(DScene sc, DState s, DCamera cam, Tile t, uint32_t iter, uint32_t idle_thresh) {
  extern __shared__ uint2 smem_stack[];
  Stack          stack(smem_stack);
  unsigned int*  ring = s.ring + RING_STRIDE * (iter % 3);
  const unsigned lane = lane_id();
  if (blockIdx.x == 0 && threadIdx.x < RING_STRIDE)  // reset the NEXT iteration's counters (last used two iterations ago)
    s.ring[RING_STRIDE * ((iter + 1) % 3) + threadIdx.x] = 0;
  unsigned int* cursor = &s.ring[3 * RING_STRIDE + (iter % 3)];  // chain fetch cursor of this iteration
  if (blockIdx.x == 0 && threadIdx.x == 0) s.ring[3 * RING_STRIDE + ((iter + 1) % 3)] = 0;

  int      chain = -1;
  float3   o = f3(0, 0, 0), d = f3(0, 0, 1);
  uint32_t seed = 0, flags = 0;
  bool     fresh = false, nullray = false;
  StepRay  ray;
  ray.idir = f3(0, 0, 0); ray.Sx = ray.Sy = ray.Sz = 0; ray.kz = 0; ray.oct_inv4 = 0;
  bool in_flight = false, pending = false;
  TravState<WIDE> st;
  st.begin(-1);
  int   phase = 0;
  float best_t = LISA_TMAX, best_u = 0, best_v = 0;
  int   best_prim = -1;
  unsigned wnext = 0, wend = 0;
  bool     exhausted = false;
  uint32_t n_rad = 0, n_null = 0, n_samp = 0, n_done = 0, n_jobs = 0, nn = 0, nt = 0;

  while (true) {
    const unsigned idle = __ballot_sync(FULL, !in_flight);
    if (idle == FULL || (uint32_t)__popc(idle) >= idle_thresh) {
      // ---- (1) shade the finished rays
      bool push = false, push_sticky = false;
      const int shaded = chain;
      if (pending) {
        pending = false;
        const int i = chain;
        float3 atten = f3(1.0f, 1.0f, 1.0f), color = f3(0.0f, 0.0f, 0.0f);
        if (!fresh) { atten = f3(ld_state(&s.a[i])); color = f3(ld_state(&s.c[i])); }
        bool finished = false;
        if (best_prim < 0) {
          finished = true;  // __miss__radiance (background 0, optix_wrapper.cc:354) or a null direction (Q7)
        } else {
          const int       mid = __float_as_int(__ldg(sc.tri_v + 3 * best_prim).w);
          const DMaterial m   = load_material(sc.mats, mid);
          if (m.emit()) {  // shader.cu:216-218
            color = add_emission(color, m.emission(), atten);
            finished = true;
          } else {
            const float3 P = madd(o, best_t, d);  // shader.cu:221
            Hit h;
            h.t = best_t; h.u = best_u; h.v = best_v; h.prim = best_prim;
            const float3 N = shading_normal(sc, h);
            uint32_t bounce = (flags & F_BOUNCE_MASK);
            if (m.alpha() < 1.0f) {  // dielectric, shader.cu:226-246
              float  cosI = dot(d, N), eta;
              float3 Nn;
              if (cosI < 0.0f) { cosI = -cosI; eta = 1.0f / m.ior(); Nn = N; }
              else { atten = atten * m.diffuse(); eta = m.ior(); Nn = -N; }
              float3 nd;
              if (eta == 1.0f) nd = d;
              else if (rnd(seed) <= bsdf::BTDF(cosI, eta)) nd = reflect(d, Nn);
              else nd = refract(cosI, d, Nn, eta);
              bounce++;
              if (bounce >= t.bounces) finished = true;
              else {
                flags = (flags & ~F_BOUNCE_MASK) | bounce;
                st_state(&s.o[i], make_float4(P.x, P.y, P.z, 0.0f));
                st_state(&s.d[i], make_float4(nd.x, nd.y, nd.z, 0.0f));
                st_state(&s.a[i], make_float4(atten.x, atten.y, atten.z, __uint_as_float(seed)));
                st_state(&s.c[i], make_float4(color.x, color.y, color.z, __uint_as_float(flags)));
              }
            } else {  // opaque, shader.cu:248-253: light sampling + bounce happen in k_tries / k_rays
              atten = atten * m.diffuse();
              st_state(&s.o[i], make_float4(P.x, P.y, P.z, 0.0f));
              if (fresh) st_state(&s.d[i], make_float4(d.x, d.y, d.z, 0.0f));
              st_state(&s.n[i], make_float4(N.x, N.y, N.z, __int_as_float(mid)));
              st_state(&s.a[i], make_float4(atten.x, atten.y, atten.z, __uint_as_float(seed)));
              st_state(&s.c[i], make_float4(color.x, color.y, color.z, __uint_as_float(flags & ~F_TRIES_MASK)));
              // RayState::hit already true (Q1): the first try is a real ray (it can clear hit) -> candidate queue
              push_sticky = (flags & F_STICKY) != 0;
              push = !push_sticky;
              n_jobs++;
            }
          }
        }
        if (finished) {
          const float4   sum4 = ld_state(&s.sum[i]);
          const uint32_t done = __float_as_uint(sum4.w) + 1;
          n_samp++;
          const float3 ns = add_sample(f3(sum4), color);
          st_state(&s.sum[i], make_float4(ns.x, ns.y, ns.z, __uint_as_float(done)));
          st_state(&s.a[i], make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed)));
          st_state(&s.c[i], make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(F_NEW)));
          if (done == t.spp) n_done++;
        }
        chain = -1;
      }
      // stream compaction into the queues: ballot + prefix popcount, one atomic per warp and queue
      queue_push(push, shaded, s.shadow_q, &ring[R_CNTJ + 0]);
      queue_push(push_sticky, shaded, s.cand_q, &ring[R_CNTC + 0]);
      // ---- (2) fetch chains
      const bool     need     = chain < 0;
      const unsigned needmask = __ballot_sync(FULL, need);
      if (needmask) {
        if (wnext == wend && !exhausted) {
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(cursor, 64u);
          base = __shfl_sync(FULL, base, 0);
          wnext = min(base, t.n_chains);
          wend  = min(base + 64u, t.n_chains);
          if (base + 64u >= t.n_chains) exhausted = true;
          // the batch's state (5 arrays x 64 chains x 16 B = 40 lines) is pulled into L2 now; the lanes that take
          // chains from it in later rounds then see L2 latency instead of HBM latency
          if (wnext < wend) {
            for (unsigned k = lane; k < 40u; k += 32u) {
              const float4* arr = k < 8u ? s.sum : k < 16u ? s.a : k < 24u ? s.c : k < 32u ? s.o : s.d;
              const unsigned idx = min(wnext + (k & 7u) * 8u, wend - 1u);
              prefetch_l2(arr + idx);
            }
          }
        }
        const unsigned avail = wend - wnext, cnt = __popc(needmask), rank = __popc(needmask & lanemask_lt());
        if (need && rank < avail) {
          const int    i = (int)(wnext + rank);
          // five independent loads in flight (o, d are wasted on a fresh chain; HBM is not the limit here)
          const float4 sum4 = ld_state(&s.sum[i]), a4 = ld_state(&s.a[i]), c4 = ld_state(&s.c[i]), o4 = ld_state(&s.o[i]),
                       d4 = ld_state(&s.d[i]);
          // chains that have all their samples, and chains whose light sampling is still running, are skipped
          if (__float_as_uint(sum4.w) < t.spp && !(__float_as_uint(c4.w) & F_DEFER)) {
            chain = i;
            flags = __float_as_uint(c4.w);
            seed  = __float_as_uint(a4.w);
            fresh = flags & F_NEW;
            if (fresh) {
              d = camera_ray(cam, t.pix0 + i % t.npix, seed);
              o = cam.eye;
              flags = 0;
            } else {
              o = f3(o4); d = f3(d4);
            }
            // ---- (3) start the radiance ray (trace_radiance, shader.cu:77-98)
            best_prim = -1; best_t = LISA_TMAX;
            nullray = (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f);  // Q7: refract() returned the null vector
            if (nullray) { n_null++; pending = true; }
            else {
              n_rad++;
              ray = step_ray(d);
              in_flight = true;
              stack.clear();
              if (hits_emitter_bounds(sc, o, d, LISA_TMIN, LISA_TMAX)) { phase = 0; st.begin(sc.root_emit); }
              else { phase = 1; st.begin(sc.root_other); }
              if (phase == 1 && sc.root_other < 0) { in_flight = false; pending = true; }
            }
          }
        }
        wnext += min(cnt, avail);
      }
      if (__ballot_sync(FULL, in_flight) == 0) {
        if (__ballot_sync(FULL, pending) != 0) continue;  // null rays / empty scene: shade them
        if (exhausted && wnext == wend) break;
        continue;
      }
    }
    // ---- one traversal quantum (closest hit)
#if LISA_VOTE_SECTIONS
    // the warp runs only the section (node visit / triangle tests) that more of its lanes are waiting for
    const unsigned vN = __ballot_sync(FULL, in_flight && st.has_nodes() && !st.has_tris());
    const unsigned vT = __ballot_sync(FULL, in_flight && st.has_tris());
    const bool     runN = __popc(vN) * LISA_VOTE_WN >= __popc(vT) * LISA_VOTE_WT, runT = !runN || vT == 0u;
#else
    const bool runN = true, runT = true;
#endif
    if (in_flight) {
      if (runN && st.has_nodes() && !st.has_tris()) {
        nn++;
        if (WIDE) wide_node_step(sc.bvh, o, ray, LISA_TMIN, best_t, *reinterpret_cast<WideState*>(&st), stack);
        else bin_node_step(sc.bvh, o, ray, LISA_TMIN, best_t, *reinterpret_cast<BinState*>(&st), stack);
      }
      if (!runT) {
      } else if (WIDE) {
        WideState& w = *reinterpret_cast<WideState*>(&st);
#pragma unroll
        for (int k = 0; k < LISA_TRI_PER_STEP; k++) {
          if (w.tg.y) {
            const uint32_t b = __ffs(w.tg.y) - 1u;
            w.tg.y &= w.tg.y - 1u;
            const int ti = (int)(w.tg.x + b);
            float tt, uu, vv;
            nt++;
            if (step_tri_uv(o, ray, sc.tri_v, ti, LISA_TMIN, best_t, tt, uu, vv)) { best_t = tt; best_u = uu; best_v = vv; best_prim = ti; }
          }
        }
        if (!w.has_tris() && !w.has_nodes() && !stack.empty()) w.ng = stack.pop();
      } else {
        BinState& b = *reinterpret_cast<BinState*>(&st);
        if (b.has_tris()) {
          const int ti = ~b.cur;
          float tt, uu, vv;
          nt++;
          if (step_tri_uv(o, ray, sc.tri_v, ti, LISA_TMIN, best_t, tt, uu, vv)) { best_t = tt; best_u = uu; best_v = vv; best_prim = ti; }
          b.cur = stack.empty() ? LISA_BIN_NONE : (int)stack.pop().x;
        }
      }
      if (!st.has_nodes() && !st.has_tris()) {
        if (phase == 0) {  // emitters done: now the closest other triangle in front of the closest emitter
          phase = 1;
          stack.clear();
          st.begin(sc.root_other);
          if (sc.root_other < 0) { in_flight = false; pending = true; }
        } else {
          in_flight = false;
          pending   = true;
        }
      }
    }
  }
  warp_add(&s.stats[ST_RADIANCE], n_rad);
  warp_add(&s.stats[ST_SAMPLES], n_samp);
  warp_add(&s.stats[ST_NULLDIR], n_null);
  warp_add(&s.stats[ST_CHAINS_DONE], n_done);
  warp_add(&s.stats[ST_JOBS], n_jobs);
  warp_add(&s.stats[ST_NODES], nn);
  warp_add(&s.stats[ST_TRIS], nt);
}

// ------------------------------------------------------------------------------------------------
// Light sampling (shoot_ray_to_light, shader.cu:196-209) in two kernels per pass.
//
// A shadow try can only matter if it may SET RayState::hit (the ray can reach an emitter) or CLEAR it (hit is
// currently true, Q1).  While hit is false, a try whose direction lies outside the cone around the emitter
// bounds cannot hit an emitter, so its outcome (miss or non-emitter) leaves hit false: it is resolved by
// consuming its three LCG draws, without traversal, bit-identically (LISA_FLAG_NO_CULL disables this).
//
//   k_tries  one WARP per job, one LANE per try: lane i jumps the job's LCG ahead by 3i draws (A^k, C_k per
//            lane), builds try i's direction and tests it against the cone; a ballot gives the first
//            candidate try.  Jobs without a candidate are finished here (BSDF bounce, write-back) by the lane
//            that owns them; jobs with one go to the candidate queue with the RNG state of that try.
//   k_rays   persistent state machine (one traversal quantum per iteration, dynamic job fetch): traces the
//            candidate try — phase 0 closest emitter, phase 1 any occluder in front of it — retires it into
//            RayState::hit, finishes lit jobs, and sends jobs that need more tries back to k_tries (next pass).
// Passes shrink geometrically; the last pass finishes its leftovers inline (default: a single pass, see lisa_rt.cu).
struct JobCounters { uint32_t samples, done; };

// End of the opaque branch of __closesthit__radiance for one job (shader.cu:251-252): add the light term,
// draw the BSDF bounce (also after the last bounce: it consumes RNG), end the sample or store the next ray.
__device__ __forceinline__ void finish_job(const DScene& sc, const DState& s, const Tile& t, int job, const float3& N, int mid,
                                           uint32_t seed, uint32_t flags, float brdf, JobCounters& jc) {
  const bool      lit = flags & F_STICKY;
  const float4    a4 = ld_state(&s.a[job]), c4 = ld_state(&s.c[job]), d4 = ld_state(&s.d[job]);
  const float3    atten = f3(a4);
  float3          color = f3(c4);
  const MatRef    m{sc.mats, mid};
  if (lit) {  // shader.cu:205,251: emission of the last light found (Q1) * BRDF(N, w) * attenuation
    const DMaterial lm = load_material(sc.mats, (int)(flags >> F_LIGHT_SHIFT));
    color = add_light(color, lm.emission(), brdf, atten);
  }
  const float3   nd = bsdf::bounce(f3(d4), N, seed, m);
  const uint32_t bounce = (flags & F_BOUNCE_MASK) + 1;
  if (bounce >= t.bounces) {
    const float4   sum4 = ld_state(&s.sum[job]);
    const uint32_t done = __float_as_uint(sum4.w) + 1;
    const float3   ns = add_sample(f3(sum4), color);
    st_state(&s.sum[job], make_float4(ns.x, ns.y, ns.z, __uint_as_float(done)));
    st_state(&s.a[job], make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed)));
    st_state(&s.c[job], make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(F_NEW)));
    jc.samples++;
    if (done == t.spp) jc.done++;
  } else {
    flags = (flags & ~(F_BOUNCE_MASK | F_TRIES_MASK | F_DEFER)) | bounce;
    st_state(&s.d[job], make_float4(nd.x, nd.y, nd.z, 0.0f));
    st_state(&s.a[job], make_float4(atten.x, atten.y, atten.z, __uint_as_float(seed)));
    st_state(&s.c[job], make_float4(color.x, color.y, color.z, __uint_as_float(flags)));
  }
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

__global__ void __launch_bounds__(256) k_tries(DScene sc, DState s, Tile t, uint32_t iter, uint32_t pass) {
  __shared__ uint32_t lcg_a[32], lcg_c[32];  // x -> A^(3k) x + C_(3k): jump ahead by k tries
  unsigned int*       ring = s.ring + RING_STRIDE * (iter % 3);
  const unsigned int  qn   = ring[R_CNTJ + pass];
  const unsigned      lane = lane_id();
  uint32_t my_a = 1u, my_c = 0u;
  for (unsigned k = 0; k < 3 * lane; k++) { my_c = 1664525u * my_c + 1013904223u; my_a *= 1664525u; }
  if (threadIdx.x < 32) { lcg_a[lane] = my_a; lcg_c[lane] = my_c; }
  __syncthreads();
  JobCounters jc = {0, 0};
  uint32_t    n_sh = 0, n_cull = 0;
  while (true) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&ring[R_CURJ + pass], 32u);
    base = __shfl_sync(FULL, base, 0);
    if (base >= qn) break;
    // lane j owns job j of the batch
    int      job = base + lane < qn ? s.shadow_q[base + lane] : -1;
    float3   P = f3(0, 0, 0), N = f3(0, 1, 0), axis = f3(0, 0, 0);
    float    cosa = 2.0f;
    uint32_t seed = 0, flags = 0, start = LISA_SHADOW_TRIES;
    int      mid = 0;
    if (job >= 0) {
      const float4 o4 = ld_state(&s.o[job]), n4 = ld_state(&s.n[job]), a4 = ld_state(&s.a[job]), c4 = ld_state(&s.c[job]);
      P = f3(o4); N = f3(n4);
      mid   = __float_as_int(n4.w);
      seed  = __float_as_uint(a4.w);
      flags = __float_as_uint(c4.w);
      start = (flags & F_TRIES_MASK) >> F_TRIES_SHIFT;
      emitter_cone(sc, P, axis, cosa);
      // every try lies in the hemisphere of N: if the whole cone is below that horizon no try can be a candidate
      if (cosa > -1.0f && cosa <= 1.0f && dot(N, axis) < -sqrtf(fmaxf(1.0f - cosa * cosa, 0.0f)) - 1e-3f) cosa = 2.0f;
    }
    // Pass A: the warp walks the jobs; lane i evaluates try (start + i) of job j against the cone.  Lane j keeps the
    // ballot: bit i set = try start+i of MY job points into the cone.
    unsigned cone_mask = 0;
    unsigned valid = __ballot_sync(FULL, job >= 0 && cosa <= 1.0f);  // cosa == 2: nothing can pass (no emitter in reach)
    while (valid) {
      const int j = __ffs(valid) - 1;
      valid &= valid - 1;
      const float3   Nj = f3(__shfl_sync(FULL, N.x, j), __shfl_sync(FULL, N.y, j), __shfl_sync(FULL, N.z, j));
      const float3   Aj = f3(__shfl_sync(FULL, axis.x, j), __shfl_sync(FULL, axis.y, j), __shfl_sync(FULL, axis.z, j));
      const float    cj = __shfl_sync(FULL, cosa, j);
      const uint32_t sj = __shfl_sync(FULL, seed, j), stj = __shfl_sync(FULL, start, j);
      uint32_t       sd = my_a * sj + my_c;  // LCG state before try (start + lane)
      const float3   w  = shoot_ray_hemisphere(Nj, sd);
      const unsigned m  = __ballot_sync(FULL, (stj + lane < LISA_SHADOW_TRIES) && dot(w, Aj) >= cj);
      if ((int)lane == j) cone_mask = m;
    }
    // Pass B: every lane confirms the (few) cone hits of its own job against the emitter box itself, in try order
    int first = -1;  // index (relative to start) of the first candidate try of MY job
    while (__any_sync(FULL, cone_mask != 0u && first < 0)) {
      if (cone_mask != 0u && first < 0) {
        const int b = __ffs(cone_mask) - 1;
        cone_mask &= cone_mask - 1;
        bool ok = true;
        if (sc.cull) {
          uint32_t     sd = lcg_a[b] * seed + lcg_c[b];
          const float3 w  = shoot_ray_hemisphere(N, sd);
          ok = hits_emitter_bounds(sc, P, w, LISA_TMIN, LISA_TMAX);
        }
        if (ok) first = b;
      }
    }
    // every lane settles its own job
    bool push = false;
    if (job >= 0) {
      const uint32_t consumed = first >= 0 ? (uint32_t)first : LISA_SHADOW_TRIES - start;  // tries resolved here
      n_sh += consumed;
      n_cull += consumed;
      seed = lcg_a[consumed] * seed + lcg_c[consumed];
      if (first >= 0) {  // candidate: k_rays regenerates the direction from this state
        flags = (flags & ~F_TRIES_MASK) | ((start + consumed) << F_TRIES_SHIFT);
        s.a[job].w = __uint_as_float(seed);
        s.c[job].w = __uint_as_float(flags);
        push = true;
      } else {
        finish_job(sc, s, t, job, N, mid, seed, flags, 0.0f, jc);  // 30 tries, no light (hit is false)
      }
    }
    queue_push(push, job, s.cand_q, &ring[R_CNTC + pass]);
  }
  warp_add(&s.stats[ST_SHADOW], n_sh);
  warp_add(&s.stats[ST_CULLED], n_cull);
  warp_add(&s.stats[ST_SAMPLES], jc.samples);
  warp_add(&s.stats[ST_CHAINS_DONE], jc.done);
}

template <bool WIDE>
__global__ void __launch_bounds__(128, LISA_MIN_BLOCKS) k_rays(DScene sc, DState s, Tile t, uint32_t iter, uint32_t pass, uint32_t last,
                                              uint32_t idle_thresh) {
  extern __shared__ uint2 smem_stack[];
  Stack              stack(smem_stack);
  unsigned int*      ring = s.ring + RING_STRIDE * (iter % 3);
  const unsigned int qn   = ring[R_CNTC + pass];
  const unsigned     lane = lane_id();
  // job
  int      job = -1;
  float3   P = f3(0, 0, 0), N = f3(0, 0, 0);
  uint32_t seed = 0, flags = 0, tries = 0;
  int      mid = 0;
  // ray of the current try
  StepRay ray;
  ray.idir = f3(0, 0, 0); ray.Sx = ray.Sy = ray.Sz = 0; ray.kz = 0; ray.oct_inv4 = 0;
  float brdf_w = 0.0f;  // BRDF(N, w) of the try in flight
  bool  in_flight = false, pending = false;
  // traversal
  TravState<WIDE> st;
  st.begin(-1);
  int   phase = 0;           // 0: emitter BVH (closest), 1: other BVH (any)
  float tlimit = LISA_TMAX;  // closest emitter distance found so far
  int   light_prim = -1, outcome = 0;
  // inline tries of the last pass
  float3 cone_axis = f3(0, 0, 0);
  float  cone_cos = -2.0f;
  // queue
  unsigned    wnext = 0, wend = 0;
  bool        exhausted = (qn == 0);
  JobCounters jc = {0, 0};
  uint32_t    n_sh = 0, n_cull = 0, nn = 0, nt = 0;

  while (true) {
    const unsigned idle = __ballot_sync(FULL, !in_flight);
    if (idle == FULL || (uint32_t)__popc(idle) >= idle_thresh) {
      {
        // ---- (1) retire the finished ray into RayState::hit
        bool finish = false, requeue = false;
        if (pending) {
          pending = false;
          tries++;
          if (outcome == 0) flags &= ~F_STICKY;  // __miss__occlusion
          else if (outcome == 1) {               // __closesthit__occlusion on an emitter
            const int light = __float_as_int(__ldg(sc.tri_v + 3 * light_prim).w);
            flags = (flags & 0x0000ffffu) | F_STICKY | ((uint32_t)light << F_LIGHT_SHIFT);
          }                                      // outcome 2: RayState::hit keeps its value (Q1)
          finish = (flags & F_STICKY) || tries == LISA_SHADOW_TRIES;
          requeue = !finish && last != 1u;  // last: 0 = next pass of this iteration, 1 = finish inline, 2 = next iteration
        }
        // ---- (2) last pass only: the remaining tries run here, one lane per job
        bool still_trying = false;
        if (last == 1u && job >= 0 && !in_flight && !finish) {
          still_trying = true;
          for (int k = 0; k < LISA_INLINE_TRIES; k++) {  // bounded: the lanes in flight are waiting for this section
            const uint32_t before = seed;
            const float3   w = shoot_ray_hemisphere(N, seed);
            const bool     sticky = flags & F_STICKY;
            bool           cand = sticky || dot(w, cone_axis) >= cone_cos;
            if (cand && !sticky && sc.cull) cand = hits_emitter_bounds(sc, P, w, LISA_TMIN, LISA_TMAX);
            if (cand) { seed = before; still_trying = false; break; }  // the ray is started below from `before`
            n_sh++; n_cull++;
            tries++;
            if (tries == LISA_SHADOW_TRIES) { finish = true; still_trying = false; break; }
          }
        }
        // ---- (3) finish / hand back
        if (finish) { finish_job(sc, s, t, job, N, mid, seed, flags, brdf_w, jc); job = -1; }
        if (requeue) {
          flags = (flags & ~F_TRIES_MASK) | (tries << F_TRIES_SHIFT) | (last == 2u ? F_DEFER : 0u);
          s.a[job].w = __uint_as_float(seed);
          s.c[job].w = __uint_as_float(flags);
        }
        // next pass of this iteration, or (last == 2) the job queue of the NEXT iteration: chains are independent, so a
        // bounce may take more than one iteration; its k_tries then runs at full width together with the new jobs
        queue_push(requeue, job, s.shadow_q,
                   last == 2u ? &s.ring[RING_STRIDE * ((iter + 1) % 3) + R_CNTJ + 0] : &ring[R_CNTJ + pass + 1]);
        if (requeue) job = -1;
        // ---- (4) fetch jobs
        const bool     need     = job < 0;
        const unsigned needmask = __ballot_sync(FULL, need);
        if (needmask) {
          if (wnext == wend && !exhausted) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&ring[R_CURC + pass], SHADOW_BATCH);
            base = __shfl_sync(FULL, base, 0);
            wnext = min(base, qn);
            wend  = min(base + SHADOW_BATCH, qn);
            if (base + SHADOW_BATCH >= qn) exhausted = true;
            if (wnext + lane < wend) {  // pull the state of the whole batch into L2 while the first lanes consume it
              const int pj = s.cand_q[wnext + lane];
              prefetch_l2(s.o + pj); prefetch_l2(s.n + pj); prefetch_l2(s.a + pj); prefetch_l2(s.c + pj); prefetch_l2(s.d + pj);
            }
          }
          const unsigned avail = wend - wnext, cnt = __popc(needmask), rank = __popc(needmask & lanemask_lt());
          if (need && rank < avail) {
            job = s.cand_q[wnext + rank];
            const float4 o4 = ld_state(&s.o[job]), n4 = ld_state(&s.n[job]), a4 = ld_state(&s.a[job]), c4 = ld_state(&s.c[job]);
            P = f3(o4); N = f3(n4);
            mid   = __float_as_int(n4.w);
            seed  = __float_as_uint(a4.w);
            flags = __float_as_uint(c4.w);
            tries = (flags & F_TRIES_MASK) >> F_TRIES_SHIFT;
            if (last == 1u) emitter_cone(sc, P, cone_axis, cone_cos);
          }
          wnext += min(cnt, avail);
        }
        // ---- (5) start the ray of the current try (its direction is regenerated from the stored LCG state)
        if (job >= 0 && !in_flight && !still_trying) {
          const float3 w = shoot_ray_hemisphere(N, seed);
          n_sh++;
          brdf_w = bsdf::BRDF(N, w, MatRef{sc.mats, mid});  // evaluated now (w is not kept), used if this try lights the job
          ray   = step_ray(w);
          in_flight  = true;
          light_prim = -1;
          tlimit     = LISA_TMAX;
          stack.clear();
          if (hits_emitter_bounds(sc, P, w, LISA_TMIN, LISA_TMAX)) { phase = 0; st.begin(sc.root_emit); }
          else { phase = 1; st.begin(sc.root_other); }
        }
      }
      // after the management section a lane with a job has a ray in flight (or is between two slices of its tries)
      if (__ballot_sync(FULL, in_flight) == 0) {
        if (exhausted && wnext == wend && __ballot_sync(FULL, job >= 0) == 0) break;  // queue drained, nothing pending
        continue;  // the warp's batch ran dry mid-fetch, or tries are still running: go round again
      }
    }
    // ---- one traversal quantum
#if LISA_VOTE_SECTIONS
    const unsigned vN = __ballot_sync(FULL, in_flight && st.has_nodes() && !st.has_tris());
    const unsigned vT = __ballot_sync(FULL, in_flight && st.has_tris());
    const bool     runN = __popc(vN) * LISA_VOTE_WN >= __popc(vT) * LISA_VOTE_WT, runT = !runN || vT == 0u;
#else
    const bool runN = true, runT = true;
#endif
    if (in_flight) {
      if (runN && st.has_nodes() && !st.has_tris()) {
        nn++;
        if (WIDE) wide_node_step(sc.bvh, P, ray, LISA_TMIN, tlimit, *reinterpret_cast<WideState*>(&st), stack);
        else bin_node_step(sc.bvh, P, ray, LISA_TMIN, tlimit, *reinterpret_cast<BinState*>(&st), stack);
      }
      bool occluded = false;
      if (!runT) {
      } else if (WIDE) {
        WideState& w = *reinterpret_cast<WideState*>(&st);
#pragma unroll
        for (int k = 0; k < LISA_TRI_PER_STEP; k++) {
          if (w.tg.y && !occluded) {
            const uint32_t b = __ffs(w.tg.y) - 1u;
            w.tg.y &= w.tg.y - 1u;
            const int ti = (int)(w.tg.x + b);
            float tt;
            nt++;
            if (step_tri(P, ray, sc.tri_v, ti, LISA_TMIN, tlimit, tt)) {
              if (phase == 0 && !sc.shadow_first_found) { tlimit = tt; light_prim = ti; }
              else if (phase == 0) { light_prim = ti; w.tg.y = 0; w.ng.y = 0; stack.clear(); }
              else occluded = true;
            }
          }
        }
        if (!w.has_tris() && !w.has_nodes() && !stack.empty() && !occluded) w.ng = stack.pop();
      } else {
        BinState& b = *reinterpret_cast<BinState*>(&st);
        if (b.has_tris()) {
          const int ti = ~b.cur;
          float tt;
          nt++;
          if (step_tri(P, ray, sc.tri_v, ti, LISA_TMIN, tlimit, tt)) {
            if (phase == 0 && !sc.shadow_first_found) { tlimit = tt; light_prim = ti; }
            else if (phase == 0) { light_prim = ti; stack.clear(); }
            else occluded = true;
          }
          b.cur = stack.empty() ? LISA_BIN_NONE : (int)stack.pop().x;
        }
      }
      const bool trav_done = occluded || (!st.has_nodes() && !st.has_tris());
      if (trav_done) {
        if (phase == 0 && !(sc.shadow_first_found && light_prim >= 0)) {  // emitters done: now any occluder in front
          phase = 1;
          stack.clear();
          st.begin(sc.root_other);
          if (sc.root_other < 0) { in_flight = false; pending = true; outcome = light_prim >= 0 ? 1 : 0; }
        } else {
          in_flight = false;
          pending   = true;
          outcome   = occluded ? 2 : (light_prim >= 0 ? 1 : 0);
        }
      }
    }
  }
  warp_add(&s.stats[ST_SHADOW], n_sh);
  warp_add(&s.stats[ST_CULLED], n_cull);
  warp_add(&s.stats[ST_SAMPLES], jc.samples);
  warp_add(&s.stats[ST_CHAINS_DONE], jc.done);
  warp_add(&s.stats[ST_NODES], nn);
  warp_add(&s.stats[ST_TRIS], nt);
}

// ------------------------------------------------------------------------------------------------
// k_path: the whole estimator in ONE persistent launch per tile (the default pipeline; the three kernels above are the
// wavefront pipeline, kept for ablation: LISA_PIPELINE=wavefront).
//
// A lane owns one chain from its first camera ray to its last sample; the chain's state (LCG, attenuation, radiance,
// RayState::hit, sum of finished samples) never leaves the SM, so HBM sees one 16-byte sum per chain instead of
// ~580 bytes per bounce.  The loop body is the same work quantum as above (<= 1 node visit + <= LISA_TRI_PER_STEP
// triangle tests) for the lanes with a ray in flight — radiance rays (closest hit over both BVHs) and the shadow rays of
// light-sampling candidates (closest emitter, then any occluder in front of it) share it — and, when `wait_thresh`
// lanes are waiting, one management section that runs, for those lanes together: the material dispatch of
// __closesthit__radiance (shader.cu:211-253), a slice of shoot_ray_to_light's tries with the exact cone/box culling
// described above (one lane per job, sequential LCG: no jump-ahead, no shuffles), the BSDF bounce, the end of the
// sample, the next camera ray (__raygen__rg, shader.cu:141-152) and the fetch of a new chain.
// Per-chain arithmetic and its order are those of k_extend / k_tries / k_rays: both pipelines produce bit-identical
// accumulators (tests/test_gpu_properties.py).
#ifndef LISA_PATH_MIN_BLOCKS
#define LISA_PATH_MIN_BLOCKS 5
#endif
template <bool WIDE>
__global__ void __launch_bounds__(128, LISA_PATH_MIN_BLOCKS) k_path(DScene sc, DState s, DCamera cam, Tile t, uint32_t wait_thresh) {
  extern __shared__ uint2 smem_stack[];
  __shared__ uint32_t lcg_a[32], lcg_c[32];  // x -> A^(3k) x + C_(3k): skip k tries
  __shared__ float4   jobbuf[4][96];         // per warp: 32 light-sampling jobs x (N | cos|cos|, cone axis | LCG, tries left)
  Stack          stack(smem_stack);
  const unsigned lane = lane_id();
  if (threadIdx.x < 32) {
    uint32_t a = 1u, c = 0u;
    for (unsigned k = 0; k < 3 * threadIdx.x; k++) { c = 1664525u * c + 1013904223u; a *= 1664525u; }
    lcg_a[threadIdx.x] = a; lcg_c[threadIdx.x] = c;
  }
  __syncthreads();
  unsigned int* cursor = &s.ring[0];  // zeroed by the host before the launch

  // chain
  int      chain = -1;
  uint32_t pixel = 0, done = 0, seed = 0;
  float3   sum = f3(0, 0, 0);
  // sample in flight
  float3   atten = f3(1, 1, 1), color = f3(0, 0, 0);
  uint32_t flags = 0;  // F_BOUNCE_MASK | F_STICKY | light << F_LIGHT_SHIFT
  // ray in flight
  float3  o = f3(0, 0, 0), d = f3(0, 0, 1);  // d: direction of the RADIANCE ray (kept while its shadow rays are traced)
  StepRay ray;
  ray.idir = f3(0, 0, 0); ray.Sx = ray.Sy = ray.Sz = 0; ray.kz = 0; ray.oct_inv4 = 0;
  bool shadow_ray = false, occluded = false;
  TravState<WIDE> st;
  st.begin(-1);
  int   phase = 0;
  float best_t = LISA_TMAX, best_u = 0, best_v = 0;
  int   best_prim = -1;
  // light sampling of the current opaque hit
  float3   N = f3(0, 1, 0), cone_axis = f3(0, 0, 0);
  float    cone_cos = 2.0f, brdf_w = 0.0f;
  int      mid = 0;
  uint32_t tries = 0;
  bool     in_flight = false, pending = false, trying = false;
  // chain queue
  unsigned wnext = 0, wend = 0;
  bool     exhausted = false;
  uint32_t n_rad = 0, n_null = 0, n_samp = 0, n_done = 0, n_jobs = 0, nn = 0, nt = 0, n_sh = 0, n_cull = 0;

  while (true) {
    const unsigned fly  = __ballot_sync(FULL, in_flight);
    const bool     more = !(exhausted && wnext == wend);
    const unsigned want = __ballot_sync(FULL, pending || trying || (chain < 0 && more));
    if (fly == 0u || (uint32_t)__popc(want) >= wait_thresh) {
      if (fly == 0u && want == 0u) break;
      bool end_sample = false, start_rad = false, start_shd = false, finish = false;
      float3 w = f3(0, 0, 0);  // direction of the shadow ray to start
      if (pending) {
        pending = false;
        if (!shadow_ray) {
          // ---- (1) material dispatch of the finished radiance ray
          if (best_prim < 0) {
            end_sample = true;  // __miss__radiance (background 0, optix_wrapper.cc:354) or a null direction (Q7)
          } else {
            mid = __float_as_int(__ldg(sc.tri_v + 3 * best_prim).w);
            const DMaterial m = load_material(sc.mats, mid);
            if (m.emit()) {  // shader.cu:216-218
              color = add_emission(color, m.emission(), atten);
              end_sample = true;
            } else {
              const float3 P = madd(o, best_t, d);  // shader.cu:221
              Hit h;
              h.t = best_t; h.u = best_u; h.v = best_v; h.prim = best_prim;
              N = shading_normal(sc, h);
              if (m.alpha() < 1.0f) {  // dielectric, shader.cu:226-246
                float  cosI = dot(d, N), eta;
                float3 Nn;
                if (cosI < 0.0f) { cosI = -cosI; eta = 1.0f / m.ior(); Nn = N; }
                else { atten = atten * m.diffuse(); eta = m.ior(); Nn = -N; }
                float3 nd;
                if (eta == 1.0f) nd = d;
                else if (rnd(seed) <= bsdf::BTDF(cosI, eta)) nd = reflect(d, Nn);
                else nd = refract(cosI, d, Nn, eta);
                const uint32_t bounce = (flags & F_BOUNCE_MASK) + 1;
                if (bounce >= t.bounces) end_sample = true;
                else {
                  flags = (flags & ~F_BOUNCE_MASK) | bounce;
                  o = P; d = nd;
                  start_rad = true;
                }
              } else {  // opaque, shader.cu:248-253
                atten = atten * m.diffuse();
                o = P;
                tries = 0;
                trying = true;
                n_jobs++;
                emitter_cone(sc, P, cone_axis, cone_cos);
                // every try lies in the hemisphere of N: if the whole cone is below that horizon no try can be a candidate
                if (cone_cos > -1.0f && cone_cos <= 1.0f &&
                    dot(N, cone_axis) < -sqrtf(fmaxf(1.0f - cone_cos * cone_cos, 0.0f)) - 1e-3f) cone_cos = 2.0f;
              }
            }
          }
        } else {
          // ---- (2) retire the finished shadow ray into RayState::hit
          tries++;
          if (occluded) {                  // a non-emitter decides: RayState::hit keeps its value (Q1)
          } else if (best_prim >= 0) {     // __closesthit__occlusion on an emitter
            const int light = __float_as_int(__ldg(sc.tri_v + 3 * best_prim).w);
            flags = (flags & 0x0000ffffu) | F_STICKY | ((uint32_t)light << F_LIGHT_SHIFT);
          } else flags &= ~F_STICKY;       // __miss__occlusion
          if ((flags & F_STICKY) || tries == LISA_SHADOW_TRIES) finish = true;
          else trying = true;
        }
      }
      // ---- (3) shoot_ray_to_light (shader.cu:196-209): tries that cannot change RayState::hit are resolved here,
      // all that are left of a job at once.  The warp walks its jobs; lane i evaluates try (tries + i) of job j
      // against the cone (LCG jump-ahead by 3i draws); the job's owner then confirms the few cone hits against the
      // emitter box, in try order, with the reference's own expression for the direction.
      if (trying && (flags & F_STICKY)) {  // RayState::hit is true (Q1), only at the first try of a bounce: a real ray
        w = shoot_ray_hemisphere(N, seed);
        start_shd = true; trying = false;
      } else if (trying && cone_cos > 1.0f) {  // no try can reach an emitter: consume the draws of all that are left
        const uint32_t k = LISA_SHADOW_TRIES - tries;
        seed = lcg_a[k] * seed + lcg_c[k];
        n_sh += k; n_cull += k;
        tries = LISA_SHADOW_TRIES;
        finish = true; trying = false;
      }
      const unsigned jobs = __ballot_sync(FULL, trying);
      if (jobs) {
        float4* jb = jobbuf[threadIdx.x >> 5];
        if (trying) {
          jb[3 * lane]     = make_float4(N.x, N.y, N.z, cone_cos * fabsf(cone_cos));
          jb[3 * lane + 1] = make_float4(cone_axis.x, cone_axis.y, cone_axis.z, __uint_as_float(seed));
          jb[3 * lane + 2].x = __uint_as_float(LISA_SHADOW_TRIES - tries);
        }
        __syncwarp();
        const uint32_t my_a = lcg_a[lane], my_c = lcg_c[lane];
        unsigned cone_mask = 0;
        for (unsigned rem = jobs; rem; rem &= rem - 1u) {
          const int      j  = __ffs(rem) - 1;
          const float4   r0 = jb[3 * j], r1 = jb[3 * j + 1];
          const uint32_t left = __float_as_uint(jb[3 * j + 2].x);
          uint32_t       sd = my_a * __float_as_uint(r1.w) + my_c;  // LCG state before try (tries_j + lane)
          const float    a = rng_fast(sd), b = rng_fast(sd), c = rng_fast(sd);
          // w = +-v/|v| with the sign of v.N (shoot_ray_hemisphere); w.A >= cos  <=>  q|q| >= cos|cos| * v.v with
          // q = +-v.A, no normalisation needed.  Where the sign of v.N is within rounding of zero the try is kept.
          const float    sN = fmaf(c, r0.z, fmaf(b, r0.y, a * r0.x));
          const float    qA = fmaf(c, r1.z, fmaf(b, r1.y, a * r1.x));
          const float    vv = fmaf(c, c, fmaf(b, b, a * a));
          const float    q  = __uint_as_float(__float_as_uint(qA) ^ (__float_as_uint(sN) & 0x80000000u));
          const bool     in_cone = q * fabsf(q) >= r0.w * vv || fabsf(sN) < 4e-6f;
          const unsigned m = __ballot_sync(FULL, in_cone && lane < left);
          if ((int)lane == j) cone_mask = m;
        }
        __syncwarp();
        // the few cone hits are confirmed against the emitter box, in try order, by the job's owner, with the reference's
        // own expression for the direction (measured: spreading these events over the warp's lanes gains nothing)
        int first = -1;
        if (trying) {
          while (cone_mask) {
            const int b = __ffs(cone_mask) - 1;
            cone_mask &= cone_mask - 1u;
            uint32_t     sd = lcg_a[b] * seed + lcg_c[b];
            const float3 wb = shoot_ray_hemisphere(N, sd);
            if (!sc.cull || hits_emitter_bounds(sc, o, wb, LISA_TMIN, LISA_TMAX)) { first = b; w = wb; seed = sd; break; }
          }
          const uint32_t consumed = first >= 0 ? (uint32_t)first : LISA_SHADOW_TRIES - tries;
          n_sh += consumed; n_cull += consumed;
          tries += consumed;
          if (first >= 0) start_shd = true;
          else { seed = lcg_a[consumed] * seed + lcg_c[consumed]; finish = true; }
          trying = false;
        }
      }
      // ---- (4) end of the opaque branch (shader.cu:251-252): light term, BSDF bounce
      if (finish) {
        const MatRef m{sc.mats, mid};
        if (flags & F_STICKY) {  // emission of the last light found (Q1) * BRDF(N, w) * attenuation
          const DMaterial lm = load_material(sc.mats, (int)(flags >> F_LIGHT_SHIFT));
          color = add_light(color, lm.emission(), brdf_w, atten);
        }
        const float3   nd = bsdf::bounce(d, N, seed, m);  // also after the last bounce: it consumes RNG
        const uint32_t bounce = (flags & F_BOUNCE_MASK) + 1;
        if (bounce >= t.bounces) end_sample = true;
        else {
          flags = (flags & ~F_BOUNCE_MASK) | bounce;
          d = nd;
          start_rad = true;
        }
      }
      // ---- (5) end of the sample
      bool fresh = false;
      if (end_sample) {
        sum = add_sample(sum, color);
        done++;
        n_samp++;
        if (done == t.spp) {
          st_state(&s.sum[chain], make_float4(sum.x, sum.y, sum.z, __uint_as_float(done)));
          n_done++;
          chain = -1;
        } else fresh = true;
      }
      // ---- (6) fetch chains (warp batches of consecutive ids)
      const bool     need     = chain < 0;
      const unsigned needmask = __ballot_sync(FULL, need);
      if (needmask && more) {
        if (wnext == wend) {
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(cursor, 32u);
          base = __shfl_sync(FULL, base, 0);
          wnext = min(base, t.n_chains);
          wend  = min(base + 32u, t.n_chains);
          if (base + 32u >= t.n_chains) exhausted = true;
        }
        const unsigned avail = wend - wnext, cnt = __popc(needmask), rank = __popc(needmask & lanemask_lt());
        if (need && rank < avail) {
          chain = (int)(wnext + rank);
          const uint32_t p = t.pix0 + (uint32_t)chain % t.npix;
          seed  = chain_seed(cam, p, t.f0 + (uint32_t)chain / t.npix);
          pixel = (p % cam.width) | ((p / cam.width) << 16);  // x | y << 16 (images are at most 65535 wide/high here)
          sum   = f3(0, 0, 0);
          done  = 0;
          fresh = true;
        }
        wnext += min(cnt, avail);
      }
      // ---- (7) next camera ray (shader.cu:149-152)
      if (fresh) {
        d = camera_ray_xy(cam, pixel & 0xffffu, pixel >> 16, seed);
        o = cam.eye;
        flags = 0;
        atten = f3(1.0f, 1.0f, 1.0f);
        color = f3(0.0f, 0.0f, 0.0f);
        start_rad = true;
      }
      // ---- (8) start the ray: trace_radiance (shader.cu:77-98) along d, or trace_occlusion (shader.cu:53-74) along w;
      // one section for both kinds, so that the lanes starting either run it together
      if (start_rad || start_shd) {
        const float3 rd = start_rad ? d : w;
        shadow_ray = !start_rad; occluded = false;
        best_prim = -1; best_t = LISA_TMAX;
        if (start_shd) {
          n_sh++;
          brdf_w = bsdf::BRDF(N, w, MatRef{sc.mats, mid});  // evaluated now (w is not kept), used if this try lights the job
        }
        if (start_rad && rd.x == 0.0f && rd.y == 0.0f && rd.z == 0.0f) { n_null++; pending = true; }  // Q7: refract() returned the null vector
        else {
          if (start_rad) n_rad++;
          ray = step_ray(rd);
          in_flight = true;
          stack.clear();
          if (hits_emitter_bounds(sc, o, rd, LISA_TMIN, LISA_TMAX)) { phase = 0; st.begin(sc.root_emit); }
          else { phase = 1; st.begin(sc.root_other); }
          if (phase == 1 && sc.root_other < 0) { in_flight = false; pending = true; }
        }
      }
      continue;
    }
    // ---- one traversal quantum
    if (in_flight) {
      if (st.has_nodes() && !st.has_tris()) {
        nn++;
        if (WIDE) wide_node_step(sc.bvh, o, ray, LISA_TMIN, best_t, *reinterpret_cast<WideState*>(&st), stack);
        else bin_node_step(sc.bvh, o, ray, LISA_TMIN, best_t, *reinterpret_cast<BinState*>(&st), stack);
      }
      // phase 0: closest emitter (first found under LISA_SHADOW_FIRST_FOUND); phase 1: closest other triangle in front
      // of it for a radiance ray, ANY other triangle in front of it for a shadow ray
      const bool any_hit = shadow_ray && (phase == 1 || sc.shadow_first_found);
      if (WIDE) {
        WideState& ws = *reinterpret_cast<WideState*>(&st);
#pragma unroll
        for (int k = 0; k < LISA_TRI_PER_STEP; k++) {
          if (ws.tg.y && !occluded) {
            const uint32_t b = __ffs(ws.tg.y) - 1u;
            ws.tg.y &= ws.tg.y - 1u;
            const int ti = (int)(ws.tg.x + b);
            float tt, uu, vv;
            nt++;
            if (step_tri_uv(o, ray, sc.tri_v, ti, LISA_TMIN, best_t, tt, uu, vv)) {
              if (!any_hit) { best_t = tt; best_u = uu; best_v = vv; best_prim = ti; }
              else if (phase == 0) { best_prim = ti; ws.tg.y = 0; ws.ng.y = 0; stack.clear(); }
              else occluded = true;
            }
          }
        }
        if (!ws.has_tris() && !ws.has_nodes() && !stack.empty() && !occluded) ws.ng = stack.pop();
      } else {
        BinState& b = *reinterpret_cast<BinState*>(&st);
        if (b.has_tris()) {
          const int ti = ~b.cur;
          float tt, uu, vv;
          nt++;
          bool stop = false;
          if (step_tri_uv(o, ray, sc.tri_v, ti, LISA_TMIN, best_t, tt, uu, vv)) {
            if (!any_hit) { best_t = tt; best_u = uu; best_v = vv; best_prim = ti; }
            else if (phase == 0) { best_prim = ti; stack.clear(); stop = true; }
            else occluded = true;
          }
          b.cur = (stop || stack.empty()) ? LISA_BIN_NONE : (int)stack.pop().x;
        }
      }
      if (occluded || (!st.has_nodes() && !st.has_tris())) {
        if (phase == 0 && !(any_hit && best_prim >= 0)) {  // emitters done: now the other triangles in front
          phase = 1;
          stack.clear();
          st.begin(sc.root_other);
          if (sc.root_other < 0) { in_flight = false; pending = true; }
        } else {
          in_flight = false;
          pending   = true;
        }
      }
    }
  }
  warp_add(&s.stats[ST_RADIANCE], n_rad);
  warp_add(&s.stats[ST_SHADOW], n_sh);
  warp_add(&s.stats[ST_SAMPLES], n_samp);
  warp_add(&s.stats[ST_NULLDIR], n_null);
  warp_add(&s.stats[ST_CHAINS_DONE], n_done);
  warp_add(&s.stats[ST_NODES], nn);
  warp_add(&s.stats[ST_TRIS], nt);
  warp_add(&s.stats[ST_JOBS], n_jobs);
  warp_add(&s.stats[ST_CULLED], n_cull);
}

#include "pool.cuh"

// ------------------------------------------------------------------------------------------------
// Chain sums -> accumulators.  accum.xyz += mean of every subframe of the tile, accum.w += subframes,
// in subframe order (fixed order => deterministic sums).
__global__ void k_finalize(DState s, Tile t, float4* accum) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.npix) return;
  float4 acc = accum[t.pix0 + k];
  for (uint32_t f = 0; f < t.nf; f++) {
    const float4 sm = s.sum[f * t.npix + k];
    const float  inv = 1.0f / (float)t.spp;  // shader.cu:158
    acc.x = fmaf(sm.x, inv, acc.x); acc.y = fmaf(sm.y, inv, acc.y); acc.z = fmaf(sm.z, inv, acc.z); acc.w += 1.0f;
  }
  accum[t.pix0 + k] = acc;
}

// dst += src where src may live on ANOTHER GPU: the loads go over NVLink through the peer mapping (P2P), so the
// reduce of the sample-space partition is one kernel on the root device, no staging copy
__global__ void k_accum_add(float4* dst, const float4* __restrict__ src, uint32_t npix) {
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
    const float4 b = src[p];
    float4       a = dst[p];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    dst[p] = a;
  }
}

// accumulators -> mean image (float4, alpha 1) and/or sRGB8 (shader.cu:165-166)
__global__ void k_resolve(const float4* __restrict__ accum, uint32_t npix, float4* mean_out, uint32_t* rgba8_out) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const float4 a   = accum[p];
  const float  inv = a.w > 0.0f ? 1.0f / a.w : 0.0f;
  const float3 m   = f3(a.x * inv, a.y * inv, a.z * inv);
  if (mean_out) mean_out[p] = make_float4(m.x, m.y, m.z, 1.0f);
  if (rgba8_out) rgba8_out[p] = make_color(m);
}

// ------------------------------------------------------------------------------------------------
// diagnostics
template <bool WIDE>
__global__ void k_trace_closest(DScene sc, const float* __restrict__ org, const float* __restrict__ dir, uint32_t n, float tmin,
                                float tmax, int* prim, float* tt) {
  extern __shared__ uint2 smem_stack[];
  Stack    stack(smem_stack);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t nn = 0, nt = 0;
  float3   o = f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d = f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
  Hit      h = closest_hit<WIDE>(sc, o, d, tmin, tmax, stack, nn, nt);
  prim[i] = h.prim;
  if (tt) tt[i] = h.t;
}
template <bool WIDE>
__global__ void k_trace_shadow(DScene sc, const float* __restrict__ org, const float* __restrict__ dir, uint32_t n, float tmin,
                               float tmax, int* outcome, int* light) {
  extern __shared__ uint2 smem_stack[];
  Stack    stack(smem_stack);
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t nn = 0, nt = 0;
  float3   o = f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d = f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
  int      l = -1;
  int      oc = shadow_query<WIDE>(sc, o, d, tmin, tmax, l, stack, nn, nt);
  outcome[i] = oc;
  if (light) light[i] = oc == 1 ? l : -1;
}
__global__ void k_primary_rays(DCamera cam, uint32_t subframe, float* dirs, uint32_t* seeds) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= cam.width * cam.height) return;
  uint32_t seed = chain_seed(cam, p, subframe);
  float3   d    = camera_ray(cam, p, seed);
  dirs[3 * p] = d.x; dirs[3 * p + 1] = d.y; dirs[3 * p + 2] = d.z;
  seeds[p] = seed;
}
__global__ void k_kat(int what, uint32_t n, const float* __restrict__ in_f, const uint32_t* __restrict__ in_u, float* out_f,
                      uint32_t* out_u) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (what) {
    case 0: out_u[i] = tea16(in_u[2 * i], in_u[2 * i + 1]); break;
    case 1: { uint32_t s = in_u[i]; out_f[3 * i] = rnd(s); out_f[3 * i + 1] = rnd(s); out_f[3 * i + 2] = rnd(s); out_u[i] = s; } break;
    case 2: { uint32_t s = in_u[i]; float3 h = shoot_ray_hemisphere(f3(in_f[3 * i], in_f[3 * i + 1], in_f[3 * i + 2]), s);
              out_f[3 * i] = h.x; out_f[3 * i + 1] = h.y; out_f[3 * i + 2] = h.z; out_u[i] = s; } break;
    case 3: out_f[i] = bsdf::BTDF(in_f[2 * i], in_f[2 * i + 1]); break;
    case 4: { const float* p = in_f + 8 * i; float3 r = refract(p[0], f3(p[1], p[2], p[3]), f3(p[4], p[5], p[6]), p[7]);
              out_f[3 * i] = r.x; out_f[3 * i + 1] = r.y; out_f[3 * i + 2] = r.z; } break;
    case 5: { const float* p = in_f + 7 * i; uint32_t s = in_u[i];
              // the material table for this selector is the caller's roughness array viewed as DMaterial::a.w
              float3 r = lerp(reflect(f3(p[0], p[1], p[2]), f3(p[3], p[4], p[5])), shoot_ray_hemisphere(f3(p[3], p[4], p[5]), s), p[6]);
              out_f[3 * i] = r.x; out_f[3 * i + 1] = r.y; out_f[3 * i + 2] = r.z; out_u[i] = s; } break;
    case 6: { const float* p = in_f + 6 * i; out_f[i] = bsdf::BRDF(f3(p[0], p[1], p[2]), f3(p[3], p[4], p[5]), MatRef{nullptr, 0}); } break;
    case 7: out_u[i] = make_color(f3(in_f[3 * i], in_f[3 * i + 1], in_f[3 * i + 2])); break;
    case 9: { uint32_t sd = in_u[i]; out_f[3 * i] = rng_fast(sd); out_f[3 * i + 1] = rng_fast(sd); out_f[3 * i + 2] = rng_fast(sd); out_u[i] = sd; } break;
    case 8: {  // shading normal at P: watertight-test barycentrics of a ray through P, then interpolation
      const float* p = in_f + 21 * i;
      float3 P = f3(p[0], p[1], p[2]);
      float3 v0 = f3(p[12], p[13], p[14]), v1 = f3(p[15], p[16], p[17]), v2 = f3(p[18], p[19], p[20]);
      float3 gn = normalize(cross(v1 - v0, v2 - v0));
      float3 o = P + gn, d = -gn;
      RayPre pre = ray_precompute(o, d);
      float t = 0, u = 0, v = 0;
      intersect_tri(pre, v0, v1, v2, 0.0f, 1e30f, t, u, v);
      float3 nrm = normalize((1.0f - u - v) * f3(p[3], p[4], p[5]) + u * f3(p[6], p[7], p[8]) + v * f3(p[9], p[10], p[11]));
      out_f[3 * i] = nrm.x; out_f[3 * i + 1] = nrm.y; out_f[3 * i + 2] = nrm.z;
    } break;
  }
}

// ------------------------------------------------------------------------------------------------
static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }
static inline size_t   stack_smem(int block) { return (size_t)block * LISA_STACK_SMEM_PER_THREAD; }

int configure_kernels(char* err, size_t errlen) {
  cudaError_t e = cudaSuccess;
  const int   smem = (int)stack_smem(128);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_extend<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_extend<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_rays<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_rays<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_path<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_path<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int psm = smem + 4 * (int)sizeof(PoolWarp);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pool<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pool<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, psm);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pool<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pool<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

static inline size_t pool_smem() { return stack_smem(128) + 4 * sizeof(PoolWarp); }
int pool_chains_per_cta() { return POOL_SLOTS * 4; }
int pool_occupancy(bool wide) {
  int n = 0;
  cudaError_t e = wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pool<true>, 128, pool_smem())
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pool<false>, 128, pool_smem());
  return (e == cudaSuccess && n > 0) ? n : 1;
}
void launch_pool(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, const LaunchCfg& cfg, cudaStream_t st) {
  unsigned grid = (unsigned)(cfg.sm_count * cfg.pool_blocks_per_sm);
  grid = min(grid, max(1u, cdiv(t.n_chains, POOL_SLOTS * 4u)));
  cudaMemsetAsync(s.ring, 0, sizeof(unsigned int), st);  // chain fetch cursor
  if (sc.wide) k_pool<true><<<grid, 128, pool_smem(), st>>>(sc, s, cam, t, (uint32_t)cfg.pool_dry_thresh);
  else k_pool<false><<<grid, 128, pool_smem(), st>>>(sc, s, cam, t, (uint32_t)cfg.pool_dry_thresh);
}

int path_occupancy(bool wide, int block) {
  int n = 0;
  cudaError_t e = wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_path<true>, block, stack_smem(block))
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_path<false>, block, stack_smem(block));
  return (e == cudaSuccess && n > 0) ? n : 2;
}
void launch_path(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, const LaunchCfg& cfg, cudaStream_t st) {
  const int b = 128;
  unsigned grid = (unsigned)(cfg.sm_count * cfg.path_blocks_per_sm);
  grid = min(grid, max(1u, cdiv(t.n_chains, 32u * (b / 32))));
  cudaMemsetAsync(s.ring, 0, sizeof(unsigned int), st);  // chain fetch cursor
  if (sc.wide) k_path<true><<<grid, b, stack_smem(b), st>>>(sc, s, cam, t, (uint32_t)cfg.path_wait_thresh);
  else k_path<false><<<grid, b, stack_smem(b), st>>>(sc, s, cam, t, (uint32_t)cfg.path_wait_thresh);
}

int shadow_occupancy(bool wide, int block) {
  int n = 0;
  cudaError_t e = wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_rays<true>, block, stack_smem(block))
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_rays<false>, block, stack_smem(block));
  return (e == cudaSuccess && n > 0) ? n : 2;
}
int extend_occupancy(bool wide, int block) {
  int n = 0;
  cudaError_t e = wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_extend<true>, block, stack_smem(block))
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_extend<false>, block, stack_smem(block));
  return (e == cudaSuccess && n > 0) ? n : 2;
}
int tries_occupancy(int block) {
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tries, block, 0);
  return (e == cudaSuccess && n > 0) ? n : 2;
}

void launch_init_chains(const DState& s, const DCamera& cam, const Tile& t, cudaStream_t st) {
  k_init_chains<<<cdiv(t.n_chains, 256), 256, 0, st>>>(s, cam, t);
}
void launch_extend(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, uint32_t iter, const LaunchCfg& cfg,
                   cudaStream_t st) {
  const int b = cfg.extend_block;
  unsigned grid = (unsigned)(cfg.sm_count * cfg.extend_blocks_per_sm);
  grid = min(grid, max(1u, cdiv(t.n_chains, 64u * (b / 32))));
  if (sc.wide) k_extend<true><<<grid, b, stack_smem(b), st>>>(sc, s, cam, t, iter, (uint32_t)cfg.idle_thresh);
  else k_extend<false><<<grid, b, stack_smem(b), st>>>(sc, s, cam, t, iter, (uint32_t)cfg.idle_thresh);
}
int launch_shadow(const DScene& sc, const DState& s, const Tile& t, uint32_t iter, const LaunchCfg& cfg, cudaStream_t st) {
  const int b = cfg.shadow_block;
  int launches = 0;
  const int passes = max(1, min(cfg.shadow_passes, LISA_SHADOW_PASSES));
  for (int p = 0; p < passes; p++) {
    const unsigned last = p == passes - 1 ? (cfg.defer_retries ? 2u : 1u) : 0u;
    // work shrinks roughly 4x per pass: later passes get smaller persistent grids
    unsigned gt = (unsigned)(cfg.sm_count * cfg.tries_blocks_per_sm), gr = (unsigned)(cfg.sm_count * cfg.shadow_blocks_per_sm);
    gt = min(gt, max(1u, cdiv(t.n_chains >> (2 * p), 32u * (256 / 32))));
    gr = min(gr, max(1u, cdiv(t.n_chains >> (2 * p), SHADOW_BATCH * (b / 32))));
    k_tries<<<gt, 256, 0, st>>>(sc, s, t, iter, (uint32_t)p);
    launches++;
    if (sc.wide) k_rays<true><<<gr, b, stack_smem(b), st>>>(sc, s, t, iter, (uint32_t)p, last, (uint32_t)cfg.idle_thresh_rays);
    else k_rays<false><<<gr, b, stack_smem(b), st>>>(sc, s, t, iter, (uint32_t)p, last, (uint32_t)cfg.idle_thresh_rays);
    launches++;
  }
  return launches;
}
void launch_finalize(const DState& s, const DCamera&, const Tile& t, float4* accum, cudaStream_t st) {
  k_finalize<<<cdiv(t.npix, 256), 256, 0, st>>>(s, t, accum);
}
void launch_resolve(const float4* accum, uint32_t npix, float4* mean_out, uint32_t* rgba8_out, cudaStream_t st) {
  k_resolve<<<cdiv(npix, 256), 256, 0, st>>>(accum, npix, mean_out, rgba8_out);
}
void launch_accum_add(float4* dst, const float4* src, uint32_t npix, int sm_count, cudaStream_t st) {
  k_accum_add<<<min(cdiv(npix, 256u), (unsigned)sm_count * 8u), 256, 0, st>>>(dst, src, npix);
}
void launch_trace_closest(const DScene& sc, const float* d_org, const float* d_dir, uint32_t n, float tmin, float tmax,
                          int* d_prim, float* d_t, cudaStream_t st) {
  if (!n) return;
  if (sc.wide) k_trace_closest<true><<<cdiv(n, 128), 128, stack_smem(128), st>>>(sc, d_org, d_dir, n, tmin, tmax, d_prim, d_t);
  else k_trace_closest<false><<<cdiv(n, 128), 128, stack_smem(128), st>>>(sc, d_org, d_dir, n, tmin, tmax, d_prim, d_t);
}
void launch_trace_shadow(const DScene& sc, const float* d_org, const float* d_dir, uint32_t n, float tmin, float tmax,
                         int* d_outcome, int* d_light, cudaStream_t st) {
  if (!n) return;
  if (sc.wide) k_trace_shadow<true><<<cdiv(n, 128), 128, stack_smem(128), st>>>(sc, d_org, d_dir, n, tmin, tmax, d_outcome, d_light);
  else k_trace_shadow<false><<<cdiv(n, 128), 128, stack_smem(128), st>>>(sc, d_org, d_dir, n, tmin, tmax, d_outcome, d_light);
}
void launch_primary_rays(const DCamera& cam, uint32_t subframe, float* d_dirs, uint32_t* d_seeds, cudaStream_t st) {
  k_primary_rays<<<cdiv(cam.width * cam.height, 256), 256, 0, st>>>(cam, subframe, d_dirs, d_seeds);
}
int launch_kat(int what, uint32_t n, const float* in_f, const uint32_t* in_u, float* out_f, uint32_t* out_u, cudaStream_t st) {
  if (what < 0 || what > 9) return -1;
  if (n) k_kat<<<cdiv(n, 128), 128, 0, st>>>(what, n, in_f, in_u, out_f, out_u);
  return 0;
}

}  // namespace lisa
