// lisa_b200/csrc/sched_wavefront.cuh — the estimator as a wavefront pipeline (included by estimator.cu; ablation:
// LISA_FLAG_WAVEFRONT / LISA_PIPELINE=wavefront).
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
#pragma once
#include "estimator.cuh"

namespace lisa {

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
            const float3 N = shading_normal(sc, h, P);
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
        LISA_COUNT(nn);
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
            LISA_COUNT(nt);
            if (step_tri_uv(o, ray, sc.tri_v, ti, LISA_TMIN, best_t, tt, uu, vv)) { best_t = tt; best_u = uu; best_v = vv; best_prim = ti; }
          }
        }
        if (!w.has_tris() && !w.has_nodes() && !stack.empty()) w.ng = stack.pop();
      } else {
        BinState& b = *reinterpret_cast<BinState*>(&st);
        if (b.has_tris()) {
          const int ti = ~b.cur;
          float tt, uu, vv;
          LISA_COUNT(nt);
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
        LISA_COUNT(nn);
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
            LISA_COUNT(nt);
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
          LISA_COUNT(nt);
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

}  // namespace lisa
