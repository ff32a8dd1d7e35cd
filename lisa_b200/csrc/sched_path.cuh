// lisa_b200/csrc/sched_path.cuh — k_path: the estimator in one persistent launch, one chain per lane (included by estimator.cu).
#pragma once
#include "estimator.cuh"

namespace lisa {

// ------------------------------------------------------------------------------------------------
// k_path: the whole estimator in ONE persistent launch per tile (the default for tiles too small to fill k_pool's chain
// slots, sched_pool.cuh; sched_wavefront.cuh is the same estimator as three kernels per bounce, kept for ablation).
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
template <bool WIDE, bool FLAT>
__global__ void __launch_bounds__(128, LISA_PATH_MIN_BLOCKS) k_path(DScene sc, DState s, DCamera cam, Tile t, uint32_t wait_thresh) {
  extern __shared__ uint2 smem_stack[];
  __shared__ uint32_t lcg_a[32], lcg_c[32];  // x -> A^(3k) x + C_(3k): skip k tries
  __shared__ float4   jobbuf[4][96];         // per warp: 32 light-sampling jobs x (N | cos|cos|, cone axis | LCG, tries left)
  Stack          stack(smem_stack);
  const unsigned lane = lane_id();
  fill_lcg_tables(lcg_a, lcg_c);
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
  float3   N = f3(0, 1, 0);
  float    brdf_w = 0.0f;
  int      mid = 0;
  uint32_t tries = 0;
  bool     in_flight = false, pending = false;
  // chain queue
  unsigned wnext = 0, wend = 0;
  bool     exhausted = false;
  uint32_t n_rad = 0, n_null = 0, n_samp = 0, n_done = 0, nn = 0, nt = 0;
  EventCounters ec = {0, 0, 0};

  while (true) {
    const unsigned fly  = __ballot_sync(FULL, in_flight);
    const bool     more = !(exhausted && wnext == wend);
    const unsigned want = __ballot_sync(FULL, pending || (chain < 0 && more));
    if (fly == 0u || (uint32_t)__popc(want) >= wait_thresh) {
      if (fly == 0u && want == 0u) break;
      // ---- (1)-(4) material dispatch / shadow-ray retirement, light sampling, light term + bounce (estimator.cuh)
      ChainEvent ev;
      ev.shadow_ray = shadow_ray; ev.occluded = occluded;
      ev.t = best_t; ev.u = best_u; ev.v = best_v; ev.prim = best_prim;
      ChainRegs cr;
      cr.o = o; cr.d = d; cr.atten = atten; cr.color = color; cr.N = N;
      cr.seed = seed; cr.flags = flags; cr.tries = tries; cr.mid = mid; cr.brdf_w = brdf_w;
      const bool      has_event = pending;
      pending = false;
      const ChainNext nx = chain_event<FLAT>(sc, t, lcg_a, lcg_c, jobbuf[threadIdx.x >> 5], has_event, ev, cr, ec);
      o = cr.o; d = cr.d; atten = cr.atten; color = cr.color; N = cr.N;
      seed = cr.seed; flags = cr.flags; tries = cr.tries; mid = cr.mid;
      const bool   end_sample = nx.end_sample, start_shd = nx.start_shd;
      bool         start_rad = nx.start_rad;
      const float3 w = nx.w;  // direction of the shadow ray to start
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
          ec.shadow++;
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
        LISA_COUNT(nn);
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
            LISA_COUNT(nt);
            // the outcome, by selects: closest hit so far | first-found emitter (LISA_SHADOW_FIRST_FOUND: stop) | occluder
            const bool hit = step_tri_uv(o, ray, sc.tri_v, ti, LISA_TMIN, best_t, tt, uu, vv);
            const bool take = hit & !any_hit, stop0 = hit & any_hit & (phase == 0);
            best_t = take ? tt : best_t; best_u = take ? uu : best_u; best_v = take ? vv : best_v;
            best_prim = (take | stop0) ? ti : best_prim;
            ws.tg.y = stop0 ? 0u : ws.tg.y; ws.ng.y = stop0 ? 0u : ws.ng.y; stack.sp = stop0 ? 0 : stack.sp;
            occluded |= hit & any_hit & (phase != 0);
          }
        }
        if (!ws.has_tris() && !ws.has_nodes() && !stack.empty() && !occluded) ws.ng = stack.pop();
      } else {
        BinState& b = *reinterpret_cast<BinState*>(&st);
        if (b.has_tris()) {
          const int ti = ~b.cur;
          float tt, uu, vv;
          LISA_COUNT(nt);
          bool stop = false;
          if (step_tri_uv(o, ray, sc.tri_v, ti, LISA_TMIN, best_t, tt, uu, vv)) {
            if (!any_hit) { best_t = tt; best_u = uu; best_v = vv; best_prim = ti; }
            else if (phase == 0) { best_prim = ti; stack.clear(); stop = true; }
            else occluded = true;
          }
          b.cur = (stop || stack.empty()) ? LISA_BIN_NONE : (int)stack.pop().x;
        }
      }
      if (WIDE) {  // end of a phase, by selects
        WideState& ws = *reinterpret_cast<WideState*>(&st);
        const bool done = occluded | (!ws.has_nodes() & !ws.has_tris());
        // emitters done: now the other triangles in front (if there are any)
        const bool to1 = done & (phase == 0) & !(any_hit & (best_prim >= 0)) & (sc.root_other >= 0);
        const bool fin = done & !to1;
        phase    = (done & (phase == 0) & !(any_hit & (best_prim >= 0))) ? 1 : phase;
        stack.sp = to1 ? 0 : stack.sp;
        ws.ng.x = to1 ? (uint32_t)sc.root_other : ws.ng.x; ws.ng.y = to1 ? 0x80000000u : ws.ng.y;
        ws.tg.x = to1 ? 0u : ws.tg.x;                      ws.tg.y = to1 ? 0u : ws.tg.y;
        in_flight = !fin;
        pending   = pending | fin;
      } else if (occluded || (!st.has_nodes() && !st.has_tris())) {
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
  warp_add(&s.stats[ST_SHADOW], ec.shadow);
  warp_add(&s.stats[ST_SAMPLES], n_samp);
  warp_add(&s.stats[ST_NULLDIR], n_null);
  warp_add(&s.stats[ST_CHAINS_DONE], n_done);
  warp_add(&s.stats[ST_NODES], nn);
  warp_add(&s.stats[ST_TRIS], nt);
  warp_add(&s.stats[ST_JOBS], ec.jobs);
  warp_add(&s.stats[ST_CULLED], ec.culled);
}

}  // namespace lisa
