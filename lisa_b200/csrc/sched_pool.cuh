// lisa_b200/csrc/sched_pool.cuh — k_pool: the persistent estimator with warp-local chain slots (included by estimator.cu).
//
// k_path ties a chain to a lane: a lane whose ray has finished waits for the warp's next management section, and that
// section runs on the waiting lanes only (profiles/r01_k_path_regions.txt: node visits at 21/32 lanes, management at
// 12/32).  Here a WARP owns 64 chains whose state lives in shared memory (8 x float4 per slot), and two warp-local
// ring queues decouple the two kinds of work:
//   ready queue    slots with a ray to trace.  A lane whose ray finishes writes the hit into the slot, appends the slot
//                  to the pending queue and takes the next ready slot at once: traversal lanes stay busy.
//   pending queue  slots with a finished ray.  When 32 are waiting the warp runs the management section with one slot per
//                  lane, i.e. at full width, whatever the lanes were tracing (their traversal state stays in registers).
// The management section is k_path's, stage for stage (material dispatch, retiring a shadow ray into RayState::hit,
// warp-cooperative tries with exact culling, light term + BSDF bounce, end of sample, camera ray, chain fetch); the ray
// set-up (reciprocal direction, shear, first phase) is done there too, at full width, and stored in the slot, so taking a
// ready slot costs three LDS.128.  Per-chain arithmetic and its order are unchanged: accumulators are bit-identical to
// k_path's and to the wavefront pipeline's (tests/test_gpu_properties.py).
//
// Slot (8 x float4, [field][slot] per warp):
//   A  o.xyz | LCG state            B  d.xyz (radiance direction) | flags (bounce, RayState::hit, light)
//   C  attenuation | mid + tries<<16  D  radiance of the sample | BRDF(N, w) of the shadow ray in flight
//   E  N.xyz | ray kind (bit 0: shadow ray, bit 1: first phase is the other-BVH)
//   F  1/dir.xyz | octant x4   -> after the traversal: t, u, v, prim (prim -1 miss, -2 occluded)
//   H  Sx, Sy, Sz | kz          G  chain id, pixel (x | y<<16), finished samples, -
// The sum of a chain's finished samples lives in s.sum[chain] (read-modify-write once per sample, L2 resident).
#pragma once
#include "estimator.cuh"

namespace lisa {

#ifndef LISA_POOL_MIN_BLOCKS
#define LISA_POOL_MIN_BLOCKS 5
#endif
// Two flavours of the kernel, chosen per scene by the builder's surface-area estimate of the node visits per ray
// (BuildOutput::sah_nodes_per_ray, lisa_rt.cu):
//   shallow  64 chains per warp, 5 traversal-stack entries per thread in shared memory.  Scenes whose rays visit a
//            few nodes (the Cornell box: 2, the 871k-triangle knot: 2.2): the management section wants 32 pending slots often.
//   deep     48 chains per warp and 12 stack entries: scenes whose rays visit tens of nodes (the C4 soups: 30) push deep, and
//            what does not fit in shared memory spills to the local array — through an L1 that the chain slots leave ~30 KB
//            of; measured on the 10M soup: 63.7 -> 72.4 Msamples/s, and -9 % on the knot, hence two flavours.
// Both run 5 CTAs per SM (44.8 KB / 43.3 KB of shared memory per CTA) and are bit-identical in their results.
#ifndef POOL_SLOTS
#define POOL_SLOTS 64  // chains per warp, shallow flavour (ring positions wrap with % SLOTS: a power of two costs one AND)
#endif
#ifndef POOL_SLOTS_DEEP
#define POOL_SLOTS_DEEP 48
#endif
#ifndef POOL_STACK_SM
#define POOL_STACK_SM 5  // what fits beside 5 CTAs x 64 chains per warp (4 -> 5: knot +1.5 %, Cornell +0.1 %)
#endif
#ifndef POOL_STACK_SM_DEEP
#define POOL_STACK_SM_DEEP 12
#endif

template <int SLOTS>
struct PoolWarp {
  float4        A[SLOTS], B[SLOTS], C[SLOTS], D[SLOTS], E[SLOTS], F[SLOTS], H[SLOTS], G[SLOTS];
  unsigned char rq[SLOTS], pq[SLOTS];
};
template <bool DEEP>
struct PoolShape {
  static constexpr int SLOTS = DEEP ? POOL_SLOTS_DEEP : POOL_SLOTS;
  static constexpr int NS0   = DEEP ? POOL_STACK_SM_DEEP : POOL_STACK_SM;
  static constexpr int NS    = NS0 < LISA_STACK_TOTAL ? NS0 : LISA_STACK_TOTAL - 1;  // (the tiny-stack test build has 3 entries in all)
  typedef TravStack<uint2, NS, LISA_STACK_TOTAL - NS> Stk;
  static constexpr size_t smem_bytes = (size_t)NS * 128 * sizeof(uint2) + 4 * sizeof(PoolWarp<SLOTS>);
};

template <bool WIDE, bool FLAT, bool DEEP>
__global__ void __launch_bounds__(128, LISA_POOL_MIN_BLOCKS) k_pool(DScene sc, DState s, DCamera cam, Tile t, uint32_t dry_thresh) {
  constexpr int POOL_N = PoolShape<DEEP>::SLOTS;
  typedef typename PoolShape<DEEP>::Stk Stack;
  typedef lisa::PoolWarp<POOL_N> PoolWarp;
  extern __shared__ uint2 smem_stack[];  // [NS][128] traversal stacks, then 4 x PoolWarp
  __shared__ uint32_t lcg_a[32], lcg_c[32];  // x -> A^(3k) x + C_(3k): skip k tries
  __shared__ float4   jobbuf[4][96];         // per warp: 32 light-sampling jobs x (N | cos|cos|, cone axis | LCG, tries left)
  Stack          stack(smem_stack);
  PoolWarp&      pw   = reinterpret_cast<PoolWarp*>(smem_stack + PoolShape<DEEP>::NS * 128)[threadIdx.x >> 5];
  const unsigned lane = lane_id();
  fill_lcg_tables(lcg_a, lcg_c);
  // every slot starts in the pending queue without a chain: the first management sections fetch chains for them
  for (unsigned k = lane; k < POOL_N; k += 32) {
    pw.pq[k] = (unsigned char)k;
    pw.G[k]  = make_float4(__int_as_float(-1), 0.0f, 0.0f, 0.0f);
    pw.F[k]  = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
    pw.E[k]  = make_float4(0.0f, 1.0f, 0.0f, __uint_as_float(0u));
  }
  __syncthreads();
  unsigned int* cursor = &s.ring[0];  // zeroed by the host before the launch

  // queues (warp-uniform)
  unsigned rq_head = 0, rq_cnt = 0, pq_head = 0, pq_cnt = POOL_N;
  // ray in flight on this lane
  int      slot = -1;
  bool     in_flight = false, shadow_ray = false, occluded = false;
  float3   o = f3(0, 0, 0);
  StepRay  ray;
  ray.idir = f3(0, 0, 0); ray.Sx = ray.Sy = ray.Sz = 0; ray.kz = 0; ray.oct_inv4 = 0;
  TravState<WIDE> st;
  st.begin(-1);
  int   phase = 0;
  float best_t = LISA_TMAX, best_u = 0, best_v = 0;
  int   best_prim = -1;
  // chain queue
  unsigned wnext = 0, wend = 0;
  bool     exhausted = false;
  uint32_t n_rad = 0, n_null = 0, n_samp = 0, n_done = 0, nn = 0, nt = 0;
  EventCounters ec = {0, 0, 0};

  while (true) {
    // ---- (a) lanes without a ray take ready slots
    const unsigned freemask = __ballot_sync(FULL, !in_flight);
    if (freemask && rq_cnt) {
      const unsigned rank = __popc(freemask & lanemask_lt());
      if (!in_flight && rank < rq_cnt) {
        slot = pw.rq[(rq_head + rank) % POOL_N];
        const float4   a4 = pw.A[slot], f4 = pw.F[slot], h4 = pw.H[slot];
        const uint32_t kind = __float_as_uint(pw.E[slot].w);
        o = f3(a4);
        ray.idir = f3(f4); ray.oct_inv4 = __float_as_uint(f4.w);
        ray.Sx = h4.x; ray.Sy = h4.y; ray.Sz = h4.z; ray.kz = __float_as_int(h4.w);
        shadow_ray = kind & 1u; occluded = false;
        best_prim = -1; best_t = LISA_TMAX;
        in_flight = true;
        stack.clear();
        if (kind & 2u) { phase = 1; st.begin(sc.root_other); }
        else { phase = 0; st.begin(sc.root_emit); }
      }
      const unsigned n = min((unsigned)__popc(freemask), rq_cnt);
      rq_head += n; rq_cnt -= n;
    }
    const unsigned fly = __ballot_sync(FULL, in_flight);
    // ---- (b) management section: 32 pending slots, or fewer when the ready queue has run dry
    if (pq_cnt >= 32u || (rq_cnt == 0u && (pq_cnt >= dry_thresh || (fly == 0u && pq_cnt > 0u)))) {
      const unsigned np = min(pq_cnt, 32u);
      const bool     active = lane < np;
      const int      ms = active ? (int)pw.pq[(pq_head + lane) % POOL_N] : -1;
      pq_head += np; pq_cnt -= np;
      // slot -> registers
      int      chain = -1;
      uint32_t pixel = 0, done = 0, seed = 0, flags = 0, tries = 0, kind = 0;
      int      mid = 0;
      float3   mo = f3(0, 0, 0), d = f3(0, 0, 1), atten = f3(1, 1, 1), color = f3(0, 0, 0), N = f3(0, 1, 0);
      float    brdf_w = 0.0f, r_t = LISA_TMAX, r_u = 0.0f, r_v = 0.0f;
      int      r_prim = -1;
      if (active) {
        const float4 a4 = pw.A[ms], b4 = pw.B[ms], c4 = pw.C[ms], d4 = pw.D[ms], e4 = pw.E[ms], f4 = pw.F[ms], g4 = pw.G[ms];
        mo = f3(a4); seed = __float_as_uint(a4.w);
        d = f3(b4); flags = __float_as_uint(b4.w);
        atten = f3(c4); mid = (int)(__float_as_uint(c4.w) & 0xffffu); tries = __float_as_uint(c4.w) >> 16;
        color = f3(d4); brdf_w = d4.w;
        N = f3(e4); kind = __float_as_uint(e4.w);
        r_t = f4.x; r_u = f4.y; r_v = f4.z; r_prim = __float_as_int(f4.w);
        chain = __float_as_int(g4.x); pixel = __float_as_uint(g4.y); done = __float_as_uint(g4.z);
      }
      // ---- (1)-(4) material dispatch / shadow-ray retirement, light sampling, light term + bounce (estimator.cuh)
      ChainEvent ev;
      ev.shadow_ray = kind & 1u; ev.occluded = r_prim == -2;
      ev.t = r_t; ev.u = r_u; ev.v = r_v; ev.prim = r_prim;
      ChainRegs cr;
      cr.o = mo; cr.d = d; cr.atten = atten; cr.color = color; cr.N = N;
      cr.seed = seed; cr.flags = flags; cr.tries = tries; cr.mid = mid; cr.brdf_w = brdf_w;
      const ChainNext nx = chain_event<FLAT>(sc, t, lcg_a, lcg_c, jobbuf[threadIdx.x >> 5], active && chain >= 0, ev, cr, ec);
      mo = cr.o; d = cr.d; atten = cr.atten; color = cr.color; N = cr.N;
      seed = cr.seed; flags = cr.flags; tries = cr.tries; mid = cr.mid;
      const bool   end_sample = nx.end_sample, start_shd = nx.start_shd;
      bool         start_rad = nx.start_rad;
      const float3 w = nx.w;  // direction of the shadow ray to start
      // ---- (5) end of the sample: the chain's sum lives in s.sum[chain]
      bool fresh = false;
      if (end_sample) {
        float3 sum = f3(0, 0, 0);
        if (done) sum = f3(__ldcg(&s.sum[chain]));
        sum = add_sample(sum, color);
        done++;
        n_samp++;
        __stcg(&s.sum[chain], make_float4(sum.x, sum.y, sum.z, __uint_as_float(done)));
        if (done == t.spp) { n_done++; chain = -1; }
        else fresh = true;
      }
      // ---- (6) fetch chains (warp batches of consecutive ids)
      const bool     more     = !(exhausted && wnext == wend);
      const bool     need     = active && chain < 0;
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
          pixel = (p % cam.width) | ((p / cam.width) << 16);
          done  = 0;
          fresh = true;
        }
        wnext += min(cnt, avail);
      }
      // ---- (7) next camera ray (shader.cu:149-152)
      if (fresh) {
        d = camera_ray_xy(cam, pixel & 0xffffu, pixel >> 16, seed);
        mo = cam.eye;
        flags = 0;
        atten = f3(1.0f, 1.0f, 1.0f);
        color = f3(0.0f, 0.0f, 0.0f);
        start_rad = true;
      }
      // ---- (8) set the ray up and hand the slot over: ready queue, or straight back to the pending queue when there is
      // nothing to trace (null direction, Q7) or the slot still needs a chain; a slot without a chain when none is left dies
      bool to_ready = false, to_pending = false;
      if (active) {
        float4 f4 = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1)), h4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (start_rad || start_shd) {
          const float3 rd = start_rad ? d : w;
          kind = start_rad ? 0u : 1u;
          if (start_shd) {
            ec.shadow++;
            brdf_w = bsdf::BRDF(N, w, MatRef{sc.mats, mid});  // evaluated now (w is not kept), used if this try lights the job
          }
          if (start_rad && rd.x == 0.0f && rd.y == 0.0f && rd.z == 0.0f) { n_null++; to_pending = true; }
          else {
            if (start_rad) n_rad++;
            if (!hits_emitter_bounds(sc, mo, rd, LISA_TMIN, LISA_TMAX)) kind |= 2u;
            // nothing to traverse (no emitter in reach and no other triangle at all): the slot goes straight back to the
            // pending queue and F must keep the MISS record (0, 0, 0, -1) — the ray set-up is stored only for the ready queue
            if ((kind & 2u) && sc.root_other < 0) to_pending = true;
            else {
              const StepRay r = step_ray(rd);
              f4 = make_float4(r.idir.x, r.idir.y, r.idir.z, __uint_as_float(r.oct_inv4));
              h4 = make_float4(r.Sx, r.Sy, r.Sz, __int_as_float(r.kz));
              to_ready = true;
            }
          }
        } else if (chain < 0 && !(exhausted && wnext == wend)) {
          to_pending = true;  // this batch ran dry mid-fetch: ask again
        }
        if (to_ready || to_pending) {
          pw.A[ms] = make_float4(mo.x, mo.y, mo.z, __uint_as_float(seed));
          pw.B[ms] = make_float4(d.x, d.y, d.z, __uint_as_float(flags));
          pw.C[ms] = make_float4(atten.x, atten.y, atten.z, __uint_as_float((uint32_t)mid | (tries << 16)));
          pw.D[ms] = make_float4(color.x, color.y, color.z, brdf_w);
          pw.E[ms] = make_float4(N.x, N.y, N.z, __uint_as_float(kind));
          pw.F[ms] = f4;
          pw.H[ms] = h4;
          pw.G[ms] = make_float4(__int_as_float(chain), __uint_as_float(pixel), __uint_as_float(done), 0.0f);
        }
      }
      {
        const unsigned rm = __ballot_sync(FULL, to_ready), pm = __ballot_sync(FULL, to_pending);
        if (to_ready) pw.rq[(rq_head + rq_cnt + __popc(rm & lanemask_lt())) % POOL_N] = (unsigned char)ms;
        if (to_pending) pw.pq[(pq_head + pq_cnt + __popc(pm & lanemask_lt())) % POOL_N] = (unsigned char)ms;
        rq_cnt += __popc(rm); pq_cnt += __popc(pm);
      }
      __syncwarp();
      continue;
    }
    if (fly == 0u) {
      if (rq_cnt == 0u && pq_cnt == 0u) break;  // every slot is dead: the tile is done for this warp
      continue;
    }
    // ---- (c) one traversal quantum
    bool finished = false;
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
      if (WIDE) {  // end of a phase, by selects (the branchy form below cost ~3 % of the issue slots at 16 lanes)
        WideState& ws = *reinterpret_cast<WideState*>(&st);
        const bool done = occluded | (!ws.has_nodes() & !ws.has_tris());
        // emitters done: now the other triangles in front
        const bool to1 = done & (phase == 0) & !(any_hit & (best_prim >= 0)) & (sc.root_other >= 0);
        const bool fin = done & !to1;
        phase    = to1 ? 1 : phase;
        stack.sp = to1 ? 0 : stack.sp;
        ws.ng.x = to1 ? (uint32_t)sc.root_other : ws.ng.x; ws.ng.y = to1 ? 0x80000000u : ws.ng.y;
        ws.tg.x = to1 ? 0u : ws.tg.x;                      ws.tg.y = to1 ? 0u : ws.tg.y;
        finished  = fin;
        in_flight = !fin;
        if (fin) pw.F[slot] = make_float4(best_t, best_u, best_v, __int_as_float(occluded ? -2 : best_prim));
      } else if (occluded || (!st.has_nodes() && !st.has_tris())) {
        if (phase == 0 && !(any_hit && best_prim >= 0) && sc.root_other >= 0) {  // emitters done: now the other triangles in front
          phase = 1;
          stack.clear();
          st.begin(sc.root_other);
        } else {
          finished = true;
          in_flight = false;
          pw.F[slot] = make_float4(best_t, best_u, best_v, __int_as_float(occluded ? -2 : best_prim));
        }
      }
    }
    // the slots whose ray has just finished join the pending queue
    const unsigned fm = __ballot_sync(FULL, finished);
    if (fm) {
      if (finished) pw.pq[(pq_head + pq_cnt + __popc(fm & lanemask_lt())) % POOL_N] = (unsigned char)slot;
      pq_cnt += __popc(fm);
      __syncwarp();
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
