// lisa_b200/csrc/sched_cta.cuh — k_cta: k_pool with the chain slots and the queues shared by the whole CTA, and the
// finished rays queued BY KIND (included by estimator.cu; LISA_PIPELINE=cta).
//
// k_pool's management section runs on the 32 slots its warp has pending, whatever they are: ~20 closest-hit results and
// ~12 shadow-ray results on the Cornell box, so the material dispatch runs at 11 lanes, the retirement of a shadow ray at 4,
// the end of a sample at 9 (profiles/r02_k_pool_regions.txt: a quarter of the warp instructions at ~10 lanes).  Here the
// CTA's 255 slots are one pool and there are three CTA-wide rings in shared memory:
//   ready    slots with a ray to trace            (any lane of any warp without a ray takes the next one)
//   hits     slots whose RADIANCE ray has finished (closest hit or miss; also slots waiting for a chain)
//   shadows  slots whose SHADOW ray has finished
// and a management section takes 32 slots of ONE kind — whichever ring is fuller — so that each of its branches runs at
// two to three times the lanes.  The rings are multi-producer / multi-consumer: a warp reserves positions with one atomic
// on the ring's tail (push) or a compare-and-swap on its head (pop); an entry is valid once it differs from EMPTY, the
// consumer puts EMPTY back.  A slot is in at most one ring, so a ring of 256 positions for 255 slots never overflows and a position is
// consumed before it can be written again.  __threadfence_block() orders a slot's data before its index (push) and the
// index before the data (pop).  Per-chain arithmetic and its order are k_pool's: the accumulators are bit-identical.
#pragma once
#include "sched_pool.cuh"

namespace lisa {

#define CTA_SLOTS 255   // slot ids 0..254: a ring entry is one byte and 255 means EMPTY
#define CTA_RING 256    // ring positions (a power of two; more than there are slots, so a ring never overflows)
#define CTA_EMPTY 0xffu

struct CtaShared {
  float4         A[CTA_RING], B[CTA_RING], C[CTA_RING], D[CTA_RING], E[CTA_RING], F[CTA_RING], H[CTA_RING], G[CTA_RING];
  unsigned       h_head, h_tail, s_head, s_tail;  // one LDS.128
  unsigned       r_head, r_tail, dead, pad;
  unsigned char  rq[CTA_RING], hq[CTA_RING], sq[CTA_RING];
};

// Ordering between a slot's data and its index in a ring.  LISA_CTA_FENCE=1: __threadfence_block() (the PTX memory model's
// answer).  Default: a compiler barrier only — shared-memory accesses of one thread are issued to the SM's one shared-memory
// pipe in program order and performed there in order, which is what the protocol needs; the fence also waits for the
// warp's outstanding GLOBAL loads (the node it has just fetched), and that serialises the traversal.
#ifndef LISA_CTA_FENCE
#define LISA_CTA_FENCE 0
#endif
__device__ __forceinline__ void cta_order() {
#if LISA_CTA_FENCE
  __threadfence_block();
#else
  asm volatile("" ::: "memory");
#endif
}

__device__ __forceinline__ uint4 lds_volatile_v4(const void* p) {
  uint4          v;
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

// all lanes call; lanes with p append their slot
__device__ __forceinline__ void cta_push(unsigned char* e, unsigned* tail, bool p, int slot, unsigned lane) {
  const unsigned m = __ballot_sync(FULL, p);
  if (!m) return;
  cta_order();  // the slot's data before its index
  const int leader = __ffs(m) - 1;
  unsigned  base = 0;
  if ((int)lane == leader) base = atomicAdd(tail, (unsigned)__popc(m));
  base = __shfl_sync(FULL, base, leader);
  if (p) reinterpret_cast<volatile unsigned char*>(e)[(base + __popc(m & lanemask_lt())) & (CTA_RING - 1)] = (unsigned char)slot;
}

// all lanes call; takes up to `want` entries, lane `rank` (< the count returned) gets the rank-th of them
__device__ __forceinline__ unsigned cta_pop(unsigned char* e, unsigned* head, unsigned* tail, unsigned want, unsigned lane, unsigned rank,
                                            int& slot) {
  unsigned h = 0, n = 0;
  if (lane == 0) {
    while (true) {
      h = *reinterpret_cast<volatile unsigned*>(head);
      n = min(*reinterpret_cast<volatile unsigned*>(tail) - h, want);
      if (n == 0u || atomicCAS(head, h, h + n) == h) break;
    }
  }
  h = __shfl_sync(FULL, h, 0);
  n = __shfl_sync(FULL, n, 0);
  if (rank < n) {
    volatile unsigned char* p = reinterpret_cast<volatile unsigned char*>(e) + ((h + rank) & (CTA_RING - 1));
    unsigned v;
#ifdef LISA_CTA_WATCHDOG
    unsigned spins = 0;
    do { v = *p; if (++spins > LISA_CTA_WATCHDOG) { printf("k_cta watchdog: pop spin block %d pos %u\n", blockIdx.x, (h + rank) & (CTA_RING - 1)); v = 0; break; } } while (v == CTA_EMPTY);
#else
    do { v = *p; } while (v == CTA_EMPTY);  // reserved by a producer that has not stored it yet (a few cycles)
#endif
    *p = (unsigned char)CTA_EMPTY;
    slot = (int)v;
  }
  if (n) cta_order();  // the index before the slot's data
  return n;
}

template <bool WIDE, bool FLAT>
__global__ void __launch_bounds__(128, LISA_POOL_MIN_BLOCKS) k_cta(DScene sc, DState s, DCamera cam, Tile t, uint32_t dry_thresh) {
  extern __shared__ uint2 smem_stack[];  // [LISA_STACK_SM][128] traversal stacks, then CtaShared
  __shared__ uint32_t lcg_a[32], lcg_c[32];
  __shared__ float4   jobbuf[4][96];
  Stack          stack(smem_stack);
  CtaShared&     cs   = *reinterpret_cast<CtaShared*>(smem_stack + LISA_STACK_SM * 128);
  const unsigned lane = lane_id();
  fill_lcg_tables(lcg_a, lcg_c);
  // every slot starts in the hits ring without a chain: the first management sections fetch chains for them
  for (unsigned k = threadIdx.x; k < CTA_RING; k += 128) {
    cs.hq[k] = (unsigned char)k;  // position 255 holds EMPTY (= 255): there are 255 slots
    cs.rq[k] = (unsigned char)CTA_EMPTY;
    cs.sq[k] = (unsigned char)CTA_EMPTY;
    cs.G[k]  = make_float4(__int_as_float(-1), 0.0f, 0.0f, 0.0f);
    cs.F[k]  = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
    cs.E[k]  = make_float4(0.0f, 1.0f, 0.0f, __uint_as_float(0u));
  }
  if (threadIdx.x == 0) {
    cs.h_head = 0; cs.h_tail = CTA_SLOTS; cs.s_head = cs.s_tail = 0; cs.r_head = cs.r_tail = 0; cs.dead = 0; cs.pad = 0;
  }
  __syncthreads();
  unsigned int* cursor = &s.ring[0];  // zeroed by the host before the launch

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
  bool     exhausted = false;  // this warp has seen the chain cursor pass the end of the tile
  unsigned idle_spins = 0;
#ifdef LISA_CTA_WATCHDOG
  unsigned wd_iter = 0;
#endif
#ifdef LISA_CTA_STATS
  unsigned long long d_iter = 0, d_man = 0, d_np = 0, d_idle = 0, d_fail = 0, d_fly = 0, d_manH = 0, d_takes = 0, d_taken = 0;
#endif
  uint32_t n_rad = 0, n_null = 0, n_samp = 0, n_done = 0, nn = 0, nt = 0;
  EventCounters ec = {0, 0, 0};

  while (true) {
    // ---- (a) lanes without a ray take ready slots
#ifdef LISA_CTA_STATS
    d_iter++;
#endif
    const unsigned freemask = __ballot_sync(FULL, !in_flight);
    unsigned r_left = 1;  // entries seen in the ready ring after this warp took its share (0: it ran dry)
    if (freemask) {
      const unsigned rank = in_flight ? 0xffffu : (unsigned)__popc(freemask & lanemask_lt());
      int            ns = -1;
      const unsigned n = cta_pop(cs.rq, &cs.r_head, &cs.r_tail, (unsigned)__popc(freemask), lane, rank, ns);
      if (rank < n) {
        slot = ns;
        const float4   a4 = cs.A[slot], f4 = cs.F[slot], h4 = cs.H[slot];
        const uint32_t kind = __float_as_uint(cs.E[slot].w);
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
      r_left = n < (unsigned)__popc(freemask) ? 0u : 1u;
#ifdef LISA_CTA_STATS
      d_takes++; d_taken += n;
#endif
    }
    const unsigned fly = __ballot_sync(FULL, in_flight);
    // ---- (b) management section: 32 finished slots of one kind, or fewer when the ready ring has run dry.
    // Lane 0 looks at the rings and broadcasts: every lane must take the same decision.
    unsigned info = 0;
    if (lane == 0) {
      const uint4 qc = lds_volatile_v4(&cs.h_head);
      info = min(qc.y - qc.x, 511u) | (min(qc.w - qc.z, 511u) << 9);
      if (fly == 0u && *reinterpret_cast<volatile unsigned*>(&cs.dead) == CTA_SLOTS) info |= 0x80000000u;
    }
    info = __shfl_sync(FULL, info, 0);
    const unsigned nH = info & 511u, nS = (info >> 9) & 511u, nmax = max(nH, nS);
    bool manage = nmax >= 32u || (r_left == 0u && nmax >= dry_thresh);
#ifdef LISA_CTA_WATCHDOG
    if (++wd_iter > LISA_CTA_WATCHDOG) {
      if (lane == 0) printf("k_cta watchdog: block %d warp %d fly %08x nH %u nS %u rq %u dead %u\n", blockIdx.x, threadIdx.x >> 5, fly, nH, nS,
                            *reinterpret_cast<volatile unsigned*>(&cs.r_tail) - *reinterpret_cast<volatile unsigned*>(&cs.r_head),
                            *reinterpret_cast<volatile unsigned*>(&cs.dead));
      break;
    }
#endif
    if (!manage && fly == 0u) {
      // nothing to trace here.  The other warps' rays will fill the rings; when nothing moves for a while (they are idle
      // too) the leftovers are taken as they are
      if (info & 0x80000000u) break;  // every slot is dead: the tile is done
      if (nmax > 0u && ++idle_spins >= 8u) manage = true;
      else {
#ifdef LISA_CTA_STATS
        d_idle++;
#endif
        __nanosleep(100); continue; }
    }
    if (manage) {
      idle_spins = 0;
      const bool take_s = nS > nH;
      int        ms = -1;
      const unsigned np = take_s ? cta_pop(cs.sq, &cs.s_head, &cs.s_tail, 32u, lane, lane, ms)
                                 : cta_pop(cs.hq, &cs.h_head, &cs.h_tail, 32u, lane, lane, ms);
#ifdef LISA_CTA_STATS
      if (np == 0u) d_fail++; else { d_man++; d_np += np; d_manH += take_s ? 0 : 1; }
#endif
      if (np == 0u) continue;  // another warp was faster
      const bool active = lane < np;
      // slot -> registers
      int      chain = -1;
      uint32_t pixel = 0, done = 0, seed = 0, flags = 0, tries = 0, kind = 0;
      int      mid = 0;
      float3   mo = f3(0, 0, 0), d = f3(0, 0, 1), atten = f3(1, 1, 1), color = f3(0, 0, 0), N = f3(0, 1, 0);
      float    brdf_w = 0.0f, r_t = LISA_TMAX, r_u = 0.0f, r_v = 0.0f;
      int      r_prim = -1;
      if (active) {
        const float4 a4 = cs.A[ms], b4 = cs.B[ms], c4 = cs.C[ms], d4 = cs.D[ms], e4 = cs.E[ms], f4 = cs.F[ms], g4 = cs.G[ms];
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
      // ---- (6) fetch chains: consecutive ids for the lanes of this section that need one
      const bool     need     = active && chain < 0;
      const unsigned needmask = __ballot_sync(FULL, need);
      bool           die      = false;
      if (needmask) {
        unsigned base = t.n_chains;
        if (!exhausted) {
          if (lane == 0) base = atomicAdd(cursor, (unsigned)__popc(needmask));
          base = __shfl_sync(FULL, base, 0);
          if (base + (unsigned)__popc(needmask) >= t.n_chains) exhausted = true;
        }
        if (need) {
          const unsigned id = base + (unsigned)__popc(needmask & lanemask_lt());
          if (id < t.n_chains) {
            chain = (int)id;
            const uint32_t p = t.pix0 + (uint32_t)chain % t.npix;
            seed  = chain_seed(cam, p, t.f0 + (uint32_t)chain / t.npix);
            pixel = (p % cam.width) | ((p / cam.width) << 16);
            done  = 0;
            fresh = true;
          } else die = true;  // no chain left: the slot is retired
        }
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
      // ---- (8) set the ray up and hand the slot over: ready ring, or straight back to a finished ring when there is
      // nothing to trace (null direction, Q7; no triangle in reach)
      bool to_ready = false, to_pending = false;
      if (active && !die) {
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
            if ((kind & 2u) && sc.root_other < 0) to_pending = true;  // F keeps the MISS record
            else {
              const StepRay r = step_ray(rd);
              f4 = make_float4(r.idir.x, r.idir.y, r.idir.z, __uint_as_float(r.oct_inv4));
              h4 = make_float4(r.Sx, r.Sy, r.Sz, __int_as_float(r.kz));
              to_ready = true;
            }
          }
        }
        if (to_ready || to_pending) {
          cs.A[ms] = make_float4(mo.x, mo.y, mo.z, __uint_as_float(seed));
          cs.B[ms] = make_float4(d.x, d.y, d.z, __uint_as_float(flags));
          cs.C[ms] = make_float4(atten.x, atten.y, atten.z, __uint_as_float((uint32_t)mid | (tries << 16)));
          cs.D[ms] = make_float4(color.x, color.y, color.z, brdf_w);
          cs.E[ms] = make_float4(N.x, N.y, N.z, __uint_as_float(kind));
          cs.F[ms] = f4;
          cs.H[ms] = h4;
          cs.G[ms] = make_float4(__int_as_float(chain), __uint_as_float(pixel), __uint_as_float(done), 0.0f);
        }
      }
      cta_push(cs.rq, &cs.r_tail, to_ready, ms, lane);
      cta_push(cs.hq, &cs.h_tail, to_pending && !(kind & 1u), ms, lane);
      cta_push(cs.sq, &cs.s_tail, to_pending && (kind & 1u), ms, lane);
      {
        const unsigned dm = __ballot_sync(FULL, die);
        if (dm && lane == 0) atomicAdd(&cs.dead, (unsigned)__popc(dm));
      }
      continue;
    }
#ifdef LISA_CTA_STATS
    d_fly += __popc(fly);
#endif
    // ---- (c) one traversal quantum
    bool finished = false;
    if (in_flight) {
      if (st.has_nodes() && !st.has_tris()) {
        LISA_COUNT(nn);
        if (WIDE) wide_node_step(sc.bvh, o, ray, LISA_TMIN, best_t, *reinterpret_cast<WideState*>(&st), stack);
        else bin_node_step(sc.bvh, o, ray, LISA_TMIN, best_t, *reinterpret_cast<BinState*>(&st), stack);
      }
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
      if (WIDE) {
        WideState& ws = *reinterpret_cast<WideState*>(&st);
        const bool done = occluded | (!ws.has_nodes() & !ws.has_tris());
        const bool to1 = done & (phase == 0) & !(any_hit & (best_prim >= 0)) & (sc.root_other >= 0);
        const bool fin = done & !to1;
        phase    = to1 ? 1 : phase;
        stack.sp = to1 ? 0 : stack.sp;
        ws.ng.x = to1 ? (uint32_t)sc.root_other : ws.ng.x; ws.ng.y = to1 ? 0x80000000u : ws.ng.y;
        ws.tg.x = to1 ? 0u : ws.tg.x;                      ws.tg.y = to1 ? 0u : ws.tg.y;
        finished  = fin;
        in_flight = !fin;
        if (fin) cs.F[slot] = make_float4(best_t, best_u, best_v, __int_as_float(occluded ? -2 : best_prim));
      } else if (occluded || (!st.has_nodes() && !st.has_tris())) {
        if (phase == 0 && !(any_hit && best_prim >= 0) && sc.root_other >= 0) {
          phase = 1;
          stack.clear();
          st.begin(sc.root_other);
        } else {
          finished = true;
          in_flight = false;
          cs.F[slot] = make_float4(best_t, best_u, best_v, __int_as_float(occluded ? -2 : best_prim));
        }
      }
    }
    // the slots whose ray has just finished join the ring of their kind
    if (__any_sync(FULL, finished)) {
      cta_push(cs.hq, &cs.h_tail, finished && !shadow_ray, slot, lane);
      cta_push(cs.sq, &cs.s_tail, finished && shadow_ray, slot, lane);
    }
  }
#ifdef LISA_CTA_STATS
  if (lane == 0 && blockIdx.x == 7 && t.n_chains > 1000000u)
    printf("k_cta stats warp %d: iter %llu, traversal iters %llu at %.1f lanes, manage %llu (hits %llu) avg np %.1f, failed pops %llu, idle %llu, ready pops %llu avg %.2f\n", threadIdx.x >> 5,
           d_iter, d_iter - d_man - d_fail - d_idle, (double)d_fly / (double)(d_iter - d_man - d_fail - d_idle), d_man, d_manH, (double)d_np / (double)d_man, d_fail, d_idle, d_takes, (double)d_taken / (double)d_takes);
#endif
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
