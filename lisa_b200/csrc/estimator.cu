// lisa_b200/csrc/estimator.cu — the translation unit of the render kernels (sm_100a): the three schedules of the
// estimator (sched_pool.cuh: a warp owns 64 chains in shared memory; sched_path.cuh: a lane owns a chain in registers;
// sched_wavefront.cuh: three kernels per bounce over chain state in HBM), the accumulate / resolve kernels, the
// diagnostic queries, and every launcher declared in estimator.h.
#include "sched_wavefront.cuh"
#include "sched_path.cuh"
#include "sched_pool.cuh"

namespace lisa {

// ------------------------------------------------------------------------------------------------
// Chain sums -> accumulators.  accum.xyz += the SUM of the samples of every subframe of the tile, accum.w += their number,
// in subframe order (fixed order => deterministic sums).  The image is accum.xyz / accum.w: the mean over all samples,
// which is what the reference's running mean of equally sized subframes is (shader.cu:158-164) — and stays the plain
// sample mean when subframes differ in size (N samples split over G GPUs with G not dividing N).
__global__ void k_finalize(DState s, Tile t, float4* accum) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.npix) return;
  float4 acc = accum[t.pix0 + k];
  for (uint32_t f = 0; f < t.nf; f++) {
    const float4 sm = s.sum[f * t.npix + k];
    acc.x = __fadd_rn(acc.x, sm.x); acc.y = __fadd_rn(acc.y, sm.y); acc.z = __fadd_rn(acc.z, sm.z); acc.w += (float)t.spp;
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
  auto path_attr = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); };
  auto pool_attr = [&](const void* f, size_t bytes) {
    if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  };
  path_attr((const void*)k_path<true, true>); path_attr((const void*)k_path<true, false>);
  path_attr((const void*)k_path<false, true>); path_attr((const void*)k_path<false, false>);
  pool_attr((const void*)k_pool<true, true, false>, PoolShape<false>::smem_bytes); pool_attr((const void*)k_pool<true, false, false>, PoolShape<false>::smem_bytes);
  pool_attr((const void*)k_pool<false, true, false>, PoolShape<false>::smem_bytes); pool_attr((const void*)k_pool<false, false, false>, PoolShape<false>::smem_bytes);
  pool_attr((const void*)k_pool<true, true, true>, PoolShape<true>::smem_bytes); pool_attr((const void*)k_pool<true, false, true>, PoolShape<true>::smem_bytes);
  if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
  return 0;
}

// the deep flavour exists for the 8-wide BVH only (the binary BVH is an ablation path)
static inline bool   pool_deep(const DScene& sc, const LaunchCfg& cfg) { return cfg.pool_deep && sc.wide; }
int pool_chains_per_cta(bool deep) { return (deep ? PoolShape<true>::SLOTS : PoolShape<false>::SLOTS) * 4; }
int pool_occupancy(bool wide, bool deep) {
  int n = 0;
  cudaError_t e = (wide && deep) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pool<true, true, true>, 128, PoolShape<true>::smem_bytes)
                  : wide         ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pool<true, true, false>, 128, PoolShape<false>::smem_bytes)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_pool<false, true, false>, 128, PoolShape<false>::smem_bytes);
  return (e == cudaSuccess && n > 0) ? n : 1;
}
void launch_pool(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, const LaunchCfg& cfg, cudaStream_t st) {
  const bool deep = pool_deep(sc, cfg);
  unsigned grid = (unsigned)(cfg.sm_count * (deep ? cfg.pool_blocks_per_sm_deep : cfg.pool_blocks_per_sm));
  grid = min(grid, max(1u, cdiv(t.n_chains, (unsigned)pool_chains_per_cta(deep))));
  cudaMemsetAsync(s.ring, 0, sizeof(unsigned int), st);  // chain fetch cursor
  const bool flat = sc.emit_flat >= 0;  // each kernel is instantiated with and without the flat-bounds filter of the shadow tries
  if (deep) {
    auto k = flat ? k_pool<true, true, true> : k_pool<true, false, true>;
    k<<<grid, 128, PoolShape<true>::smem_bytes, st>>>(sc, s, cam, t, (uint32_t)cfg.pool_dry_thresh_deep);
  } else {
    auto k = sc.wide ? (flat ? k_pool<true, true, false> : k_pool<true, false, false>) : (flat ? k_pool<false, true, false> : k_pool<false, false, false>);
    k<<<grid, 128, PoolShape<false>::smem_bytes, st>>>(sc, s, cam, t, (uint32_t)cfg.pool_dry_thresh);
  }
}

int path_occupancy(bool wide, int block) {
  int n = 0;
  cudaError_t e = wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_path<true, true>, block, stack_smem(block))
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_path<false, true>, block, stack_smem(block));
  return (e == cudaSuccess && n > 0) ? n : 2;
}
void launch_path(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, const LaunchCfg& cfg, cudaStream_t st) {
  const int b = 128;
  unsigned grid = (unsigned)(cfg.sm_count * cfg.path_blocks_per_sm);
  grid = min(grid, max(1u, cdiv(t.n_chains, 32u * (b / 32))));
  cudaMemsetAsync(s.ring, 0, sizeof(unsigned int), st);  // chain fetch cursor
  const bool flat = sc.emit_flat >= 0;
  auto k = sc.wide ? (flat ? k_path<true, true> : k_path<true, false>) : (flat ? k_path<false, true> : k_path<false, false>);
  k<<<grid, b, stack_smem(b), st>>>(sc, s, cam, t, (uint32_t)cfg.path_wait_thresh);
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
// traversal-stack overflow flag of the current device (traverse.cuh): copied into *h_pinned on `st`, then cleared
cudaError_t fetch_trav_overflow(unsigned int* h_pinned, cudaStream_t st) {
  cudaError_t e = cudaMemcpyFromSymbolAsync(h_pinned, g_trav_overflow, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return e;
  static const unsigned int zero = 0u;
  return cudaMemcpyToSymbolAsync(g_trav_overflow, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, st);
}
int traversal_stack_entries() { return LISA_STACK_TOTAL; }
int launch_kat(int what, uint32_t n, const float* in_f, const uint32_t* in_u, float* out_f, uint32_t* out_u, cudaStream_t st) {
  if (what < 0 || what > 9) return -1;
  if (n) k_kat<<<cdiv(n, 128), 128, 0, st>>>(what, n, in_f, in_u, out_f, out_u);
  return 0;
}

}  // namespace lisa
