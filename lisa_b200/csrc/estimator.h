// lisa_b200/csrc/estimator.h — host-visible declarations of the render kernels (estimator.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "scene.cuh"

namespace lisa {

// Per-chain state, structure of arrays of 16-byte words.  A "chain" is one (pixel, subframe) sample
// sequence: the reference's per-thread loop `for i < samples_per_launch` (shader.cu:147-155) with its
// strictly sequential LCG stream seeded by tea<16>(pixel, subframe) (shader.cu:141).
struct DState {
  float4* o;    // ray origin .xyz (hit point P after the extend stage)
  float4* d;    // ray direction .xyz (not unit, Q5)
  float4* a;    // attenuation .xyz | LCG state in .w
  float4* c;    // radiance of the sample in flight .xyz | flags in .w (see F_* in estimator.cuh)
  float4* n;    // shading normal .xyz | material id in .w   (extend -> shadow stage hand-off)
  float4* sum;  // sum of finished samples .xyz | number of finished samples in .w
  int*    shadow_q;  // job queue: chain ids with an opaque hit whose next light-sampling tries are still to be drawn
  int*    cand_q;    // candidate queue: chain ids whose current try must be traced
  // ring of 3 per-iteration blocks of 16 counters: queue lengths and fetch cursors of every pass (estimator.cuh R_*)
  unsigned int* ring;
  // cumulative: [0] radiance rays [1] shadow rays [2] samples [3] null directions [4] finished chains
  // [5] BVH nodes visited [6] triangles tested [7] shadow jobs (opaque hits queued) [8] shadow tries resolved without traversal
  unsigned long long* stats;
};

struct Tile {
  uint32_t pix0, npix;      // pixel range of the tile (pixel id = y*W + x)
  uint32_t f0, nf;          // subframe range
  uint32_t spp;             // samples per chain
  uint32_t bounces;
  uint32_t n_chains;        // npix * nf
};

struct LaunchCfg {
  int sm_count;
  int extend_block, shadow_block;
  int shadow_blocks_per_sm;
  int tries_blocks_per_sm;
  int defer_retries;  // 1: jobs that need more tries after a failed candidate wait for the next iteration's k_tries
  int shadow_passes;  // k_tries/k_rays rounds per iteration (<= LISA_SHADOW_PASSES); the last one finishes leftovers inline
  int extend_blocks_per_sm;
  int idle_thresh;       // k_extend: lanes that must be idle before the warp runs its management section
  int idle_thresh_rays;  // same for k_rays
  int path_blocks_per_sm;  // k_path: resident CTAs per SM
  int path_wait_thresh;    // k_path: lanes that must be waiting before the warp runs its management section
  int pool_blocks_per_sm;  // k_pool: resident CTAs per SM
  int pool_dry_thresh;     // k_pool: pending slots that trigger a management section once the ready queue is empty
  int pool_deep;           // k_pool: 1 = the deep flavour (48 chains per warp, 12 stack entries in shared memory; sched_pool.cuh)
  int pool_blocks_per_sm_deep, pool_dry_thresh_deep;
};

void launch_init_chains(const DState& s, const DCamera& cam, const Tile& t, cudaStream_t st);
void launch_extend(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, uint32_t iter, const LaunchCfg& cfg,
                   cudaStream_t st);
// all light-sampling passes of one iteration; returns the number of kernels launched
int  launch_shadow(const DScene& sc, const DState& s, const Tile& t, uint32_t iter, const LaunchCfg& cfg, cudaStream_t st);
#define LISA_SHADOW_PASSES 4
void launch_finalize(const DState& s, const DCamera& cam, const Tile& t, float4* accum, cudaStream_t st);
void launch_resolve(const float4* accum, uint32_t npix, float4* mean_out, uint32_t* rgba8_out, cudaStream_t st);
void launch_accum_add(float4* dst, const float4* src, uint32_t npix, int sm_count, cudaStream_t st);
int  configure_kernels(char* err, size_t errlen);
int  shadow_occupancy(bool wide, int block);  // resident CTAs of k_rays per SM
int  extend_occupancy(bool wide, int block);
int  tries_occupancy(int block);               // resident CTAs of k_tries per SM
int  path_occupancy(bool wide, int block);     // resident CTAs of k_path per SM
int  pool_occupancy(bool wide, bool deep);     // resident CTAs of k_pool per SM
int  pool_chains_per_cta(bool deep);           // chains a k_pool CTA keeps in shared memory
void launch_pool(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, const LaunchCfg& cfg, cudaStream_t st);
// the whole tile in one persistent launch (chains fetched from a cursor in s.ring[0]; only s.sum, s.ring, s.stats are used)
void launch_path(const DScene& sc, const DState& s, const DCamera& cam, const Tile& t, const LaunchCfg& cfg, cudaStream_t st);

// Copies the device's "a traversal stack was full" flag (traverse.cuh: g_trav_overflow) to pinned host memory on `st` and
// clears it.  Non-zero after the stream is synchronised means some ray skipped a subtree: the caller must fail.
cudaError_t fetch_trav_overflow(unsigned int* h_pinned, cudaStream_t st);
int         traversal_stack_entries();  // capacity of a ray's traversal stack (LISA_STACK_TOTAL)

// diagnostics
void launch_trace_closest(const DScene& sc, const float* d_org, const float* d_dir, uint32_t n, float tmin, float tmax,
                          int* d_prim, float* d_t, cudaStream_t st);
void launch_trace_shadow(const DScene& sc, const float* d_org, const float* d_dir, uint32_t n, float tmin, float tmax,
                         int* d_outcome, int* d_light, cudaStream_t st);
void launch_primary_rays(const DCamera& cam, uint32_t subframe, float* d_dirs, uint32_t* d_seeds, cudaStream_t st);
int  launch_kat(int what, uint32_t n, const float* in_f, const uint32_t* in_u, float* out_f, uint32_t* out_u, cudaStream_t st);

}  // namespace lisa
