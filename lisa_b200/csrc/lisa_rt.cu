// lisa_b200/csrc/lisa_rt.cu — implementation of the C ABI in include/lisa_rt.h.
//
// Host-side driver of the render path: scene upload, device BVH build (bvh_build.cu), camera frame,
// the render drivers (estimator.cu) and read-back.  It stands where the reference has
// OptixWrapper (src/LiSA/src/optix_wrapper.cc) and launchSubframe/render (src/LiSA/src/render.cc).
// There is no CPU path: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lisa_rt.h"
#include "build.h"
#include "devmem.h"
#include "sort_scan.h"
#include "scene.cuh"
#include "estimator.h"
#include "internal.h"

#include <nvtx3/nvToolsExt.h>

#ifndef LISA_POOL_DEEP_SAH
#define LISA_POOL_DEEP_SAH 40.0f
#endif
#ifndef LISA_LEAF_SAH_DEFAULT
#define LISA_LEAF_SAH_DEFAULT 0.5f
#endif

using namespace lisa;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) return fail(LISA_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_));    \
  } while (0)

struct lisa_ctx {
  int          device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev0 = nullptr, ev1 = nullptr;
  DScene       scene{};
  DCamera      cam{};
  BuildOutput  bvh{};
  DMaterial*   d_mats = nullptr;
  float4*      d_accum = nullptr;   // W*H float4: sum of all samples | number of samples
  float4*      d_mean = nullptr;    // scratch for read-back
  uint32_t*    d_rgba8 = nullptr;
  DState       state{};
  size_t       state_chains = 0;    // capacity of the state arrays, in chains
  uint32_t     max_chains = 0;
  unsigned long long* h_stats = nullptr;  // pinned mirror of state.stats
  LaunchCfg    cfg{};
  lisa_stats   stats{};
  uint32_t     width = 0, height = 0, num_samples = 0, num_bounces = 0;
  std::string  output_image;
  uint32_t     emit_hash = 0;       // of the materials' emitter flags (what a serialised BVH is valid for)
  bool         profile_stages = false;
  float        pool_min_chains_factor = 1.2f;  // k_pool from this many times its slots (run_tile; LISA_POOL_MIN_FACTOR)
  bool         pool_flavour_auto = true;       // k_pool's flavour follows the measured node visits per ray (lisa_render_subframes)
  int          pipeline = 3;        // 3: per tile, k_pool when the tile has enough chains to fill its slots 1.2 times, else k_path;
                                    // 2: k_pool, 1: k_path (one persistent launch per tile either way), 0: wavefront (three kernels per bounce)
  bool         state_full = false;  // the wavefront arrays are allocated (k_path needs only state.sum)
  std::vector<cudaEvent_t> ev_pool;  // stage profiling: 3 events per iteration (extend start, shadow start, shadow end)
  size_t       ev_used = 0;
  double       stage_ms[2] = {0, 0};
  uint64_t     stage_launches[2] = {0, 0};
};

static cudaEvent_t next_event(lisa_ctx* c) {
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev_pool.push_back(e);
  }
  return c->ev_pool[c->ev_used++];
}
// events were recorded as triples (extend start, shadow start, shadow end); call after a stream sync
static void drain_stage_events(lisa_ctx* c) {
  for (size_t i = 0; i + 2 < c->ev_used + 0 && i + 2 < c->ev_pool.size() + 0; i += 3) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, c->ev_pool[i], c->ev_pool[i + 1]);
    cudaEventElapsedTime(&b, c->ev_pool[i + 1], c->ev_pool[i + 2]);
    c->stage_ms[0] += a; c->stage_ms[1] += b;
    c->stage_launches[0]++; c->stage_launches[1]++;
  }
  c->ev_used = 0;
}

// the pinned slot behind h_stats is 256 bytes: the 16 counters, then the traversal-stack overflow flag
static unsigned int* overflow_slot(lisa_ctx* c) { return reinterpret_cast<unsigned int*>(c->h_stats + 16); }
static int stack_overflow_error() {
  return fail(LISA_ERR_STATE, "traversal stack overflow: the BVH is deeper than %d levels of 8-wide nodes (a chain-shaped "
              "hierarchy); some rays skipped a subtree, so the result is discarded", traversal_stack_entries());
}

extern "C" const char* lisa_last_error(void) { return g_err; }
extern "C" void        lisa_internal_set_last_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }
extern "C" int         lisa_version(void) { return LISA_RT_VERSION; }

// src/sutil/Camera.cpp:34-45 with up = (0,1,0) and aspect = width/height (optix_wrapper.cc:433-442)
static void camera_frame(const lisa_camera& c, uint32_t w, uint32_t h, DCamera* out) {
  auto  sub = [](const float* a, const float* b, float* r) { for (int i = 0; i < 3; i++) r[i] = a[i] - b[i]; };
  auto  crs = [](const float* a, const float* b, float* r) {
    r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
  };
  auto  len = [](const float* a) { return sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
  auto  nrm = [&](float* a) { float inv = 1.0f / len(a); for (int i = 0; i < 3; i++) a[i] *= inv; };
  float W[3], U[3], V[3];
  const float up[3] = {0.0f, 1.0f, 0.0f};
  sub(c.look_at, c.eye, W);
  const float wlen = len(W);
  crs(W, up, U); nrm(U);
  crs(U, W, V); nrm(V);
  const float vlen = wlen * tanf(0.5f * c.fov * (float)M_PI / 180.0f);
  for (int i = 0; i < 3; i++) V[i] *= vlen;
  const float ulen = vlen * ((float)w / (float)h);
  for (int i = 0; i < 3; i++) U[i] *= ulen;
  out->eye = make_float3(c.eye[0], c.eye[1], c.eye[2]);
  out->U = make_float3(U[0], U[1], U[2]);
  out->V = make_float3(V[0], V[1], V[2]);
  out->W = make_float3(W[0], W[1], W[2]);
  out->width = w; out->height = h;
}

static void free_state(lisa_ctx* c) {
  dev_free(c->state.o); dev_free(c->state.d);
  dev_free(c->state.a); dev_free(c->state.c);
  dev_free(c->state.n); dev_free(c->state.sum);
  dev_free(c->state.shadow_q);
  dev_free(c->state.cand_q);
  c->state.o = c->state.d = c->state.a = c->state.c = c->state.n = c->state.sum = nullptr;
  c->state.shadow_q = c->state.cand_q = nullptr;
  c->state_chains = 0;
  c->state_full = false;
}

static int ensure_state(lisa_ctx* c, size_t chains) {
  const bool full = c->pipeline == 0;
  if (chains <= c->state_chains && (c->state_full || !full)) return LISA_OK;
  chains = std::max(chains, c->state_chains);
  free_state(c);
  CU(dev_alloc((void**)&c->state.sum, sizeof(float4) * chains));
  if (full) {
    CU(dev_alloc((void**)&c->state.o, sizeof(float4) * chains));
    CU(dev_alloc((void**)&c->state.d, sizeof(float4) * chains));
    CU(dev_alloc((void**)&c->state.a, sizeof(float4) * chains));
    CU(dev_alloc((void**)&c->state.c, sizeof(float4) * chains));
    CU(dev_alloc((void**)&c->state.n, sizeof(float4) * chains));
    CU(dev_alloc((void**)&c->state.shadow_q, sizeof(int) * chains));
    CU(dev_alloc((void**)&c->state.cand_q, sizeof(int) * chains));
  }
  c->state_chains = chains;
  c->state_full = full;
  c->stats.state_bytes = full ? chains * (6 * sizeof(float4) + 2 * sizeof(int)) : chains * sizeof(float4);
  return LISA_OK;
}

extern "C" void lisa_destroy(lisa_ctx* c) {
  if (!c) return;
  const bool dbg = getenv("LISA_DEBUG_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  auto t0 = now();
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  auto t1 = now();
  free_state(c);
  {
    dev_free(c->d_accum);
    dev_free(c->d_mean);
    dev_free(c->d_rgba8);
  }
  auto t2 = now();
  dev_free(c->state.ring); dev_free(c->state.stats);
  dev_free(c->bvh.d_nodes); dev_free(c->bvh.d_tri_v); dev_free(c->bvh.d_tri_n); dev_free(c->bvh.d_final_to_orig);
  dev_free(c->d_mats);
  auto t3 = now();
  pinned_slot_free(c->h_stats);
  auto t4 = now();
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  auto t5 = now();
  if (dbg)
    fprintf(stderr, "lisa_destroy: sync %.2f  cache %.2f  cudaFree %.2f  freeHost %.2f  events+stream %.2f ms\n", ms(t0, t1), ms(t1, t2),
            ms(t2, t3), ms(t3, t4), ms(t4, t5));
  delete c;
}

// ---- serialised BVH -------------------------------------------------------------------
// File: BvhFileHeader, then the node array, tri_v, tri_n (3 float4 per triangle each, leaf order), final_to_orig.
struct BvhFileHeader {
  char     magic[8];  // "LISABVH1"
  uint32_t version, header_bytes;
  uint64_t num_tris, num_nodes;
  int32_t  nodes_other, nodes_emit, root_other, root_emit, num_emit_tris, wide;
  int32_t  num_materials;
  uint32_t emit_hash;  // FNV-1a over the materials' emitter flags: the emitter / non-emitter partition is baked into the file
  float    box_other[6], box_emit[6];
  uint64_t file_bytes;
  uint64_t num_input_tris;  // triangles of the soup the BVH was built from (< num_tris when triangles were split into references)
  float    sah_nodes_per_ray;  // BuildOutput::sah_nodes_per_ray (version 3): chooses k_pool's flavour
  uint32_t reserved;
};
static uint32_t emit_flags_hash(const lisa_scene_desc* sd) {
  uint32_t h = 2166136261u;
  for (int i = 0; i < sd->num_materials; i++) h = (h ^ (sd->materials[i].emit ? 1u : 0u)) * 16777619u;
  return h;
}
static uint64_t bvh_file_bytes(const BvhFileHeader& h) {
  return sizeof(BvhFileHeader) + h.num_nodes * (h.wide ? 80ull : 64ull) + h.num_tris * (48ull + 48ull + 4ull);
}

// reads a file written by lisa_save_bvh into device arrays (c->bvh), in place of the upload of the soup and the build
static int load_bvh_file(const lisa_scene_desc* sd, const char* path, lisa_ctx* c, int* wide_out) {
  FILE* f = fopen(path, "rb");
  if (!f) return fail(LISA_ERR_IO, "cannot open %s", path);
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{f};
  BvhFileHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "LISABVH1", 8) != 0 || h.version != 3 || h.header_bytes != sizeof(h))
    return fail(LISA_ERR_IO, "%s is not a serialised BVH of this library", path);
  if (fseek(f, 0, SEEK_END) != 0 || (uint64_t)ftell(f) != h.file_bytes || h.file_bytes != bvh_file_bytes(h))
    return fail(LISA_ERR_IO, "%s is truncated or damaged", path);
  if (sd->num_vertices && (uint64_t)(sd->num_vertices / 3) != h.num_input_tris)
    return fail(LISA_ERR_ARG, "%s holds %llu triangles, the scene has %d", path, (unsigned long long)h.num_input_tris, sd->num_vertices / 3);
  if (h.num_materials != sd->num_materials || h.emit_hash != emit_flags_hash(sd))
    return fail(LISA_ERR_ARG, "%s was built for %d materials with other emitter flags (the emitter partition is part of the BVH)", path, h.num_materials);
  fseek(f, (long)sizeof(h), SEEK_SET);
  BuildOutput& b = c->bvh;
  memset(&b, 0, sizeof(b));
  const size_t T = (size_t)h.num_tris, nb = (size_t)h.num_nodes * (h.wide ? 80 : 64);
  b.num_nodes = (int)h.num_nodes; b.nodes_other = h.nodes_other; b.nodes_emit = h.nodes_emit;
  b.root_other = h.root_other; b.root_emit = h.root_emit; b.num_emit_tris = h.num_emit_tris; b.node_bytes = nb;
  b.num_tris = (int)h.num_tris;
  b.num_input_tris = (int)h.num_input_tris;
  b.sah_nodes_per_ray = h.sah_nodes_per_ray;
  memcpy(b.box_other, h.box_other, sizeof(b.box_other));
  memcpy(b.box_emit, h.box_emit, sizeof(b.box_emit));
  *wide_out = h.wide;
  if (T == 0) return LISA_OK;
  CU(dev_alloc((void**)&b.d_nodes, std::max<size_t>(nb, 16)));
  CU(dev_alloc((void**)&b.d_tri_v, 48 * T));
  CU(dev_alloc((void**)&b.d_tri_n, 48 * T));
  CU(dev_alloc((void**)&b.d_final_to_orig, 4 * T));
  std::vector<char> buf(std::min<size_t>(64u << 20, std::max<size_t>(48 * T, nb)));
  auto pump = [&](void* dst, size_t bytes) -> int {  // file -> device in chunks (the copy of a chunk overlaps the read of the next)
    for (size_t off = 0; off < bytes; off += buf.size()) {
      const size_t n = std::min(buf.size(), bytes - off);
      if (fread(buf.data(), 1, n, f) != n) return fail(LISA_ERR_IO, "%s: short read", path);
      CU(cudaMemcpyAsync((char*)dst + off, buf.data(), n, cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));
    }
    return LISA_OK;
  };
  int rc;
  if ((rc = pump(b.d_nodes, nb)) || (rc = pump(b.d_tri_v, 48 * T)) || (rc = pump(b.d_tri_n, 48 * T)) || (rc = pump(b.d_final_to_orig, 4 * T))) return rc;
  return LISA_OK;
}

static int create_impl(const lisa_scene_desc* sd, const lisa_options* opt, lisa_ctx* c, const char* bvh_path = nullptr) {
  lisa_options o{};
  o.struct_size = sizeof(o);
  o.device = -1;
  if (opt) memcpy(&o, opt, std::min<size_t>(opt->struct_size ? opt->struct_size : sizeof(o), sizeof(o)));
  if (const char* e = getenv("LISA_BVH")) o.bvh_kind = !strcmp(e, "binary") ? LISA_BVH_BINARY : LISA_BVH_WIDE8;
  if (const char* e = getenv("LISA_SHADOW")) o.shadow_mode = !strcmp(e, "first") ? LISA_SHADOW_FIRST_FOUND : LISA_SHADOW_CLOSEST;
  if (const char* e = getenv("LISA_MAX_CHAINS")) o.max_chains = (uint32_t)strtoul(e, nullptr, 10);
  c->profile_stages = (o.flags & LISA_FLAG_PROFILE_STAGES) || (getenv("LISA_PROFILE_STAGES") && atoi(getenv("LISA_PROFILE_STAGES")));

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(LISA_ERR_CUDA, "no CUDA device (%s): lisa_rt has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (o.device >= 0) { CU(cudaSetDevice(o.device)); }
  CU(cudaGetDevice(&c->device));
  // per-device constants are queried once per process (cudaGetDeviceProperties alone can take ~100 ms)
  struct DevInfo { bool ok = false; int sms = 0, occ_rays[2] = {0, 0}, occ_ext[2] = {0, 0}, occ_path[2] = {0, 0}, occ_pool[2] = {0, 0}, occ_pool_deep = 0, occ_tries = 0; };
  static DevInfo dev_info[64];
  DevInfo& di = dev_info[c->device & 63];
  if (!di.ok) {
    CU(cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, c->device));
    if (configure_kernels(g_err, sizeof(g_err))) return LISA_ERR_CUDA;
  }
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));

  int T = sd->num_vertices / 3;
  c->width = sd->width; c->height = sd->height;
  c->num_samples = sd->num_samples; c->num_bounces = sd->num_bounces;
  if (sd->output_image) c->output_image = sd->output_image;
  camera_frame(sd->camera, sd->width, sd->height, &c->cam);

  std::vector<DMaterial>     mats((size_t)std::max(sd->num_materials, 1));
  std::vector<unsigned char> emit((size_t)std::max(sd->num_materials, 1), 0);
  int first_light = -1, single_light = 1;
  for (int i = 0; i < sd->num_materials; i++) {
    const lisa_material& m = sd->materials[i];
    mats[i].a = make_float4(m.diffuse_color[0], m.diffuse_color[1], m.diffuse_color[2], m.roughness);
    mats[i].b = make_float4(m.emission_color[0], m.emission_color[1], m.emission_color[2], m.n);
    mats[i].c = make_float4(m.alpha, m.emit ? 1.0f : 0.0f, 0.0f, 0.0f);
    emit[i]   = m.emit ? 1 : 0;
    if (m.emit) { if (first_light < 0) first_light = i; else single_light = 0; }
  }
  c->emit_hash = emit_flags_hash(sd);
  CU(dev_alloc((void**)&c->d_mats, sizeof(DMaterial) * mats.size()));
  CU(cudaMemcpyAsync(c->d_mats, mats.data(), sizeof(DMaterial) * mats.size(), cudaMemcpyHostToDevice, c->stream));
  int wide = o.bvh_kind == LISA_BVH_WIDE8 ? 1 : 0;
  if (bvh_path) {
    // ---- a BVH built earlier (lisa_save_bvh): no soup upload, no build
    const auto t_load = std::chrono::steady_clock::now();
    int rc = load_bvh_file(sd, bvh_path, c, &wide);
    if (rc) return rc;
    T = c->bvh.num_tris;
    CU(cudaStreamSynchronize(c->stream));
    c->stats.upload_ms = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_load).count();
    c->stats.bvh_build_ms = 0.0f;
  } else {
  // ---- upload (Q11: one material index per triangle; Q12: the caller zero-fills unused material fields)
  struct Ev3 {  // destroyed on every exit path
    cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
    ~Ev3() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
  } ev3;
  for (cudaEvent_t& x : ev3.e) CU(cudaEventCreate(&x));
  cudaEvent_t t0 = ev3.e[0], t1 = ev3.e[1], t2 = ev3.e[2];
  float *d_verts = nullptr, *d_normals = nullptr;
  int*   d_mat_idx = nullptr;
  unsigned char* d_emit = nullptr;
  CU(cudaEventRecord(t0, c->stream));
  const size_t vb = sizeof(float) * 9 * (size_t)std::max(T, 1);
  CU(dev_alloc((void**)&d_verts, vb)); CU(dev_alloc((void**)&d_normals, vb));
  CU(dev_alloc((void**)&d_mat_idx, sizeof(int) * (size_t)std::max(T, 1)));
  CU(dev_alloc((void**)&d_emit, emit.size()));
  if (T) {
    // large soups go through a ring of pinned buffers filled by worker threads (devmem.cu: upload_async)
    CU(upload_async(d_verts, sd->vertices, sizeof(float) * 9 * (size_t)T, c->stream));
    CU(upload_async(d_normals, sd->normals, sizeof(float) * 9 * (size_t)T, c->stream));
    CU(upload_async(d_mat_idx, sd->mat_indices, sizeof(int) * (size_t)T, c->stream));
  }
  CU(cudaMemcpyAsync(d_emit, emit.data(), emit.size(), cudaMemcpyHostToDevice, c->stream));
  CU(cudaEventRecord(t1, c->stream));

  // ---- BVH
  // PLOC search radius, measured on B200 (render Msamples/s at radius 2 / 3 / 4 / 5 / 8 / 16 / 24, one box per group of columns):
  //   Cornell 1569 / 1563 / 1570 / 1560 / 1564 | 1562 / 1566 / 1560     knot  1639 / 1688 / 1685 / 1702 / 1687 | 1698 / 1675 / 1676
  //   1M soup 136.0 / 134.0 / 135.9 / 134.4 / 130.8 | 130.6 / 126.1 / 124.2     10M soup 79.3 / 79.0 / 77.9 / 77.7 / 74.6 | 73.4 / 72.1 / 67.1
  // A wider search buys nothing on meshes and LOSES on overlapping soups; 2 starts to cost on the knot.  4: PLOC 25.8 -> 17.4 ms at 100M triangles.
  int lbvh = (o.flags & LISA_FLAG_LBVH) ? 1 : 0, radius = 4;
  if (const char* e = getenv("LISA_BUILDER")) lbvh = !strcmp(e, "lbvh");
  if (const char* e = getenv("LISA_PLOC_RADIUS")) radius = std::max(1, atoi(e));
  int rotate = 0;
  if (const char* e = getenv("LISA_BVH_ROTATE")) rotate = std::max(0, std::min(8, atoi(e)));
  // triangle splitting (split.cu), opt-in: the builder's primitives become references to the triangles
  float split_budget = (o.flags & LISA_FLAG_SPLIT_TRIANGLES) ? 1.5f : 0.0f;
  if (const char* e = getenv("LISA_SPLIT")) { const float v = (float)atof(e); split_budget = v <= 0.0f ? 0.0f : v <= 1.0f ? 1.5f : std::min(v, 8.0f); }
  const int   T_in = T;
  SplitOutput sp{};
  int         rc = 0;
  if (split_budget > 0.0f && T > 0) {
    rc = split_triangles(d_verts, T, split_budget, &sp, c->stream, g_err, sizeof(g_err));
    if (!rc) T = sp.num_refs;
  }
  // leaves by the surface-area heuristic (k_collapse8): the cost of one more child slot, in triangle tests; LISA_LEAF_SAH=-1: off
  float leaf_sah = LISA_LEAF_SAH_DEFAULT;
  if (const char* e = getenv("LISA_LEAF_SAH")) leaf_sah = (float)atof(e);
  BuildInput bi{d_verts, d_normals, d_mat_idx, d_emit, T, sd->num_materials, wide, lbvh, radius, rotate, leaf_sah, sp.d_ref_tri, sp.d_ref_lo, sp.d_ref_hi};
  if (!rc) rc = build_bvh(bi, &c->bvh, c->stream, g_err, sizeof(g_err));
  if (rc) {  // keep the builder's message: a sticky CUDA error would otherwise be reported by the next call instead
    dev_free(d_verts); dev_free(d_normals); dev_free(d_mat_idx); dev_free(d_emit); split_free(&sp);
    return rc;
  }
  c->bvh.num_input_tris = T_in;
  CU(cudaEventRecord(t2, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  dev_free(d_verts); dev_free(d_normals); dev_free(d_mat_idx); dev_free(d_emit); split_free(&sp);
  cudaEventElapsedTime(&c->stats.upload_ms, t0, t1);
  cudaEventElapsedTime(&c->stats.bvh_build_ms, t1, t2);
  c->stats.build_sort_ms = c->bvh.stage_ms[0]; c->stats.build_hierarchy_ms = c->bvh.stage_ms[1];
  c->stats.build_collapse_ms = c->bvh.stage_ms[2]; c->stats.build_pack_ms = c->bvh.stage_ms[3];
  }

  c->scene.tri_v = c->bvh.d_tri_v;
  c->scene.tri_n = c->bvh.d_tri_n;
  c->scene.mats = c->d_mats;
  c->scene.bvh = c->bvh.d_nodes;
  c->scene.root_other = c->bvh.root_other;
  c->scene.root_emit = c->bvh.root_emit;
  c->scene.wide = wide;
  c->scene.num_tris = T;
  c->scene.num_mats = sd->num_materials;
  c->scene.single_light = single_light;
  c->scene.shadow_first_found = o.shadow_mode == LISA_SHADOW_FIRST_FOUND;
  {  // emitter bounds, padded by a relative + absolute margin so the quick reject is conservative
    const float* b = c->bvh.box_emit;
    float pad[3];
    for (int k = 0; k < 3; k++) pad[k] = 1e-5f * (fabsf(b[k]) + fabsf(b[3 + k])) + 1e-6f * (b[3 + k] - b[k]) + 1e-30f;
    c->scene.emit_lo = make_float3(b[0] - pad[0], b[1] - pad[1], b[2] - pad[2]);
    c->scene.emit_hi = make_float3(b[3] + pad[0], b[4] + pad[1], b[5] + pad[2]);
    if (c->bvh.num_emit_tris > 0) {
      const float3 lo = c->scene.emit_lo, hi = c->scene.emit_hi;
      c->scene.emit_c = make_float3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
      const float dx = 0.5f * (hi.x - lo.x), dy = 0.5f * (hi.y - lo.y), dz = 0.5f * (hi.z - lo.z);
      c->scene.emit_r2 = (dx * dx + dy * dy + dz * dz) * 1.0001f;
    } else {  // no emitter: nothing can ever be lit; an empty sphere culls every non-sticky try
      c->scene.emit_c = make_float3(0, 0, 0);
      c->scene.emit_r2 = -1.0f;
    }
    // a quad light: the (unpadded) emitter bounds are flat along one axis
    c->scene.emit_flat = -1;
    c->scene.emit_plane = 0.0f;
    c->scene.emit_hh = 0.0f;
    if (c->bvh.num_emit_tris > 0 && !(getenv("LISA_FLAT_LIGHT") && !atoi(getenv("LISA_FLAT_LIGHT")))) {
      const float ext[3] = {b[3] - b[0], b[4] - b[1], b[5] - b[2]};
      const float big = std::max(ext[0], std::max(ext[1], ext[2]));
      for (int k = 0; k < 3; k++)
        if (big > 0.0f && ext[k] <= 1e-6f * big && ext[(k + 1) % 3] > 0.0f && ext[(k + 2) % 3] > 0.0f) {
          const float lo = (&c->scene.emit_lo.x)[k], hi = (&c->scene.emit_hi.x)[k];  // the PADDED bounds, which the confirmation tests
          c->scene.emit_flat = k;
          c->scene.emit_plane = 0.5f * (lo + hi);
          c->scene.emit_hh = 0.5f * (hi - lo) * 1.01f + 4e-6f * (big + fabsf(lo) + fabsf(hi));
          break;
        }
    }
    c->scene.cull = (o.flags & LISA_FLAG_NO_CULL) ? 0 : 1;
    if (const char* e = getenv("LISA_CULL")) c->scene.cull = atoi(e) != 0;
  }

  c->stats.struct_size = sizeof(lisa_stats);
  c->stats.num_triangles = (uint32_t)c->bvh.num_input_tris;
  c->stats.num_references = (uint32_t)T;
  c->stats.num_emitter_triangles = (uint32_t)c->bvh.num_emit_tris;
  c->stats.bvh_nodes = (uint32_t)c->bvh.num_nodes;
  c->stats.bvh_sah_nodes_per_ray = c->bvh.sah_nodes_per_ray;
  c->stats.bvh_emitter_nodes = (uint32_t)c->bvh.nodes_emit;
  c->stats.bvh_bytes = c->bvh.node_bytes;
  c->stats.triangle_bytes = (uint64_t)T * 96;

  // ---- accumulators, counters
  const size_t npix = (size_t)c->width * c->height;
  CU(dev_alloc((void**)&c->d_accum, sizeof(float4) * std::max<size_t>(npix, 1)));
  CU(cudaMemsetAsync(c->d_accum, 0, sizeof(float4) * npix, c->stream));
  CU(dev_alloc((void**)&c->state.ring, sizeof(unsigned int) * 64));
  CU(dev_alloc((void**)&c->state.stats, sizeof(unsigned long long) * 16));
  CU(cudaMemsetAsync(c->state.stats, 0, sizeof(unsigned long long) * 16, c->stream));
  c->h_stats = (unsigned long long*)pinned_slot_alloc();
  if (!c->h_stats) return fail(LISA_ERR_NOMEM, "pinned host allocation failed");
  memset(c->h_stats, 0, 256);

  c->cfg.sm_count = di.sms;
  c->cfg.extend_block = 128;  // __launch_bounds__(128, 6): 80 registers/thread, 6 CTAs = 24 warps per SM
  c->cfg.shadow_block = 128;
  if (const char* e2 = getenv("LISA_EXTEND_BLOCK")) c->cfg.extend_block = std::max(32, std::min(128, atoi(e2) / 32 * 32));
  if (const char* e2 = getenv("LISA_SHADOW_BLOCK")) c->cfg.shadow_block = std::max(32, std::min(128, atoi(e2) / 32 * 32));
  c->cfg.idle_thresh = 20;  // measured on B200 (Cornell 2000x2000): 4 -> 784, 8 -> 790, 12 -> 799, 16 -> 808, 20 -> 811, 24 -> 789 Msamples/s
  if (const char* e2 = getenv("LISA_IDLE_THRESH")) c->cfg.idle_thresh = std::max(1, std::min(32, atoi(e2)));
  c->cfg.idle_thresh_rays = c->cfg.idle_thresh;
  if (const char* e2 = getenv("LISA_IDLE_THRESH_RAYS")) c->cfg.idle_thresh_rays = std::max(1, std::min(32, atoi(e2)));
  // measured on B200 (Cornell 2000x2000): 4 passes 668, 3: 683, 2: 718, 1: 744 Msamples/s — the fill/drain of every extra
  // persistent launch (~14 us) costs more than finishing the few re-tries inline in k_rays
  c->cfg.shadow_passes = 1;
  c->cfg.defer_retries = 0;  // measured: no gain (813 vs 811 Msamples/s) for 16 % more iterations
  if (const char* e2 = getenv("LISA_DEFER")) c->cfg.defer_retries = atoi(e2) != 0;
  if (const char* e2 = getenv("LISA_SHADOW_PASSES")) c->cfg.shadow_passes = atoi(e2);
  {
    const int w = wide != 0;
    if (!di.ok || getenv("LISA_EXTEND_BLOCK") || getenv("LISA_SHADOW_BLOCK")) {
      di.occ_tries = tries_occupancy(256);
      for (int k = 0; k < 2; k++) { di.occ_ext[k] = extend_occupancy(k != 0, c->cfg.extend_block); di.occ_rays[k] = shadow_occupancy(k != 0, c->cfg.shadow_block); di.occ_path[k] = path_occupancy(k != 0, 128); di.occ_pool[k] = pool_occupancy(k != 0, false); }
      di.occ_pool_deep = pool_occupancy(true, true);
      di.ok = true;
    }
    c->cfg.tries_blocks_per_sm = di.occ_tries;
    c->cfg.extend_blocks_per_sm = di.occ_ext[w];
    c->cfg.shadow_blocks_per_sm = di.occ_rays[w];
    c->cfg.path_blocks_per_sm = di.occ_path[w];
    c->cfg.pool_blocks_per_sm = di.occ_pool[w];
    c->cfg.pool_blocks_per_sm_deep = di.occ_pool_deep;
  }
  c->pipeline = (o.flags & LISA_FLAG_WAVEFRONT) ? 0 : 3;
  if (const char* e2 = getenv("LISA_PIPELINE"))
    c->pipeline = strcmp(e2, "wavefront") == 0 ? 0 : strcmp(e2, "pool") == 0 ? 2 : strcmp(e2, "path") == 0 ? 1 : 3;
  c->cfg.pool_dry_thresh = 16;
  c->cfg.pool_dry_thresh_deep = 2;  // measured on the 10M soup (deep flavour): 16 -> 68.6, 8 -> 71.0, 4 -> 72.4, 2 -> 73.0 Msamples/s
  if (const char* e2 = getenv("LISA_DRY_THRESH")) c->cfg.pool_dry_thresh = c->cfg.pool_dry_thresh_deep = std::max(1, std::min(32, atoi(e2)));
  if (const char* e2 = getenv("LISA_POOL_BLOCKS_PER_SM")) {
    c->cfg.pool_blocks_per_sm = std::max(1, std::min(c->cfg.pool_blocks_per_sm, atoi(e2)));
    c->cfg.pool_blocks_per_sm_deep = std::max(1, std::min(c->cfg.pool_blocks_per_sm_deep, atoi(e2)));
  }
  // k_pool's flavour (sched_pool.cuh) by the builder's surface-area estimate (sum of the wide nodes' areas over the root's: the
  // node visits of a LINE through the whole scene, several times what a ray segment that ends at its first hit visits).
  // Cornell box 3.0, 871k-triangle knot 3.4 (shallow +9 %); soups of overlapping triangles: 3k 15 (shallow +2 %), 10k-100k
  // 23-52 (equal), 300k 76 (deep +3 %), 1M 116 (deep +10 %), 10M 251 (deep +14 %).
  float deep_from = LISA_POOL_DEEP_SAH;
  if (const char* e2 = getenv("LISA_POOL_DEEP_SAH")) deep_from = (float)atof(e2);
  c->cfg.pool_deep = wide && c->bvh.sah_nodes_per_ray > deep_from;
  c->pool_flavour_auto = true;
  if (const char* e2 = getenv("LISA_POOL_FLAVOUR")) { c->cfg.pool_deep = wide && !strcmp(e2, "deep"); c->pool_flavour_auto = false; }
  c->stats.pool_flavour = c->cfg.pool_deep ? 1u : 0u;
  c->pool_min_chains_factor = 1.2f;
  if (const char* e2 = getenv("LISA_POOL_MIN_FACTOR")) c->pool_min_chains_factor = (float)atof(e2);
  if (getenv("LISA_DEBUG_TIMING")) fprintf(stderr, "[lisa] pipeline %d, k_path %d CTAs/SM, k_pool %d CTAs/SM (%s flavour: SAH estimate %.2f node visits per ray)\n", c->pipeline, c->cfg.path_blocks_per_sm,
                                         c->cfg.pool_deep ? c->cfg.pool_blocks_per_sm_deep : c->cfg.pool_blocks_per_sm, c->cfg.pool_deep ? "deep" : "shallow", c->bvh.sah_nodes_per_ray);
  if ((c->width > 65535u || c->height > 65535u) && c->pipeline) c->pipeline = 0;  // k_path packs a chain's pixel as x | y << 16
  c->cfg.path_wait_thresh = 16;  // measured on B200 (Cornell 2000x2000): 8 -> 1057, 12 -> 1090, 16 -> 1115, 20 -> 1107, 24 -> 1073, 28 -> 999 Msamples/s
  if (const char* e2 = getenv("LISA_WAIT_THRESH")) c->cfg.path_wait_thresh = std::max(1, std::min(32, atoi(e2)));
  if (const char* e2 = getenv("LISA_PATH_BLOCKS_PER_SM")) c->cfg.path_blocks_per_sm = std::max(1, std::min(c->cfg.path_blocks_per_sm, atoi(e2)));
  if (const char* e2 = getenv("LISA_SHADOW_BLOCKS_PER_SM")) c->cfg.shadow_blocks_per_sm = std::max(1, atoi(e2));
  // default residency: enough chains to fill the machine several times over, bounded so the state stays
  // a small fraction of HBM (112 B per chain)
  c->max_chains = o.max_chains ? o.max_chains : (4u << 20);
  CU(cudaStreamSynchronize(c->stream));
  return LISA_OK;
}

extern "C" int lisa_create(const lisa_scene_desc* sd, const lisa_options* opt, lisa_ctx** out) {
  if (!sd || !out) return fail(LISA_ERR_ARG, "lisa_create: null argument");
  *out = nullptr;
  if (sd->num_vertices < 0 || sd->num_vertices % 3) return fail(LISA_ERR_ARG, "num_vertices (%d) must be a non-negative multiple of 3", sd->num_vertices);
  if (sd->num_vertices && (!sd->vertices || !sd->normals || !sd->mat_indices)) return fail(LISA_ERR_ARG, "null geometry array");
  if (sd->num_materials < 0 || (sd->num_materials && !sd->materials)) return fail(LISA_ERR_ARG, "bad materials");
  if (!sd->width || !sd->height) return fail(LISA_ERR_ARG, "width and height must be positive");
  if ((uint64_t)sd->width * sd->height > (1ull << 31)) return fail(LISA_ERR_ARG, "image too large");
  if (sd->num_materials > 65535) return fail(LISA_ERR_ARG, "at most 65535 materials");
  const int T = sd->num_vertices / 3;
  for (int t = 0; t < T; t++)
    if (sd->mat_indices[t] < 0 || sd->mat_indices[t] >= sd->num_materials)
      return fail(LISA_ERR_ARG, "triangle %d: material index %d out of range [0, %d)", t, sd->mat_indices[t], sd->num_materials);
  lisa_ctx* c = new lisa_ctx();
  int rc = create_impl(sd, opt, c);
  if (rc != LISA_OK) {
    std::string keep = g_err;
    lisa_destroy(c);
    snprintf(g_err, sizeof(g_err), "%s", keep.c_str());
    return rc;
  }
  *out = c;
  return LISA_OK;
}

extern "C" int lisa_create_from_bvh(const lisa_scene_desc* sd, const lisa_options* opt, const char* path, lisa_ctx** out) {
  if (!sd || !out || !path || !*path) return fail(LISA_ERR_ARG, "lisa_create_from_bvh: null argument");
  *out = nullptr;
  if (sd->num_vertices < 0 || sd->num_vertices % 3) return fail(LISA_ERR_ARG, "num_vertices (%d) must be a non-negative multiple of 3", sd->num_vertices);
  if (sd->num_materials < 0 || (sd->num_materials && !sd->materials)) return fail(LISA_ERR_ARG, "bad materials");
  if (!sd->width || !sd->height) return fail(LISA_ERR_ARG, "width and height must be positive");
  if ((uint64_t)sd->width * sd->height > (1ull << 31)) return fail(LISA_ERR_ARG, "image too large");
  lisa_ctx* c = new lisa_ctx();
  int rc = create_impl(sd, opt, c, path);
  if (rc != LISA_OK) {
    std::string keep = g_err;
    lisa_destroy(c);
    snprintf(g_err, sizeof(g_err), "%s", keep.c_str());
    return rc;
  }
  *out = c;
  return LISA_OK;
}

extern "C" int lisa_save_bvh(lisa_ctx* c, const char* path) {
  if (!c || !path || !*path) return fail(LISA_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  BvhFileHeader h{};
  memcpy(h.magic, "LISABVH1", 8);
  h.version = 3; h.header_bytes = sizeof(h);
  h.sah_nodes_per_ray = c->bvh.sah_nodes_per_ray;
  h.num_input_tris = (uint64_t)c->bvh.num_input_tris;
  h.num_tris = (uint64_t)c->scene.num_tris; h.num_nodes = (uint64_t)c->bvh.num_nodes;
  h.nodes_other = c->bvh.nodes_other; h.nodes_emit = c->bvh.nodes_emit;
  h.root_other = c->bvh.root_other; h.root_emit = c->bvh.root_emit;
  h.num_emit_tris = c->bvh.num_emit_tris; h.wide = c->scene.wide;
  h.num_materials = c->scene.num_mats; h.emit_hash = c->emit_hash;
  memcpy(h.box_other, c->bvh.box_other, sizeof(h.box_other));
  memcpy(h.box_emit, c->bvh.box_emit, sizeof(h.box_emit));
  h.file_bytes = bvh_file_bytes(h);
  const std::string tmp = std::string(path) + ".tmp";  // written aside and renamed
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return fail(LISA_ERR_IO, "cannot open %s for writing", tmp.c_str());
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  const size_t T = (size_t)h.num_tris;
  std::vector<char> buf(std::min<size_t>(64u << 20, std::max<size_t>(48 * std::max<size_t>(T, 1), 80)));
  auto pump = [&](const void* src, size_t bytes) {
    for (size_t off = 0; off < bytes && ok; off += buf.size()) {
      const size_t n = std::min(buf.size(), bytes - off);
      ok = cudaMemcpyAsync(buf.data(), (const char*)src + off, n, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess &&
           cudaStreamSynchronize(c->stream) == cudaSuccess && fwrite(buf.data(), 1, n, f) == n;
    }
  };
  if (T) {
    pump(c->bvh.d_nodes, (size_t)h.num_nodes * (h.wide ? 80 : 64));
    pump(c->bvh.d_tri_v, 48 * T);
    pump(c->bvh.d_tri_n, 48 * T);
    pump(c->bvh.d_final_to_orig, 4 * T);
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(LISA_ERR_IO, "short write to %s", path); }
  return LISA_OK;
}

extern "C" int lisa_reset_accum(lisa_ctx* c) {
  if (!c) return fail(LISA_ERR_ARG, "null ctx");
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(c->d_accum, 0, sizeof(float4) * (size_t)c->width * c->height, c->stream));
  CU(cudaMemsetAsync(c->state.stats, 0, sizeof(unsigned long long) * 16, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->stats.samples = c->stats.radiance_rays = c->stats.shadow_rays = c->stats.null_directions = 0;
  c->stats.kernel_launches = c->stats.iterations = 0;
  c->stats.render_ms = 0;
  c->stats.subframes_accumulated = 0;
  return LISA_OK;
}

// Runs one tile (pixel range x subframe range) to completion.
static int run_tile(lisa_ctx* c, const Tile& t, uint64_t* launches, uint64_t* iterations) {
  int rc = ensure_state(c, t.n_chains);
  if (rc) return rc;
  if (c->pipeline >= 1) {
    if (c->profile_stages) cudaEventRecord(next_event(c), c->stream);
    // k_pool keeps 64 chains per warp (189,440 on B200): it pays once the tile fills those slots.  Measured, k_pool vs k_path on the
    // Cornell box (chains / slots): 256x256 (0.35) 11.6 vs 8.1 ms, knot 480x270 (0.68) 24.0 vs 22.9, 512x512 x 64 spp (1.38) 15.3 vs
    // 16.0, 1280x720 (4.9) 6.8 vs 7.3, 1024x1024 (5.5) 11.6 vs 12.7, 2000x2000 (21) 1564 vs 1254 Msamples/s.  Using fewer slots per
    // warp so that a tile's rounds are equally full LOSES (512x512: 45 slots 16.6 ms, 32 slots 17.0 ms against 15.3 with all 64).
    const bool     deep = c->cfg.pool_deep != 0;
    const uint64_t pool_slots = (uint64_t)c->cfg.sm_count * (deep ? c->cfg.pool_blocks_per_sm_deep : c->cfg.pool_blocks_per_sm) * pool_chains_per_cta(deep);
    const bool use_pool = c->pipeline == 2 || (c->pipeline == 3 && (double)t.n_chains >= (double)c->pool_min_chains_factor * (double)pool_slots);
    nvtxRangePushA(use_pool ? "lisa: tile k_pool" : "lisa: tile k_path");
    if (use_pool) launch_pool(c->scene, c->state, c->cam, t, c->cfg, c->stream);
    else launch_path(c->scene, c->state, c->cam, t, c->cfg, c->stream);
    nvtxRangePop();
    if (c->profile_stages) { cudaEvent_t e = next_event(c); cudaEventRecord(e, c->stream); cudaEventRecord(next_event(c), c->stream); }
    launch_finalize(c->state, c->cam, t, c->d_accum, c->stream);
    *launches += 2;
    *iterations += 1;
    if (c->profile_stages) { CU(cudaStreamSynchronize(c->stream)); drain_stage_events(c); }
    return LISA_OK;
  }
  launch_init_chains(c->state, c->cam, t, c->stream);
  (*launches)++;
  // Every chain needs at least spp iterations (typically ~4.5 per sample); poll the finished-chain counter in bursts.
  uint32_t iter = 0;
  // a bounce takes 1 iteration, plus one for every failed light-sampling candidate when retries are deferred (<= 30)
  const uint64_t max_iter = (uint64_t)t.spp * std::max(t.bounces, 1u) * (LISA_SHADOW_TRIES + 1) + 8;
  // a sample takes at least one iteration, so nothing can finish before spp iterations
  uint32_t burst = std::max<uint32_t>(1, t.spp);
  while (true) {
    for (uint32_t k = 0; k < burst; k++, iter++) {
      if (c->profile_stages) cudaEventRecord(next_event(c), c->stream);
      launch_extend(c->scene, c->state, c->cam, t, iter, c->cfg, c->stream);
      if (c->profile_stages) cudaEventRecord(next_event(c), c->stream);
      const int nl = launch_shadow(c->scene, c->state, t, iter, c->cfg, c->stream);
      if (c->profile_stages) cudaEventRecord(next_event(c), c->stream);
      *launches += 1 + nl;
    }
    CU(cudaMemcpyAsync(c->h_stats, c->state.stats, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->profile_stages) drain_stage_events(c);
    if (c->h_stats[4] >= t.n_chains) break;
    if (iter > max_iter) return fail(LISA_ERR_STATE, "wavefront did not converge after %u iterations (%llu of %u chains done)", iter,
                                     (unsigned long long)c->h_stats[4], t.n_chains);
    // iterations after the last chain finished are near-free (every thread exits on its sample count)
    burst = std::max<uint32_t>(8, t.spp / 16);
  }
  *iterations += iter;
  launch_finalize(c->state, c->cam, t, c->d_accum, c->stream);
  (*launches)++;
  return LISA_OK;
}

extern "C" int lisa_render_subframes(lisa_ctx* c, uint32_t first, uint32_t count, uint32_t spp) {
  if (!c) return fail(LISA_ERR_ARG, "null ctx");
  if (!count || !spp) return fail(LISA_ERR_ARG, "count and spp must be positive");
  CU(cudaSetDevice(c->device));
  const uint64_t npix = (uint64_t)c->width * c->height;
  unsigned long long before[16];
  CU(cudaMemcpyAsync(c->h_stats, c->state.stats, sizeof(before), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(before, c->h_stats, sizeof(before));
  uint64_t launches = 0, iterations = 0;
  c->stage_ms[0] = c->stage_ms[1] = 0;
  c->stage_launches[0] = c->stage_launches[1] = 0;
  CU(cudaEventRecord(c->ev0, c->stream));
  if (c->num_bounces == 0) {
    // shader.cu:110: zero bounces trace nothing; every sample is black
  } else {
    // tiles: as many whole subframes as fit in max_chains, else pixel ranges of one subframe
    const uint64_t cap = std::max<uint64_t>(c->max_chains, 1024);
    if (npix <= cap) {
      const uint32_t per = (uint32_t)std::max<uint64_t>(1, cap / npix);
      for (uint32_t f = 0; f < count; f += per) {
        Tile t{0, (uint32_t)npix, first + f, std::min(per, count - f), spp, c->num_bounces, 0};
        t.n_chains = t.npix * t.nf;
        int rc = run_tile(c, t, &launches, &iterations);
        if (rc) return rc;
      }
    } else {
      for (uint32_t f = 0; f < count; f++)
        for (uint64_t p0 = 0; p0 < npix; p0 += cap) {
          Tile t{(uint32_t)p0, (uint32_t)std::min<uint64_t>(cap, npix - p0), first + f, 1, spp, c->num_bounces, 0};
          t.n_chains = t.npix;
          int rc = run_tile(c, t, &launches, &iterations);
          if (rc) return rc;
        }
    }
  }
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaMemcpyAsync(c->h_stats, c->state.stats, sizeof(before), cudaMemcpyDeviceToHost, c->stream));
  CU(fetch_trav_overflow(overflow_slot(c), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (*overflow_slot(c)) return stack_overflow_error();
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  lisa_stats& s = c->stats;
  s.last_render_ms = ms;
  s.render_ms += ms;
  s.last_radiance_rays = c->h_stats[0] - before[0];
  s.last_shadow_rays = c->h_stats[1] - before[1];
  s.last_samples = npix * count * spp;
  s.last_kernel_launches = launches;
  s.last_extend_ms = c->stage_ms[0];
  s.last_shadow_ms = c->stage_ms[1];
  s.last_extend_launches = c->profile_stages ? c->stage_launches[0] : iterations;
  s.last_shadow_launches = c->profile_stages ? c->stage_launches[1] : iterations;
  s.last_shadow_jobs = c->h_stats[7] - before[7];
  s.last_shadow_culled = c->h_stats[8] - before[8];
  s.shadow_culled = c->h_stats[8];
  s.last_nodes_visited = c->h_stats[5] - before[5];
  s.last_triangles_tested = c->h_stats[6] - before[6];
  s.nodes_visited = c->h_stats[5];
  s.triangles_tested = c->h_stats[6];
  s.radiance_rays = c->h_stats[0];
  s.shadow_rays = c->h_stats[1];
  s.null_directions = c->h_stats[3];
  s.samples += npix * count * spp;
  s.kernel_launches += launches;
  s.iterations += iterations;
  s.subframes_accumulated += count;
  // The estimate at create time is geometric (a line through the whole scene); once a call has traced enough rays the flavour
  // follows what they measured.  Crossover on B200 (soups, node visits per traversed ray): 4.1 shallow +2 %, 6.5-9.6 equal,
  // 10.6 deep +3 %, 27 deep +10 %.  Needs the traversal counters (LISA_COUNT_TRAVERSAL, on by default); same bits either way.
  if (c->pool_flavour_auto && c->scene.wide && s.last_nodes_visited > 0) {
    const uint64_t traced = s.last_radiance_rays + s.last_shadow_rays - s.last_shadow_culled;
    if (traced >= 100000) {
      const double per_ray = (double)s.last_nodes_visited / (double)traced;
      if (per_ray > 10.0) c->cfg.pool_deep = 1;
      else if (per_ray < 8.0) c->cfg.pool_deep = 0;
      s.pool_flavour = c->cfg.pool_deep ? 1u : 0u;
    }
  }
  return LISA_OK;
}

static int resolve(lisa_ctx* c, bool want_mean, bool want_rgba) {
  const size_t npix = (size_t)c->width * c->height;
  if (want_mean && !c->d_mean) CU(dev_alloc((void**)&c->d_mean, sizeof(float4) * npix));
  if (want_rgba && !c->d_rgba8) CU(dev_alloc((void**)&c->d_rgba8, sizeof(uint32_t) * npix));
  launch_resolve(c->d_accum, (uint32_t)npix, want_mean ? c->d_mean : nullptr, want_rgba ? c->d_rgba8 : nullptr, c->stream);
  return LISA_OK;
}

extern "C" int lisa_read_accum(lisa_ctx* c, float* rgba) {
  if (!c || !rgba) return fail(LISA_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  int rc = resolve(c, true, false);
  if (rc) return rc;
  CU(cudaMemcpyAsync(rgba, c->d_mean, sizeof(float4) * (size_t)c->width * c->height, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LISA_OK;
}

extern "C" int lisa_read_rgba8(lisa_ctx* c, uint8_t* rgba) {
  if (!c || !rgba) return fail(LISA_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  int rc = resolve(c, false, true);
  if (rc) return rc;
  CU(cudaMemcpyAsync(rgba, c->d_rgba8, sizeof(uint32_t) * (size_t)c->width * c->height, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return LISA_OK;
}

extern "C" int lisa_write_ppm(lisa_ctx* c, const char* path) {
  if (!c) return fail(LISA_ERR_ARG, "null ctx");
  if (!path) path = c->output_image.c_str();
  if (!path || !*path) return fail(LISA_ERR_ARG, "no output path");
  std::vector<uint8_t> px((size_t)c->width * c->height * 4);
  int rc = lisa_read_rgba8(c, px.data());
  if (rc) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(LISA_ERR_IO, "cannot open %s for writing", path);
  // sutil::savePPM (sutil.cpp:97-117) after saveImage's flip and alpha drop (sutil.cpp:523-554)
  fprintf(f, "P6\n%u %u\n255\n", c->width, c->height);
  std::vector<uint8_t> row((size_t)c->width * 3);
  for (int y = (int)c->height - 1; y >= 0; y--) {
    const uint8_t* src = px.data() + (size_t)y * c->width * 4;
    for (uint32_t x = 0; x < c->width; x++) { row[3 * x] = src[4 * x]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x + 2]; }
    if (fwrite(row.data(), 1, row.size(), f) != row.size()) { fclose(f); return fail(LISA_ERR_IO, "short write to %s", path); }
  }
  fclose(f);
  return LISA_OK;
}

// ---- PNG (RGBA8, no compression: zlib "stored" blocks) --------------------------------------------------------
static uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  static uint32_t table[256];
  static bool     ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    ready = true;
  }
  for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
  return crc;
}
static void put_be32(std::vector<uint8_t>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((uint8_t)(x >> s)); }
static bool png_chunk(FILE* f, const char* type, const std::vector<uint8_t>& data) {
  std::vector<uint8_t> head;
  put_be32(head, (uint32_t)data.size());
  uint32_t crc = crc32_update(0xffffffffu, (const uint8_t*)type, 4);
  crc = crc32_update(crc, data.data(), data.size()) ^ 0xffffffffu;
  std::vector<uint8_t> tail;
  put_be32(tail, crc);
  return fwrite(head.data(), 1, 4, f) == 4 && fwrite(type, 1, 4, f) == 4 &&
         (data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size()) && fwrite(tail.data(), 1, 4, f) == 4;
}

// sutil::saveImage's PNG branch (sutil.cpp:600-613): the uchar4 frame, all four components, flipped vertically
static int write_png(lisa_ctx* c, const char* path) {
  std::vector<uint8_t> px((size_t)c->width * c->height * 4);
  int rc = lisa_read_rgba8(c, px.data());
  if (rc) return rc;
  const size_t row = (size_t)c->width * 4;
  std::vector<uint8_t> raw;  // filter byte 0 + row, rows top to bottom
  raw.reserve((row + 1) * c->height);
  for (int y = (int)c->height - 1; y >= 0; y--) {
    raw.push_back(0);
    raw.insert(raw.end(), px.begin() + (size_t)y * row, px.begin() + (size_t)(y + 1) * row);
  }
  std::vector<uint8_t> z;  // zlib stream of stored blocks
  z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;  // adler32
  for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
    const size_t n = std::min<size_t>(65535, raw.size() - off);
    z.push_back(off + n >= raw.size() ? 1 : 0);
    z.push_back((uint8_t)(n & 0xff)); z.push_back((uint8_t)(n >> 8));
    z.push_back((uint8_t)(~n & 0xff)); z.push_back((uint8_t)((~n >> 8) & 0xff));
    z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    for (size_t i = off; i < off + n; i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
    if (raw.empty()) break;
  }
  put_be32(z, (b << 16) | a);
  FILE* f = fopen(path, "wb");
  if (!f) return fail(LISA_ERR_IO, "cannot open %s for writing", path);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, c->width); put_be32(ihdr, c->height);
  ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);  // 8 bit RGBA
  bool ok = fwrite(sig, 1, 8, f) == 8 && png_chunk(f, "IHDR", ihdr) && png_chunk(f, "IDAT", z) && png_chunk(f, "IEND", {});
  ok = (fclose(f) == 0) && ok;
  return ok ? LISA_OK : fail(LISA_ERR_IO, "short write to %s", path);
}

extern "C" int lisa_write_image(lisa_ctx* c, const char* path) {
  if (!c) return fail(LISA_ERR_ARG, "null ctx");
  if (!path) path = c->output_image.c_str();
  const std::string filename(path ? path : "");
  // sutil::saveImage (sutil.cpp:523-530, 600, 636, 685-688): the LAST THREE characters name the format
  if (filename.length() < 5) return fail(LISA_ERR_ARG, "sutil::saveImage(): Failed to determine filename extension");
  const std::string ext = filename.substr(filename.length() - 3);
  if (ext == "PPM" || ext == "ppm") return lisa_write_ppm(c, filename.c_str());
  if (ext == "PNG" || ext == "png") return write_png(c, filename.c_str());
  if (ext == "EXR" || ext == "exr") return fail(LISA_ERR_ARG, "sutil::saveImage(): saving of uchar4 images to EXR not implemented yet");
  return fail(LISA_ERR_ARG, "sutil::saveImage(): Failed unsupported filetype '%s'", ext.c_str());
}

extern "C" int lisa_write_pfm(lisa_ctx* c, const char* path) {
  if (!c || !path || !*path) return fail(LISA_ERR_ARG, "null argument");
  std::vector<float> px((size_t)c->width * c->height * 4);
  int rc = lisa_read_accum(c, px.data());
  if (rc) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(LISA_ERR_IO, "cannot open %s for writing", path);
  // Portable Float Map: "PF", width height, negative scale = little endian; scanlines bottom to top, which is the
  // accumulators' own order (row 0 = bottom)
  fprintf(f, "PF\n%u %u\n-1.0\n", c->width, c->height);
  std::vector<float> row((size_t)c->width * 3);
  for (uint32_t y = 0; y < c->height; y++) {
    const float* src = px.data() + (size_t)y * c->width * 4;
    for (uint32_t x = 0; x < c->width; x++) { row[3 * x] = src[4 * x]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x + 2]; }
    if (fwrite(row.data(), sizeof(float), row.size(), f) != row.size()) { fclose(f); return fail(LISA_ERR_IO, "short write to %s", path); }
  }
  fclose(f);
  return LISA_OK;
}

// accumulator checkpoint: "LISAACC1", width, height, subframes (u32 each), then W*H float4
extern "C" int lisa_save_accum(lisa_ctx* c, const char* path) {
  if (!c || !path || !*path) return fail(LISA_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  const size_t npix = (size_t)c->width * c->height;
  std::vector<float4> px(npix);
  CU(cudaMemcpyAsync(px.data(), c->d_accum, sizeof(float4) * npix, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const std::string tmp = std::string(path) + ".tmp";  // written aside and renamed: an interrupted save never truncates a good file
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return fail(LISA_ERR_IO, "cannot open %s for writing", tmp.c_str());
  const uint32_t hdr[3] = {c->width, c->height, c->stats.subframes_accumulated};
  bool ok = fwrite("LISAACC1", 1, 8, f) == 8 && fwrite(hdr, sizeof(uint32_t), 3, f) == 3 && fwrite(px.data(), sizeof(float4), npix, f) == npix;
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(LISA_ERR_IO, "short write to %s", path); }
  return LISA_OK;
}

extern "C" int lisa_load_accum(lisa_ctx* c, const char* path, uint32_t* subframes) {
  if (!c || !path || !*path) return fail(LISA_ERR_ARG, "null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(LISA_ERR_IO, "cannot open %s", path);
  char     magic[8];
  uint32_t hdr[3];
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "LISAACC1", 8) != 0 || fread(hdr, sizeof(uint32_t), 3, f) != 3) {
    fclose(f);
    return fail(LISA_ERR_IO, "%s is not an accumulator checkpoint", path);
  }
  if (hdr[0] != c->width || hdr[1] != c->height) {
    fclose(f);
    return fail(LISA_ERR_ARG, "%s holds a %ux%u image, the context renders %ux%u", path, hdr[0], hdr[1], c->width, c->height);
  }
  const size_t npix = (size_t)c->width * c->height;
  std::vector<float4> px(npix);
  const bool ok = fread(px.data(), sizeof(float4), npix, f) == npix;
  fclose(f);
  if (!ok) return fail(LISA_ERR_IO, "%s is truncated", path);
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->d_accum, px.data(), sizeof(float4) * npix, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->stats.subframes_accumulated = hdr[2];
  if (subframes) *subframes = hdr[2];
  return LISA_OK;
}

extern "C" int lisa_get_stats(lisa_ctx* c, lisa_stats* out) {
  if (!c || !out) return fail(LISA_ERR_ARG, "null argument");
  uint32_t sz = out->struct_size ? std::min<uint32_t>(out->struct_size, sizeof(lisa_stats)) : sizeof(lisa_stats);
  c->stats.struct_size = sizeof(lisa_stats);
  memcpy(out, &c->stats, sz);
  return LISA_OK;
}

extern "C" int lisa_accum_add_peer(lisa_ctx* dst, lisa_ctx* src) {
  if (!dst || !src) return fail(LISA_ERR_ARG, "null ctx");
  if (dst->width != src->width || dst->height != src->height) return fail(LISA_ERR_ARG, "accumulators differ in size");
  CU(cudaSetDevice(src->device));
  CU(cudaStreamSynchronize(src->stream));
  CU(cudaSetDevice(dst->device));
  const uint32_t npix = dst->width * dst->height;
  const float4*  from = src->d_accum;
  float4*        staged = nullptr;
  if (src->device != dst->device) {
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
    if (can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(LISA_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
      cudaGetLastError();
    } else {  // no peer mapping (different PCIe roots without NVLink): stage through a copy, still on the GPU
      CU(dev_alloc((void**)&staged, sizeof(float4) * (size_t)npix));
      CU(cudaMemcpyPeerAsync(staged, dst->device, src->d_accum, src->device, sizeof(float4) * (size_t)npix, dst->stream));
      from = staged;
    }
  }
  launch_accum_add(dst->d_accum, from, npix, dst->cfg.sm_count, dst->stream);
  CU(cudaStreamSynchronize(dst->stream));
  CU(cudaGetLastError());
  dev_free(staged);
  dst->stats.subframes_accumulated += src->stats.subframes_accumulated;
  dst->stats.samples += src->stats.samples;
  return LISA_OK;
}

extern "C" int lisa_accum_note_merged(lisa_ctx* c, uint32_t subframes, uint64_t samples) {
  if (!c) return fail(LISA_ERR_ARG, "null ctx");
  c->stats.subframes_accumulated += subframes;
  c->stats.samples += samples;
  return LISA_OK;
}

extern "C" void*  lisa_accum_device_ptr(lisa_ctx* c) { return c ? (void*)c->d_accum : nullptr; }
extern "C" size_t lisa_accum_bytes(lisa_ctx* c) { return c ? sizeof(float4) * (size_t)c->width * c->height : 0; }
extern "C" int    lisa_device(lisa_ctx* c) { return c ? c->device : -1; }
extern "C" int    lisa_sync(lisa_ctx* c) {
  if (!c) return fail(LISA_ERR_ARG, "null ctx");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  return LISA_OK;
}

// ---- diagnostics ---------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  ~DevBuf() { dev_free(p); }
  cudaError_t alloc(size_t n) { return dev_alloc((void**)&p, sizeof(T) * std::max<size_t>(n, 1)); }
};

extern "C" int lisa_trace_closest(lisa_ctx* c, const float* org, const float* dir, uint32_t n, float tmin, float tmax,
                                  int32_t* prim, float* t) {
  if (!c || !org || !dir || !prim) return fail(LISA_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  DevBuf<float> d_o, d_d, d_t;
  DevBuf<int>   d_p;
  CU(d_o.alloc(3ull * n)); CU(d_d.alloc(3ull * n)); CU(d_t.alloc(n)); CU(d_p.alloc(n));
  CU(cudaMemcpyAsync(d_o.p, org, sizeof(float) * 3ull * n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_d.p, dir, sizeof(float) * 3ull * n, cudaMemcpyHostToDevice, c->stream));
  launch_trace_closest(c->scene, d_o.p, d_d.p, n, tmin, tmax, d_p.p, d_t.p, c->stream);
  std::vector<int> fp(n);
  CU(cudaMemcpyAsync(fp.data(), d_p.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (t) CU(cudaMemcpyAsync(t, d_t.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  std::vector<int> f2o((size_t)c->scene.num_tris);
  if (c->scene.num_tris)
    CU(cudaMemcpyAsync(f2o.data(), c->bvh.d_final_to_orig, sizeof(int) * f2o.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(fetch_trav_overflow(overflow_slot(c), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (*overflow_slot(c)) return stack_overflow_error();
  for (uint32_t i = 0; i < n; i++) prim[i] = fp[i] >= 0 ? f2o[fp[i]] : -1;
  return LISA_OK;
}

extern "C" int lisa_trace_shadow(lisa_ctx* c, const float* org, const float* dir, uint32_t n, float tmin, float tmax,
                                 int32_t* outcome, int32_t* light) {
  if (!c || !org || !dir || !outcome) return fail(LISA_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  DevBuf<float> d_o, d_d;
  DevBuf<int>   d_oc, d_l;
  CU(d_o.alloc(3ull * n)); CU(d_d.alloc(3ull * n)); CU(d_oc.alloc(n)); CU(d_l.alloc(n));
  CU(cudaMemcpyAsync(d_o.p, org, sizeof(float) * 3ull * n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(d_d.p, dir, sizeof(float) * 3ull * n, cudaMemcpyHostToDevice, c->stream));
  launch_trace_shadow(c->scene, d_o.p, d_d.p, n, tmin, tmax, d_oc.p, d_l.p, c->stream);
  CU(cudaMemcpyAsync(outcome, d_oc.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (light) CU(cudaMemcpyAsync(light, d_l.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU(fetch_trav_overflow(overflow_slot(c), c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (*overflow_slot(c)) return stack_overflow_error();
  return LISA_OK;
}

extern "C" int lisa_primary_rays(lisa_ctx* c, uint32_t subframe, float* dirs, uint32_t* seeds_after) {
  if (!c || !dirs || !seeds_after) return fail(LISA_ERR_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  const size_t npix = (size_t)c->width * c->height;
  DevBuf<float>    d_d;
  DevBuf<uint32_t> d_s;
  CU(d_d.alloc(3 * npix)); CU(d_s.alloc(npix));
  launch_primary_rays(c->cam, subframe, d_d.p, d_s.p, c->stream);
  CU(cudaMemcpyAsync(dirs, d_d.p, sizeof(float) * 3 * npix, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(seeds_after, d_s.p, sizeof(uint32_t) * npix, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return LISA_OK;
}

extern "C" int lisa_kat_eval(int device, int what, uint32_t n, const float* in_f, const uint32_t* in_u, float* out_f,
                             uint32_t* out_u) {
  static const int fin[10] = {0, 0, 3, 2, 8, 7, 6, 3, 21, 0}, uin[10] = {2, 1, 1, 0, 0, 1, 0, 0, 0, 1};
  static const int fout[10] = {0, 3, 3, 1, 3, 3, 1, 0, 3, 3}, uout[10] = {1, 1, 1, 0, 0, 1, 0, 1, 0, 1};
  if (what < 0 || what > 9) return fail(LISA_ERR_ARG, "unknown KAT selector %d", what);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LISA_ERR_CUDA, "no CUDA device: lisa_rt has no CPU fallback");
  if (device >= 0) CU(cudaSetDevice(device));
  DevBuf<float>    d_if, d_of;
  DevBuf<uint32_t> d_iu, d_ou;
  CU(d_if.alloc((size_t)fin[what] * n)); CU(d_iu.alloc((size_t)uin[what] * n));
  CU(d_of.alloc((size_t)fout[what] * n)); CU(d_ou.alloc((size_t)uout[what] * n));
  if (fin[what] && in_f) CU(cudaMemcpy(d_if.p, in_f, sizeof(float) * fin[what] * (size_t)n, cudaMemcpyHostToDevice));
  if (uin[what] && in_u) CU(cudaMemcpy(d_iu.p, in_u, sizeof(uint32_t) * uin[what] * (size_t)n, cudaMemcpyHostToDevice));
  if (launch_kat(what, n, d_if.p, d_iu.p, d_of.p, d_ou.p, 0)) return fail(LISA_ERR_ARG, "unknown KAT selector %d", what);
  CU(cudaDeviceSynchronize());
  CU(cudaGetLastError());
  if (fout[what] && out_f) CU(cudaMemcpy(out_f, d_of.p, sizeof(float) * fout[what] * (size_t)n, cudaMemcpyDeviceToHost));
  if (uout[what] && out_u) CU(cudaMemcpy(out_u, d_ou.p, sizeof(uint32_t) * uout[what] * (size_t)n, cudaMemcpyDeviceToHost));
  return LISA_OK;
}

// ---- builder primitives, exposed for the tests --------------------------------------------------------------
extern "C" int lisa_debug_sort_pairs(int device, uint64_t* keys, uint32_t* vals, uint32_t n) {
  if ((!keys || !vals) && n) return fail(LISA_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LISA_ERR_CUDA, "no CUDA device: lisa_rt has no CPU fallback");
  if (device >= 0) CU(cudaSetDevice(device));
  DevBuf<unsigned long long> k0, k1;
  DevBuf<unsigned int>       v0, v1;
  DevBuf<char>               tmp;
  CU(k0.alloc(n)); CU(k1.alloc(n)); CU(v0.alloc(n)); CU(v1.alloc(n)); CU(tmp.alloc(radix_sort_temp_bytes(n)));
  CU(cudaMemcpy(k0.p, keys, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(v0.p, vals, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
  const int where = radix_sort_pairs(k0.p, k1.p, v0.p, v1.p, n, 0, 64, tmp.p, 0);
  CU(cudaDeviceSynchronize());
  CU(cudaGetLastError());
  CU(cudaMemcpy(keys, where ? k1.p : k0.p, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(vals, where ? v1.p : v0.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
  return LISA_OK;
}

extern "C" int lisa_debug_scan_compact(int device, uint32_t* scan_inout, int32_t* compact_inout, uint32_t n, uint32_t* total, uint32_t* kept) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LISA_ERR_CUDA, "no CUDA device: lisa_rt has no CPU fallback");
  if (device >= 0) CU(cudaSetDevice(device));
  DevBuf<uint32_t> a, tot;
  DevBuf<int>      c0, c1, cnt;
  DevBuf<char>     tmp;
  CU(a.alloc(n)); CU(tot.alloc(1)); CU(c0.alloc(n)); CU(c1.alloc(n)); CU(cnt.alloc(1));
  CU(tmp.alloc(std::max(scan_temp_bytes(n), compact_temp_bytes(n))));
  CU(cudaMemcpy(a.p, scan_inout, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c0.p, compact_inout, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
  exclusive_scan_u32(a.p, a.p, n, tot.p, tmp.p, 0);
  CU(cudaDeviceSynchronize());
  compact_nonneg(c0.p, c1.p, n, cnt.p, tmp.p, 0);
  CU(cudaDeviceSynchronize());
  CU(cudaGetLastError());
  CU(cudaMemcpy(scan_inout, a.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(total, tot.p, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  int k = 0;
  CU(cudaMemcpy(&k, cnt.p, sizeof(int), cudaMemcpyDeviceToHost));
  *kept = (uint32_t)k;
  CU(cudaMemcpy(compact_inout, c1.p, sizeof(int32_t) * (size_t)k, cudaMemcpyDeviceToHost));
  return LISA_OK;
}
