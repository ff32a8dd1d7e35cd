/* include/lisa_rt.h — C ABI of the B200-native render path (liblisa_rt.so).
 *
 * This is the drop-in boundary for the render path of gaetanserre/LiSA: the
 * four seams the reference's own `main` crosses, flattened to
 * `extern "C"`, plain pointers and sizes.  No C++ types, no CUDA/torch types
 * in any signature.  Reference file:line are relative to the reference root.
 *
 *   B1 scene hand-off   RendererParams SceneParser::get_params()
 *                       src/LiSA/include/scene_parser.hh:13, structs.hh:90-103
 *                       -> lisa_scene_desc (field-for-field, same layouts)
 *   B2 backend lifetime OptixWrapper(const RendererParams&) / ~OptixWrapper()
 *                       src/LiSA/include/optix_wrapper.hh:3-20, optix_wrapper.cc:24-37
 *                       -> lisa_create / lisa_destroy
 *   B3 render           render() / display() / launchSubframe()
 *                       src/LiSA/src/render.cc:133-148, 75-131, 35-58
 *                       -> lisa_render_subframes, lisa_read_*, lisa_write_ppm
 *   B4 BSDF interface   bounce / BRDF / BTDF, src/LiSA/src/bsdfs/lambertian.cu:7-27
 *                       -> device header lisa_b200/csrc/bsdf/lambertian.cuh
 *                          (compile-time seam, as in the reference)
 *
 * Error convention: every int-returning function returns LISA_OK (0) or a
 * negative lisa_status; lisa_last_error() gives the message of the calling
 * thread's last failure.  The reference throws sutil::Exception instead
 * (src/sutil/Exception.h:167-177); the host front-end (lisa_b200/host)
 * converts a failure back into the reference's abort-with-message behaviour.
 * There is NO CPU fallback: without a CUDA device lisa_create fails with
 * LISA_ERR_CUDA.
 *
 * Threading: one lisa_ctx is used from one host thread at a time.  All device
 * work of a context runs on the context's own CUDA stream on its device.
 */
#ifndef LISA_RT_H
#define LISA_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LISA_RT_VERSION 1

typedef enum lisa_status {
  LISA_OK            = 0,
  LISA_ERR_ARG       = -1, /* null pointer, inconsistent sizes, index out of range */
  LISA_ERR_CUDA      = -2, /* CUDA runtime failure (message holds cudaGetErrorString) */
  LISA_ERR_NOMEM     = -3,
  LISA_ERR_IO        = -4,
  LISA_ERR_STATE     = -5
} lisa_status;

/* src/LiSA/include/structs.hh:16-23 — identical 40-byte layout. alpha < 1 marks a
 * dielectric whose parameter is the IOR `n`; otherwise `roughness` is used
 * (structs.hh:41-52).  Fields the reference leaves uninitialised must be zero. */
typedef struct lisa_material {
  float   roughness;
  float   alpha;
  float   n;
  float   diffuse_color[3];
  uint8_t emit;
  uint8_t _pad[3];
  float   emission_color[3];
} lisa_material;

/* src/LiSA/include/structs.hh:54-58 */
typedef struct lisa_camera {
  float eye[3];
  float look_at[3];
  float fov; /* vertical, degrees */
} lisa_camera;

/* src/LiSA/include/structs.hh:90-103 (RendererParams).  float3 = 3 packed floats.
 * The arrays are only read during lisa_create (copy semantics). */
typedef struct lisa_scene_desc {
  const float*         vertices;     /* num_vertices * 3 floats; triangle t = vertices 3t, 3t+1, 3t+2 */
  const float*         normals;      /* num_vertices * 3 floats */
  const lisa_material* materials;    /* num_materials */
  const int32_t*       mat_indices;  /* ONE per triangle (num_vertices / 3); the reference over-reads this (Q11) */
  int32_t              num_vertices; /* 3 * triangles */
  int32_t              num_materials;
  uint32_t             width, height;
  lisa_camera          camera;
  uint32_t             num_samples;
  uint32_t             num_bounces;
  const char*          output_image; /* may be NULL; only lisa_write_ppm(ctx, NULL) uses it */
} lisa_scene_desc;

enum { LISA_SHADOW_CLOSEST = 0, LISA_SHADOW_FIRST_FOUND = 1 };
enum { LISA_BVH_WIDE8 = 0, LISA_BVH_BINARY = 1 };
enum { LISA_FLAG_PROFILE_STAGES = 1, /* CUDA events around every stage launch */
       LISA_FLAG_LBVH = 2,           /* plain LBVH hierarchy (fastest build) instead of PLOC clustering */
       LISA_FLAG_NO_CULL = 4,        /* traverse EVERY shadow try.  By default a try whose outcome provably cannot change
                                        RayState::hit (hit is false and the ray cannot reach any emitter: outside the cone
                                        around the emitter bounds) is resolved without traversal; images are bit-identical
                                        either way and lisa_stats reports how many tries were resolved that way. */
       LISA_FLAG_WAVEFRONT = 8,      /* render with the wavefront pipeline (three kernels per bounce over chain state in HBM)
                                        instead of the default single persistent kernel per tile; bit-identical images */
       LISA_FLAG_SPLIT_TRIANGLES = 16 /* build the BVH over REFERENCES: a triangle longer than about twice the scene's mean
                                        triangle spacing is handed to the builder once per cell of a grid over its bounding box,
                                        each time with the box of the part inside that cell (early split clipping; at most 1.5
                                        references per triangle on average).  Tighter boxes for long thin triangles; the hits,
                                        hence the closest-hit images, are those of the unsplit soup.  lisa_stats.num_references
                                        reports the count.  Off by default. */ };

typedef struct lisa_options {
  uint32_t struct_size;   /* sizeof(lisa_options) */
  int32_t  device;        /* CUDA device ordinal; -1 = current device */
  uint32_t shadow_mode;   /* LISA_SHADOW_CLOSEST (default): the closest hit of a shadow ray decides.
                             LISA_SHADOW_FIRST_FOUND: the first hit found in traversal order decides, which is
                             what the reference asks OptiX for (shader.cu:69) and is traversal-order dependent. */
  uint32_t bvh_kind;      /* LISA_BVH_WIDE8 (default) compressed 8-wide; LISA_BVH_BINARY for ablation */
  uint32_t max_chains;    /* upper bound on concurrently resident (pixel, subframe) sample chains; 0 = auto */
  uint32_t flags;         /* LISA_FLAG_* */
} lisa_options;

typedef struct lisa_stats {
  uint32_t struct_size;
  uint32_t num_triangles, num_emitter_triangles;
  uint32_t bvh_nodes, bvh_emitter_nodes; /* nodes of the traversal BVH(s) */
  uint64_t bvh_bytes, triangle_bytes;    /* device bytes of nodes / of packed triangles (intersection + shading) */
  float    upload_ms, bvh_build_ms;      /* lisa_create: H2D copy; device BVH build (CUDA events) */
  /* cumulative since lisa_create / lisa_reset_accum */
  uint64_t samples, radiance_rays, shadow_rays, null_directions;
  uint64_t kernel_launches, iterations;
  double   render_ms;                    /* sum over lisa_render_subframes calls, CUDA events on the ctx stream */
  /* last lisa_render_subframes call only */
  double   last_render_ms;
  uint64_t last_samples, last_radiance_rays, last_shadow_rays, last_kernel_launches;
  double   last_extend_ms, last_shadow_ms; /* per-stage device time (CUDA events around every launch of the stage),
                                              only when options.flags & LISA_FLAG_PROFILE_STAGES or LISA_PROFILE_STAGES=1 */
  uint64_t state_bytes;                  /* device bytes of chain state + queues */
  uint32_t subframes_accumulated;
  uint32_t num_references;               /* primitives of the BVH: num_triangles unless LISA_FLAG_SPLIT_TRIANGLES */
  uint64_t last_extend_launches, last_shadow_launches, last_shadow_jobs; /* jobs = opaque hits light-sampled */
  uint64_t nodes_visited, triangles_tested;           /* cumulative traversal work (both stages) */
  uint64_t last_nodes_visited, last_triangles_tested;
  uint64_t shadow_culled, last_shadow_culled;         /* shadow tries resolved without traversal (counted in shadow_rays too) */
  /* stages of the device BVH build (CUDA events; they add up to bvh_build_ms): Morton keys + radix sort, hierarchy (PLOC rounds
   * or LBVH, optional rotations), collapse to 8-wide nodes, triangle packing */
  float    build_sort_ms, build_hierarchy_ms, build_collapse_ms, build_pack_ms;
  /* the builder's surface-area estimate of the node visits of a ray that crosses the scene (sum of the 8-wide nodes' surface
   * areas over their root's; 0 for the binary BVH), and the flavour of the persistent kernel chosen from it: 0 = shallow
   * (64 chains per warp), 1 = deep (48 chains per warp and a 12-entry traversal stack in shared memory: scenes whose rays
   * visit tens of nodes).  After a render call that traced >= 100,000 rays the flavour follows the node visits per ray that
   * call measured (deep above 10, shallow below 8) unless LISA_POOL_FLAVOUR forces one.  The flavours give bit-identical images. */
  float    bvh_sah_nodes_per_ray;
  uint32_t pool_flavour;
} lisa_stats;

typedef struct lisa_ctx lisa_ctx;

/* B2.  Uploads the soup, builds the BVH on the device, derives the camera frame
 * (src/sutil/Camera.cpp:34-45 with up = (0,1,0), aspect = width/height,
 * optix_wrapper.cc:433-442) and allocates the accumulators. */
int  lisa_create(const lisa_scene_desc* scene, const lisa_options* options /* may be NULL */, lisa_ctx** out);
void lisa_destroy(lisa_ctx* ctx);
const char* lisa_last_error(void);
int  lisa_version(void);

/* B3.  Renders subframes first .. first+count-1, each with spp samples per pixel, seeded
 * tea<16>(pixel, subframe) exactly as shader.cu:141, and adds their samples to the accumulators (sum of
 * samples | number of samples per pixel): the image is the mean over ALL samples rendered since the last reset, which is
 * what the reference's running mean of equally sized subframes, shader.cu:160-164, computes.
 *   reference `-s`  ==  lisa_render_subframes(ctx, 0, 1, num_samples)            (render.cc:140)
 *   reference `-d`  ==  for f = 0.. : lisa_render_subframes(ctx, f, 1, min(16, num_samples))
 * Blocking: returns when the accumulators on the device are final. */
int lisa_render_subframes(lisa_ctx* ctx, uint32_t first_subframe, uint32_t count, uint32_t spp_per_subframe);
int lisa_reset_accum(lisa_ctx* ctx);

/* Linear mean radiance, W*H*4 floats (r, g, b, 1), row 0 = image BOTTOM like the reference's accum_buffer. */
int lisa_read_accum(lisa_ctx* ctx, float* rgba);
/* sRGB-quantised frame exactly as make_color (src/cuda/helpers.h:129-138), W*H*4 bytes, row 0 = bottom. */
int lisa_read_rgba8(lisa_ctx* ctx, uint8_t* rgba);
/* Vertically flipped binary PPM as sutil::saveImage/savePPM (src/sutil/sutil.cpp:523-554, 97-117).
 * path == NULL uses scene->output_image. */
int lisa_write_ppm(lisa_ctx* ctx, const char* path);
/* save_image() -> sutil::saveImage (src/LiSA/src/render.cc:9-17, src/sutil/sutil.cpp:523-688) for the uchar4 frame the
 * reference hands it: the last three characters of the name pick the format — "ppm"/"PPM" the flipped P6 above,
 * "png"/"PNG" an 8-bit RGBA PNG (flipped, alpha 255), "exr"/"EXR" and anything else fail with the reference's message
 * (LISA_ERR_ARG), as does a name shorter than 5 characters.  path == NULL uses scene->output_image. */
int lisa_write_image(lisa_ctx* ctx, const char* path);
/* Linear float image (Portable Float Map, RGB, little endian): the accumulators without the 8-bit sRGB quantisation,
 * for parity tooling (the reference can only write the quantised PPM; README.md:183 lists float output as a TODO). */
int lisa_write_pfm(lisa_ctx* ctx, const char* path);
/* Checkpoint / resume of a progressive render (the reference keeps its accum_buffer only in device
 * memory, src/LiSA/src/render.cc:75-131).  lisa_save_accum writes the raw accumulators (W*H float4: sum of
 * all samples | number of samples) with a header naming the image size and the number of subframes accumulated;
 * lisa_load_accum replaces the accumulators of a context of the same size with a saved file and reports through
 * *subframes how many subframes it holds, so that rendering resumes at that subframe index: the resumed image is
 * bit-identical to an uninterrupted one. */
int lisa_save_accum(lisa_ctx* ctx, const char* path);
int lisa_load_accum(lisa_ctx* ctx, const char* path, uint32_t* subframes);
int lisa_get_stats(lisa_ctx* ctx, lisa_stats* stats /* struct_size set by caller */);

/* Serialisable BVH (the reference rebuilds its GAS on every start, optix_wrapper.cc:105-145).
 * lisa_save_bvh writes what lisa_create built: the node array, the packed triangles (positions + material ids, normals +
 * original indices) in leaf order, the roots and bounds of the emitter / non-emitter partitions.  lisa_create_from_bvh makes
 * a context from such a file: no soup upload, no build — scene->vertices / normals / mat_indices may be NULL (num_vertices 0),
 * everything else of the scene (materials, camera, size) is used; the file is refused if it is truncated, if the triangle
 * count disagrees with a scene that does carry geometry, or if the materials' emitter flags differ from those the BVH was
 * built for (the partition is baked in).  Images are bit-identical to those of the context that saved the file. */
int lisa_save_bvh(lisa_ctx* ctx, const char* path);
int lisa_create_from_bvh(const lisa_scene_desc* scene, const lisa_options* options /* may be NULL */, const char* path, lisa_ctx** out);

/* Multi-GPU plumbing: every rank renders a disjoint subframe set into its own sums; the
 * caller reduces the sum buffers (one NCCL reduce) and the root reads the image.  The buffer holds W*H
 * float4 = (sum of all samples .xyz, number of samples .w; the image is .xyz / .w) and lives on the context's device. */
/* Same-process alternative to the NCCL reduce: dst's accumulators += src's, as ONE kernel on dst's device that reads
 * src's buffer over NVLink through the peer mapping (falls back to a peer copy where P2P is unavailable). */
int    lisa_accum_add_peer(lisa_ctx* dst, lisa_ctx* src);
/* Bookkeeping after the CALLER summed other contexts' accumulators into ctx's buffer (e.g. one ncclReduce onto
 * lisa_accum_device_ptr): adds `subframes` / `samples` to ctx's counters, so that lisa_get_stats and the header written by
 * lisa_save_accum describe what the buffer now holds. */
int    lisa_accum_note_merged(lisa_ctx* ctx, uint32_t subframes, uint64_t samples);
void*  lisa_accum_device_ptr(lisa_ctx* ctx);
size_t lisa_accum_bytes(lisa_ctx* ctx);
int    lisa_device(lisa_ctx* ctx);
int    lisa_sync(lisa_ctx* ctx);

/* B3 on several GPUs of one box, ONE process (SURVEY.md §8b B3 "internally splits subframes over GPUs and reduces", §8e):
 * a lisa_multi owns one context per GPU (replicated scene, each GPU builds its own BVH from the same upload) and one NCCL
 * communicator per GPU (ncclCommInitAll).  lisa_multi_render_subframes partitions subframes [first, first+count) into
 * contiguous blocks, one per GPU (block sizes differ by at most one), renders them concurrently (one host thread per GPU)
 * and combines the float4 accumulators with ONE ncclReduce(sum, fp32, root = GPU 0) of W*H*4 floats over NVLink inside an
 * ncclGroupStart/End; afterwards GPU 0's accumulators hold everything rendered so far and the others are cleared.
 * lisa_multi_render_samples is the `-s` call for G GPUs: num_samples split into G subframes of floor/ceil(N/G) spp (exactly
 * N samples in total — the accumulators weigh by sample count).  Read the image through lisa_multi_root(...) with the
 * single-GPU calls.  NCCL is loaded at run time (libnccl.so.2); where it is missing the reduce falls back to
 * lisa_accum_add_peer (one kernel per peer reading over NVLink) and lisa_multi_backend() says "peer" instead of "nccl".
 * num_gpus <= 0 means every visible device.  Deterministic: the N-GPU image equals the 1-GPU image over the same subframes
 * up to the order of the float additions. */
typedef struct lisa_multi lisa_multi;
int         lisa_multi_create(const lisa_scene_desc* scene, const lisa_options* options /* device is ignored */, int num_gpus, lisa_multi** out);
void        lisa_multi_destroy(lisa_multi* m);
int         lisa_multi_num_gpus(const lisa_multi* m);
lisa_ctx*   lisa_multi_root(lisa_multi* m);
lisa_ctx*   lisa_multi_ctx(lisa_multi* m, int gpu);
const char* lisa_multi_backend(const lisa_multi* m);
int         lisa_multi_reset_accum(lisa_multi* m);
int         lisa_multi_render_subframes(lisa_multi* m, uint32_t first_subframe, uint32_t count, uint32_t spp_per_subframe);
int         lisa_multi_render_samples(lisa_multi* m, uint32_t first_subframe, uint32_t num_samples);
/* last lisa_multi_render_* call: wall time of the render phase (slowest GPU) and of the reduce, in ms */
int         lisa_multi_last_times(const lisa_multi* m, double* render_ms, double* reduce_ms);

/* Diagnostics used by the parity tests (not part of the reference surface).
 * Batched queries against the context's BVH; host arrays; n rays.  org/dir: n*3 floats.
 * prim: index of the hit triangle IN THE CALLER'S ORDER (mat_indices order) or -1; t, u, v optional. */
int lisa_trace_closest(lisa_ctx* ctx, const float* org, const float* dir, uint32_t n, float tmin, float tmax,
                       int32_t* prim, float* t /* may be NULL */);
/* outcome per ray: 0 miss, 1 emitter decides (light = material index), 2 non-emitter decides. */
int lisa_trace_shadow(lisa_ctx* ctx, const float* org, const float* dir, uint32_t n, float tmin, float tmax,
                      int32_t* outcome, int32_t* light /* may be NULL */);
/* First camera ray of every pixel of `subframe` as raygen builds it (shader.cu:141-152):
 * dirs W*H*3 floats, seeds_after W*H (state after the two jitter draws). */
int lisa_primary_rays(lisa_ctx* ctx, uint32_t subframe, float* dirs, uint32_t* seeds_after);
/* Evaluates the device helpers on the GPU for known-answer tests.  `what`:
 *   0 tea16(in_u[2i], in_u[2i+1])                        -> out_u[i]
 *   1 rnd x3 from seed in_u[i]                            -> out_f[3i..], out_u[i] = seed after
 *   2 hemisphere(N = in_f[3i..], seed in_u[i])            -> out_f[3i..], out_u[i]
 *   3 BTDF/fresnel(cos = in_f[2i], eta = in_f[2i+1])      -> out_f[i]
 *   4 refract(cosI, dir, N, eta = in_f[8i..8i+7])         -> out_f[3i..]
 *   5 bounce(dir, N, roughness = in_f[7i..], seed in_u[i])-> out_f[3i..], out_u[i]
 *   6 BRDF(N, L = in_f[6i..])                             -> out_f[i]
 *   7 make_color(rgb = in_f[3i..])                        -> out_u[i] = r | g<<8 | b<<16 | a<<24
 *   8 shading normal(P, n1,n2,n3, v1,v2,v3 = in_f[21i..]) -> out_f[3i..]
 *   9 rng x3 (conversion-free form) from seed in_u[i]     -> out_f[3i..], out_u[i] = seed after
 * Returns LISA_ERR_ARG for an unknown selector. */
int lisa_kat_eval(int device, int what, uint32_t n, const float* in_f, const uint32_t* in_u, float* out_f,
                  uint32_t* out_u);

/* The builder's device primitives (own LSD radix sort of (u64 key, u32 value) pairs, exclusive scan, stream
 * compaction of non-negative ints), run on host arrays in place — for the tests only. */
int lisa_debug_sort_pairs(int device, uint64_t* keys, uint32_t* vals, uint32_t n);
int lisa_debug_scan_compact(int device, uint32_t* scan_inout, int32_t* compact_inout, uint32_t n, uint32_t* total, uint32_t* kept);

#ifdef __cplusplus
}
#endif
#endif /* LISA_RT_H */
