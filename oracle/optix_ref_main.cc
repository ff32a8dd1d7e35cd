/* oracle/optix_ref_main.cc — headless driver around the UNMODIFIED reference
 * (gaetanserre/LiSA) OptiX renderer.  TEST INFRASTRUCTURE, never shipped and
 * never on the product path.
 *
 * oracle/Makefile compiles the reference's own translation units where they
 * lie under /root/reference (src/LiSA/src/{optix_wrapper,scene_parser,
 * parse_obj}.cc, src/sutil/Camera.cpp) and the reference's shader.cu to PTX;
 * this file is only the glue the reference gets from its GL-dependent parts:
 *   - sutil::getInputData   (src/sutil/sutil.cpp:1001) -> read the PTX file
 *     that sits next to the executable;
 *   - render()              (src/LiSA/src/render.cc:133-148) -> same single
 *     launch of num_samples spp at subframe 0, but into a plain device frame
 *     buffer instead of a GL/zero-copy CUDAOutputBuffer, timed in ms;
 *   - display()'s progressive loop (render.cc:75-131) without the window,
 *     selected with --subframes.
 * It dumps the raw float4 accumulators (row 0 = image bottom) so parity is
 * judged on linear radiance, not on 8-bit sRGB.
 *   --soup-bin FILE --soup-material K   appends a procedural triangle soup
 *     (BASELINE C4: 1M-100M triangles; as OBJ text that is hundreds of MB and
 *     minutes in the reference's line-by-line loader) to the RendererParams the
 *     reference's SceneParser produced: FILE = u64 T, then 9T floats of
 *     vertices, 9T floats of normals (de-indexed, like parse_obj's output);
 *     every soup triangle gets material index K.  The arrays are handed to the
 *     unmodified OptixWrapper exactly as SceneParser's would be.
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <optix.h>
#include <optix_function_table_definition.h>
#include <optix_stubs.h>

#include <sutil/Exception.h>
#include <sutil/sutil.h>

#include "optix_wrapper.hh"
#include "scene_parser.hh"

static std::string g_ptx;
static std::string g_ptx_path;

namespace sutil {
const char* getInputData(const char*, const char*, const char* filename, size_t& dataSize, const char** log,
                         const std::vector<const char*>&) {
  if (log) *log = nullptr;
  std::ifstream f(g_ptx_path, std::ios::binary);
  if (!f.good()) {
    fprintf(stderr, "optix_ref: cannot read PTX %s (for %s)\n", g_ptx_path.c_str(), filename);
    exit(2);
  }
  std::stringstream ss;
  ss << f.rdbuf();
  g_ptx    = ss.str();
  dataSize = g_ptx.size();
  return g_ptx.c_str();
}
}  // namespace sutil

static const char* opt(int argc, char** argv, const char* name) {
  for (int i = 1; i + 1 < argc; i++)
    if (!strcmp(argv[i], name)) return argv[i + 1];
  return nullptr;
}
static bool flag(int argc, char** argv, const char* name) {
  for (int i = 1; i < argc; i++)
    if (!strcmp(argv[i], name)) return true;
  return false;
}

static void launch(RendererState& state, uchar4* d_frame) {
  state.params.frame_buffer = d_frame;
  CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(state.d_params), &state.params, sizeof(OptixParams),
                             cudaMemcpyHostToDevice, state.stream));
  OPTIX_CHECK(optixLaunch(state.pipeline, state.stream, reinterpret_cast<CUdeviceptr>(state.d_params),
                          sizeof(OptixParams), &state.sbt, state.params.width, state.params.height, 1));
  CUDA_SYNC_CHECK();
}

int main(int argc, char** argv) {
  const char* scene = opt(argc, argv, "-s");
  if (!scene) {
    fprintf(stderr, "usage: %s -s scene.rto [--spp S] [--subframes F] [--accum out.f32] [--ppm out.ppm] [--warmup]\n",
            argv[0]);
    return 1;
  }
  {
    std::string self(argv[0]);
    size_t      p = self.find_last_of('/');
    g_ptx_path    = (p == std::string::npos ? std::string(".") : self.substr(0, p)) + "/lisa_ref_shader.ptx";
    if (const char* e = getenv("LISA_REF_PTX")) g_ptx_path = e;
  }
  SceneParser    parser(const_cast<char*>(scene));
  RendererParams params = parser.get_params();
  std::vector<float3> soup_v, soup_n;
  std::vector<int>    soup_m;
  if (const char* sb = opt(argc, argv, "--soup-bin")) {
    const int k = opt(argc, argv, "--soup-material") ? atoi(opt(argc, argv, "--soup-material")) : 0;
    FILE* f = fopen(sb, "rb");
    unsigned long long T = 0;
    if (!f || fread(&T, sizeof(T), 1, f) != 1) { fprintf(stderr, "optix_ref: cannot read %s\n", sb); return 3; }
    const size_t nv0 = (size_t)params.num_vertices, nt0 = nv0 / 3;
    soup_v.resize(nv0 + 3 * T); soup_n.resize(nv0 + 3 * T); soup_m.assign(nv0 + 3 * T, k);  /* optix_wrapper.cc:67 copies num_vertices ints (Q11): keep that read in bounds */
    /* the soup first, the scene file's own meshes (the light) after it */
    if (fread(soup_v.data(), sizeof(float3), 3 * T, f) != 3 * T || fread(soup_n.data(), sizeof(float3), 3 * T, f) != 3 * T) {
      fprintf(stderr, "optix_ref: %s is truncated\n", sb); return 3;
    }
    fclose(f);
    memcpy(soup_v.data() + 3 * T, params.vertices, sizeof(float3) * nv0);
    memcpy(soup_n.data() + 3 * T, params.normals, sizeof(float3) * nv0);
    memcpy(soup_m.data() + T, params.mat_indices, sizeof(int) * nt0);
    params.vertices = soup_v.data(); params.normals = soup_n.data(); params.mat_indices = soup_m.data();
    params.num_vertices = (int)(nv0 + 3 * T);
  }

  auto         t_setup0 = std::chrono::steady_clock::now();
  OptixWrapper wrapper(params);
  CUDA_SYNC_CHECK();
  double setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_setup0).count();

  RendererState& state = *wrapper.get_pstate();
  const size_t   npix  = (size_t)state.params.width * state.params.height;
  uchar4*        d_frame;
  CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_frame), npix * sizeof(uchar4)));

  /* -s semantics (render.cc:140): one launch, samples_per_launch = num_samples.
   * --subframes F --spp S reproduces display()'s loop: F launches of S spp,
   * subframe_index 0..F-1, running-mean merge done by the reference shader. */
  unsigned spp       = opt(argc, argv, "--spp") ? (unsigned)atoi(opt(argc, argv, "--spp")) : params.num_samples;
  unsigned subframes = opt(argc, argv, "--subframes") ? (unsigned)atoi(opt(argc, argv, "--subframes")) : 1;
  unsigned first     = opt(argc, argv, "--first-subframe") ? (unsigned)atoi(opt(argc, argv, "--first-subframe")) : 0;

  if (flag(argc, argv, "--warmup")) {
    state.params.samples_per_launch = 1;
    state.params.subframe_index     = 0;
    launch(state, d_frame);
  }
  state.params.samples_per_launch = spp;
  /* --warmup-steps W: W untimed launches of the SAME size (bench.py's warm-up steps) */
  unsigned warm = opt(argc, argv, "--warmup-steps") ? (unsigned)atoi(opt(argc, argv, "--warmup-steps")) : 0;
  for (unsigned w = 0; w < warm; w++) {
    state.params.subframe_index = 1000000u + w;
    launch(state, d_frame);
  }
  std::vector<double> step_ms;
  auto t0                         = std::chrono::steady_clock::now();
  for (unsigned f = 0; f < subframes; f++) {
    auto ts = std::chrono::steady_clock::now();
    /* the shader merges with lerp(prev, cur, 1/(subframe_index+1)); a run that
     * starts at first>0 therefore needs prev==0 weighting handled by the
     * caller — only first==0 gives a plain mean. */
    state.params.subframe_index = first + f;
    launch(state, d_frame);
    step_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts).count());
  }
  double render_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

  std::vector<float4> accum(npix);
  std::vector<uchar4> frame(npix);
  CUDA_CHECK(cudaMemcpy(accum.data(), state.params.accum_buffer, npix * sizeof(float4), cudaMemcpyDeviceToHost));
  CUDA_CHECK(cudaMemcpy(frame.data(), d_frame, npix * sizeof(uchar4), cudaMemcpyDeviceToHost));

  if (const char* p = opt(argc, argv, "--accum")) {
    FILE* f = fopen(p, "wb");
    if (!f) { perror(p); return 3; }
    fwrite(accum.data(), sizeof(float4), npix, f);
    fclose(f);
  }
  if (const char* p = opt(argc, argv, "--ppm")) {
    /* flip + P6, as sutil::saveImage/savePPM do (sutil.cpp:523-554, 97-117) */
    FILE* f = fopen(p, "wb");
    if (!f) { perror(p); return 3; }
    fprintf(f, "P6\n%u %u\n255\n", state.params.width, state.params.height);
    for (int y = (int)state.params.height - 1; y >= 0; y--)
      for (unsigned x = 0; x < state.params.width; x++) {
        uchar4 c = frame[(size_t)y * state.params.width + x];
        fputc(c.x, f); fputc(c.y, f); fputc(c.z, f);
      }
    fclose(f);
  }
  double sum[3] = {0, 0, 0};
  for (size_t i = 0; i < npix; i++) { sum[0] += accum[i].x; sum[1] += accum[i].y; sum[2] += accum[i].z; }
  double samples = (double)npix * spp * subframes;
  printf("{\"impl\": \"optix_ref\", \"width\": %u, \"height\": %u, \"spp\": %u, \"subframes\": %u, \"bounces\": %u, "
         "\"triangles\": %d, \"setup_ms\": %.3f, \"render_ms\": %.3f, \"msamples_per_s\": %.4f, "
         "\"mean_rgb\": [%.8g, %.8g, %.8g], \"step_ms\": [",
         state.params.width, state.params.height, spp, subframes, state.params.num_bounces, params.num_vertices / 3,
         setup_ms, render_ms, samples / render_ms / 1e3, sum[0] / npix, sum[1] / npix, sum[2] / npix);
  for (size_t i = 0; i < step_ms.size(); i++) printf("%s%.3f", i ? ", " : "", step_ms[i]);
  printf("]}\n");
  fflush(stdout);
  /* skip ~OptixWrapper's teardown ordering issues on error paths: normal return runs it */
  return 0;
}
