// lisa_b200/host/render.cc — see render.hh.
#include "render.hh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

static void check(int rc, const char* what) {
  if (rc != LISA_OK) throw std::runtime_error(std::string(what) + ": " + lisa_last_error());
}

static void save_image(lisa_ctx* ctx, const lisa_scene_desc& params) {  // render.cc:9-17: sutil::saveImage picks the format by extension
  check(lisa_write_image(ctx, params.output_image), "save_image");
}

void render(lisa_ctx* ctx, const lisa_scene_desc& params) {
  auto start = std::chrono::system_clock::now();
  check(lisa_reset_accum(ctx), "reset");
  check(lisa_render_subframes(ctx, 0, 1, params.num_samples), "render");  // samples_per_launch = num_samples (render.cc:140)
  save_image(ctx, params);
  std::chrono::duration<float> total = std::chrono::system_clock::now() - start;
  printf("Rendering finished in %.2f mn.\n", total.count() / 60.0f);
}

void display(lisa_ctx* ctx, const lisa_scene_desc& params) { display(ctx, params, DisplayOptions()); }

void display(lisa_ctx* ctx, const lisa_scene_desc& params, const DisplayOptions& opt) {
  const unsigned spl = params.num_samples > 16 ? 16 : params.num_samples;  // optix_wrapper.cc:430
  auto start = std::chrono::system_clock::now();
  check(lisa_reset_accum(ctx), "reset");
  unsigned subframe = 0;
  double   render_s = 0;
  if (opt.resume) {
    check(lisa_load_accum(ctx, opt.resume, &subframe), "resume");
    printf("resumed %s at subframe %u (%llu samples)\n", opt.resume, subframe, (unsigned long long)subframe * spl);
    if ((unsigned long long)subframe * spl >= params.num_samples) {  // the checkpoint already holds the whole render
      save_image(ctx, params);
      printf("Rendering finished in %.2f mn.\n", 0.0f);
      return;
    }
  }
  do {
    auto t0 = std::chrono::steady_clock::now();
    check(lisa_render_subframes(ctx, subframe, 1, spl), "render");
    std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t0;
    render_s += dt.count();
    subframe++;
    // stands in for sutil::displayStats + the "nb sample" overlay
    printf("render %8.2f ms | nb sample   : %8u\n", dt.count() * 1e3, subframe * spl);
    fflush(stdout);
    if (opt.snapshot_every && subframe % opt.snapshot_every == 0 && (unsigned long long)subframe * spl < params.num_samples) {
      save_image(ctx, params);
      if (opt.checkpoint) check(lisa_save_accum(ctx, opt.checkpoint), "checkpoint");
    }
  } while ((unsigned long long)subframe * spl < params.num_samples);  // render.cc:121
  save_image(ctx, params);
  if (opt.checkpoint) check(lisa_save_accum(ctx, opt.checkpoint), "checkpoint");
  std::chrono::duration<float> total = std::chrono::system_clock::now() - start;
  printf("Rendering finished in %.2f mn.\n", total.count() / 60.0f);
}

double render_multi(const lisa_scene_desc& params, int ngpus) {
  if (ngpus < 1) ngpus = 1;
  lisa_multi* m = nullptr;
  auto start = std::chrono::system_clock::now();
  check(lisa_multi_create(&params, nullptr, ngpus, &m), "create");
  // num_samples split into one subframe per GPU of floor/ceil(N / G) spp — exactly N samples — then ONE ncclReduce
  int rc = lisa_multi_render_samples(m, 0, params.num_samples);
  if (rc == LISA_OK) rc = lisa_write_image(lisa_multi_root(m), params.output_image);
  const std::string err = rc == LISA_OK ? "" : lisa_last_error();
  double render_ms = 0, reduce_ms = 0;
  lisa_multi_last_times(m, &render_ms, &reduce_ms);
  const std::string backend = lisa_multi_backend(m);
  lisa_multi_destroy(m);
  if (rc != LISA_OK) throw std::runtime_error("render_multi: " + err);
  std::chrono::duration<double> total = std::chrono::system_clock::now() - start;
  printf("%d GPUs: render %.2f ms, reduce (%s) %.2f ms\n", ngpus, render_ms, backend.c_str(), reduce_ms);
  printf("Rendering finished in %.2f mn.\n", total.count() / 60.0);
  return total.count();
}
