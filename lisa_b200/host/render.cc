// lisa_b200/host/render.cc — see render.hh.
#include "render.hh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>

static void check(int rc, const char* what) {
  if (rc != LISA_OK) throw std::runtime_error(std::string(what) + ": " + lisa_last_error());
}

static void save_image(lisa_ctx* ctx, const lisa_scene_desc& params) {  // render.cc:9-17
  check(lisa_write_ppm(ctx, params.output_image), "save_image");
}

void render(lisa_ctx* ctx, const lisa_scene_desc& params) {
  auto start = std::chrono::system_clock::now();
  check(lisa_reset_accum(ctx), "reset");
  check(lisa_render_subframes(ctx, 0, 1, params.num_samples), "render");  // samples_per_launch = num_samples (render.cc:140)
  save_image(ctx, params);
  std::chrono::duration<float> total = std::chrono::system_clock::now() - start;
  printf("Rendering finished in %.2f mn.\n", total.count() / 60.0f);
}

void display(lisa_ctx* ctx, const lisa_scene_desc& params) {
  const unsigned spl = params.num_samples > 16 ? 16 : params.num_samples;  // optix_wrapper.cc:430
  auto start = std::chrono::system_clock::now();
  check(lisa_reset_accum(ctx), "reset");
  unsigned subframe = 0;
  double   render_s = 0;
  do {
    auto t0 = std::chrono::steady_clock::now();
    check(lisa_render_subframes(ctx, subframe, 1, spl), "render");
    std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t0;
    render_s += dt.count();
    subframe++;
    // stands in for sutil::displayStats + the "nb sample" overlay
    printf("render %8.2f ms | nb sample   : %8u\n", dt.count() * 1e3, subframe * spl);
    fflush(stdout);
  } while ((unsigned long long)subframe * spl < params.num_samples);  // render.cc:121
  save_image(ctx, params);
  std::chrono::duration<float> total = std::chrono::system_clock::now() - start;
  printf("Rendering finished in %.2f mn.\n", total.count() / 60.0f);
}
