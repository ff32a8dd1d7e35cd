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
  const unsigned spp = (params.num_samples + ngpus - 1) / ngpus;
  std::vector<lisa_ctx*>   ctx(ngpus, nullptr);
  std::vector<std::string> err(ngpus);
  auto start = std::chrono::system_clock::now();
  std::vector<std::thread> th;
  for (int g = 0; g < ngpus; g++)
    th.emplace_back([&, g] {
      lisa_options o{};
      o.struct_size = sizeof(o);
      o.device = g;
      if (lisa_create(&params, &o, &ctx[g]) != LISA_OK) { err[g] = lisa_last_error(); return; }
      if (lisa_render_subframes(ctx[g], (uint32_t)g, 1, spp) != LISA_OK) err[g] = lisa_last_error();
    });
  for (auto& t : th) t.join();
  for (int g = 0; g < ngpus; g++)
    if (!err[g].empty()) {
      for (lisa_ctx* c : ctx) lisa_destroy(c);
      throw std::runtime_error("GPU " + std::to_string(g) + ": " + err[g]);
    }
  for (int g = 1; g < ngpus; g++) check(lisa_accum_add_peer(ctx[0], ctx[g]), "reduce");
  save_image(ctx[0], params);
  std::chrono::duration<double> total = std::chrono::system_clock::now() - start;
  for (lisa_ctx* c : ctx) lisa_destroy(c);
  printf("Rendering finished in %.2f mn.\n", total.count() / 60.0);
  return total.count();
}
