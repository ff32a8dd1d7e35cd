// integration/lisa_main_dropin.cc — the reference-side binding of INTEGRATION.md §A, as a file that compiles.
//
// This is gaetanserre/LiSA's own main (src/LiSA/src/main.cc:6-28) with the three lines that reach OptiX
//     OptixWrapper wrapper(params);  display(*wrapper.get_pstate(), params);  render(*wrapper.get_pstate(), params);
// replaced by calls into liblisa_rt.so (include/lisa_rt.h).  Everything in front of the boundary is the REFERENCE'S code:
// oracle/Makefile (target _ref/lisa_dropin) compiles the reference's scene_parser.cc + parse_obj.cc and its headers
// (scene_parser.hh, structs.hh, parse_args.hh) where they lie under /root/reference, and links this file against
// -llisa_rt — no OptiX runtime, no GL, no sutil, no PTX directory.  tests/test_dropin.py renders a scene with it and
// requires the PPM to equal the one the product's own CLI (lisa_b200/lisa) writes, byte for byte.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "scene_parser.hh"  // reference: SceneParser, RendererParams (structs.hh:90-103)
#include "parse_args.hh"    // reference: getCmdOption / cmdOptionExists (needs scene_parser.hh's `using namespace std`, as in main.cc:2-4)

#include "lisa_rt.h"

// B1: RendererParams -> lisa_scene_desc, field for field (structs.hh:90-103; Material keeps its 40-byte layout,
// structs.hh:16-23; float3 = three packed floats)
static lisa_scene_desc to_desc(const RendererParams& p) {
  static_assert(sizeof(Material) == sizeof(lisa_material), "40-byte Material");
  static_assert(sizeof(float3) == 12, "packed float3");
  lisa_scene_desc d{};
  d.vertices      = reinterpret_cast<const float*>(p.vertices);
  d.normals       = reinterpret_cast<const float*>(p.normals);
  d.materials     = reinterpret_cast<const lisa_material*>(p.materials);
  d.mat_indices   = p.mat_indices;  // one per triangle (the reference over-reads this array, Q11; the library does not)
  d.num_vertices  = p.num_vertices;
  d.num_materials = p.num_materials;
  d.width         = p.width;
  d.height        = p.height;
  memcpy(d.camera.eye, &p.camera.eye, 12);
  memcpy(d.camera.look_at, &p.camera.look_at, 12);
  d.camera.fov    = p.camera.fov;
  d.num_samples   = p.num_samples;
  d.num_bounces   = p.num_bounces;
  d.output_image  = p.output_image;
  return d;
}

static void check(int rc) {
  if (rc != LISA_OK) throw std::runtime_error(lisa_last_error());  // sutil::Exception in the reference
}

int main(int argc, char** argv) {
  char* scene_path;
  if (cmdOptionExists(argv, argv + argc, "-s")) {
    scene_path = getCmdOption(argv, argv + argc, "-s");
  } else {
    std::cerr << "Missing scene path." << std::endl;
    std::cerr << "Usage: " << argv[0] << " -s scene_path" << std::endl;
    exit(1);
  }

  SceneParser    parser(scene_path);
  RendererParams params = parser.get_params();

  lisa_scene_desc desc = to_desc(params);
  lisa_ctx*       ctx  = nullptr;
  check(lisa_create(&desc, nullptr, &ctx));  // B2: OptixWrapper wrapper(params);

  printf("Starting rendering...\n");
  if (cmdOptionExists(argv, argv + argc, "-d")) {  // B3: display(), render.cc:75-131 — subframes of min(16, N) spp
    const unsigned spl = params.num_samples > 16 ? 16 : params.num_samples;  // optix_wrapper.cc:430
    unsigned       f   = 0;
    do { check(lisa_render_subframes(ctx, f++, 1, spl)); } while ((unsigned long long)f * spl < params.num_samples);
  } else {  // B3: render(), render.cc:133-148 — one launch of num_samples spp at subframe 0
    check(lisa_render_subframes(ctx, 0, 1, params.num_samples));
  }
  check(lisa_write_image(ctx, params.output_image));  // save_image(), render.cc:9-17
  lisa_destroy(ctx);                                  // ~OptixWrapper
  return 0;
}
