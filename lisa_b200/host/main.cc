// lisa_b200/host/main.cc — the CLI (src/LiSA/src/main.cc:6-28, include/parse_args.hh:3-16).
//   lisa -s scene.rto [-d]
// -s <scene> is mandatory; without it the usage goes to stderr and the exit code is 1.  -d selects the
// progressive mode (headless here).  Extra, optional: --gpus N splits num_samples over N GPUs of this box (one context each, in
// this process) and combines the accumulators with one ncclReduce (lisa_multi_*); --obj-cache keeps a binary copy of each OBJ's triangle soup next to
// it (<file>.lisasoup) and loads that instead of the text when it is current; --pfm <file> also writes the linear float image; --save-bvh <file> serialises the BVH after the build and
// --load-bvh <file> starts from such a file instead of loading the OBJ meshes and building (lisa_save_bvh / lisa_create_from_bvh);
// --split-triangles builds the BVH over references to long thin triangles (LISA_FLAG_SPLIT_TRIANGLES); with -d,
// --snapshot-every K rewrites the PPM every K subframes, --checkpoint <file> saves the accumulators then (and at the
// end) and --resume <file> continues an interrupted render from such a file; --stats prints one JSON line with the counters of
// include/lisa_rt.h:lisa_stats; environment variables LISA_BVH/LISA_SHADOW/LISA_MAX_CHAINS select
// ablation variants (see lisa_rt.cu).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "parse_obj.hh"
#include "render.hh"
#include "scene_parser.hh"

static char* getCmdOption(char** begin, char** end, const std::string& option) {
  char** itr = std::find(begin, end, option);
  if (itr != end && ++itr != end) return *itr;
  return nullptr;
}
static bool cmdOptionExists(char** begin, char** end, const std::string& option) { return std::find(begin, end, option) != end; }

int main(int argc, char** argv) {
  char* scene_path = nullptr;
  if (cmdOptionExists(argv, argv + argc, "-s")) scene_path = getCmdOption(argv, argv + argc, "-s");
  if (!scene_path) {
    std::cerr << "Missing scene path." << std::endl;
    std::cerr << "Usage: " << argv[0] << " -s scene_path" << std::endl;
    return 1;
  }
  try {
    if (cmdOptionExists(argv, argv + argc, "--obj-cache")) parse_obj_set_cache(1);
    char* load_bvh = getCmdOption(argv, argv + argc, "--load-bvh");
    SceneParser     parser(scene_path, load_bvh == nullptr);  // with a serialised BVH the OBJ files are not even opened
    lisa_scene_desc params = parser.get_params();
    if (char* g = getCmdOption(argv, argv + argc, "--gpus")) {  // sample-space partition over the GPUs of this box
      if (atoi(g) > 1) {
        printf("Starting rendering...\n");
        render_multi(params, atoi(g));
        return 0;
      }
    }
    lisa_ctx*       ctx = nullptr;
    lisa_options    opt{};
    opt.struct_size = sizeof(opt);
    opt.device = -1;
    if (cmdOptionExists(argv, argv + argc, "--split-triangles")) opt.flags |= LISA_FLAG_SPLIT_TRIANGLES;
    if ((load_bvh ? lisa_create_from_bvh(&params, &opt, load_bvh, &ctx) : lisa_create(&params, &opt, &ctx)) != LISA_OK) {
      // the reference throws sutil::Exception out of OptixWrapper's constructor and aborts
      std::cerr << "lisa: " << lisa_last_error() << std::endl;
      return 134;
    }
    if (char* save_bvh = getCmdOption(argv, argv + argc, "--save-bvh")) {
      if (lisa_save_bvh(ctx, save_bvh) != LISA_OK) std::cerr << "lisa: " << lisa_last_error() << std::endl;
    }
    printf("Starting rendering...\n");
    if (cmdOptionExists(argv, argv + argc, "-d")) {
      DisplayOptions dopt;
      if (char* e = getCmdOption(argv, argv + argc, "--snapshot-every")) dopt.snapshot_every = (unsigned)std::max(0, atoi(e));
      dopt.checkpoint = getCmdOption(argv, argv + argc, "--checkpoint");
      dopt.resume = getCmdOption(argv, argv + argc, "--resume");
      display(ctx, params, dopt);
    } else render(ctx, params);
    if (char* pfm = getCmdOption(argv, argv + argc, "--pfm")) {
      if (lisa_write_pfm(ctx, pfm) != LISA_OK) std::cerr << "lisa: " << lisa_last_error() << std::endl;
    }
    if (cmdOptionExists(argv, argv + argc, "--stats")) {
      lisa_stats s;
      s.struct_size = sizeof(s);
      lisa_get_stats(ctx, &s);
      printf("{\"triangles\": %u, \"references\": %u, \"bvh_nodes\": %u, \"bvh_sah_nodes_per_ray\": %.2f, \"kernel_flavour\": \"%s\", \"upload_ms\": %.3f, \"bvh_build_ms\": %.3f, \"bvh_build_stages_ms\": {\"morton_sort\": %.3f, "
             "\"hierarchy\": %.3f, \"collapse\": %.3f, \"pack\": %.3f}, \"render_ms\": %.3f, \"samples\": %llu, "
             "\"radiance_rays\": %llu, \"shadow_rays\": %llu, \"msamples_per_s\": %.3f, \"mrays_per_s\": %.3f, "
             "\"kernel_launches\": %llu}\n",
             s.num_triangles, s.num_references, s.bvh_nodes, s.bvh_sah_nodes_per_ray, s.pool_flavour ? "deep" : "shallow", s.upload_ms, s.bvh_build_ms, s.build_sort_ms, s.build_hierarchy_ms, s.build_collapse_ms, s.build_pack_ms,
             s.render_ms, (unsigned long long)s.samples,
             (unsigned long long)s.radiance_rays, (unsigned long long)s.shadow_rays, s.samples / s.render_ms / 1e3,
             (s.radiance_rays + s.shadow_rays) / s.render_ms / 1e3, (unsigned long long)s.kernel_launches);
    }
    lisa_destroy(ctx);
  } catch (const SceneError& e) {
    std::string m = e.what();
    std::cerr << m;
    if (m.empty() || m.back() != '\n') std::cerr << std::endl;
    return e.exit_code & 0xff;
  } catch (const std::exception& e) {
    std::cerr << "lisa: " << e.what() << std::endl;
    return 134;
  }
  return 0;
}
