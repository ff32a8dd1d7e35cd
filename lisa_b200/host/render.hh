// lisa_b200/host/render.hh — render()/display() of the reference (src/LiSA/include/render.hh:4-5) over the C ABI.
#pragma once
#include "lisa_rt.h"

// Headless render (render.cc:133-148): ONE subframe (index 0) of num_samples spp, then the PPM.
// Prints "Rendering finished in %.2f mn." like the reference.  Throws std::runtime_error on failure.
void render(lisa_ctx* ctx, const lisa_scene_desc& params);

// Progressive render (render.cc:75-131) without the GL window: subframes of min(16, num_samples) spp
// (optix_wrapper.cc:430) until subframe_index * spp >= num_samples, one stats line per subframe in place
// of the ImGui overlay (sutil.cpp:735-743, render.cc:110-115), then the PPM.
void display(lisa_ctx* ctx, const lisa_scene_desc& params);

// The same progressive render for a headless box: every `snapshot_every` subframes (0 = never)
// the PPM is rewritten and, if `checkpoint` is given, the accumulators are saved there (lisa_save_accum); with
// `resume` the accumulators are loaded from that file first and rendering continues at the subframe it holds, so an
// interrupted render finishes with the image it would have had.
struct DisplayOptions {
  unsigned    snapshot_every = 0;
  const char* checkpoint = nullptr;
  const char* resume = nullptr;
};
void display(lisa_ctx* ctx, const lisa_scene_desc& params, const DisplayOptions& opt);

// Sample-space partition over `ngpus` GPUs of this box, single process, through lisa_multi (include/lisa_rt.h):
// one context per GPU (full scene, own BVH), num_samples split into `ngpus` subframes of floor/ceil(num_samples / ngpus) spp
// (exactly num_samples in total), rendered concurrently, ONE ncclReduce of the float4 accumulators onto GPU 0, which writes
// the image.  Returns the wall time in seconds.
double render_multi(const lisa_scene_desc& params, int ngpus);
