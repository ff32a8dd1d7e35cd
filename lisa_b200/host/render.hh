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

// Sample-space partition over `ngpus` GPUs of this box, single process: the num_samples are split into
// `ngpus` subframes of ceil(num_samples / ngpus) spp; GPU g (own context: full scene, own BVH) renders subframe g on
// its own host thread; the sums are added onto GPU 0 by lisa_accum_add_peer (one kernel reading peer memory over
// NVLink) and GPU 0 writes the PPM.  Returns the render wall time in seconds.
double render_multi(const lisa_scene_desc& params, int ngpus);
