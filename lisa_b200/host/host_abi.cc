// lisa_b200/host/host_abi.cc — C ABI over the C++ host front-end, for bindings (ctypes in tests/bench).
// Declared in include/lisa_host.h.
#include <cstring>
#include <string>

#include "lisa_host.h"
#include "parse_obj.hh"
#include "render.hh"
#include "scene_parser.hh"

struct lisa_scene {
  SceneParser* parser;
  lisa_scene_desc desc;
};

static thread_local std::string g_host_err;
static thread_local int         g_host_exit = 0;

extern "C" const char* lisa_host_last_error(void) { return g_host_err.c_str(); }
extern "C" int         lisa_host_last_exit_code(void) { return g_host_exit; }

extern "C" int lisa_scene_parse(const char* path, int load_meshes, lisa_scene** out) {
  if (!path || !out) { g_host_err = "null argument"; return LISA_ERR_ARG; }
  *out = nullptr;
  try {
    SceneParser* p = new SceneParser(path, load_meshes != 0);
    lisa_scene*  s = new lisa_scene{p, p->get_params()};
    *out = s;
    return LISA_OK;
  } catch (const SceneError& e) {
    g_host_err = e.what(); g_host_exit = e.exit_code;
    return LISA_ERR_IO;
  } catch (const std::exception& e) {
    g_host_err = e.what(); g_host_exit = 134;  // the reference would die on the uncaught exception
    return LISA_ERR_ARG;
  }
}
extern "C" void lisa_scene_free(lisa_scene* s) {
  if (!s) return;
  delete s->parser;
  delete s;
}
extern "C" const lisa_scene_desc* lisa_scene_get_desc(const lisa_scene* s) { return s ? &s->desc : nullptr; }
extern "C" int lisa_scene_num_meshes(const lisa_scene* s) { return s ? (int)s->parser->mesh_files().size() : 0; }
extern "C" const char* lisa_scene_mesh_file(const lisa_scene* s, int i, int* mat_idx) {
  if (!s || i < 0 || i >= (int)s->parser->mesh_files().size()) return nullptr;
  if (mat_idx) *mat_idx = s->parser->mesh_files()[i].second;
  return s->parser->mesh_files()[i].first.c_str();
}
extern "C" int lisa_scene_material_index(const lisa_scene* s, const char* name) {
  if (!s || !name) return -1;
  auto it = s->parser->material_names().find(name);
  return it == s->parser->material_names().end() ? -1 : it->second;
}
extern "C" int lisa_host_render(lisa_ctx* ctx, const lisa_scene* s, int progressive) {
  if (!ctx || !s) { g_host_err = "null argument"; return LISA_ERR_ARG; }
  try {
    if (progressive) display(ctx, s->desc); else render(ctx, s->desc);
    return LISA_OK;
  } catch (const std::exception& e) {
    g_host_err = e.what();
    return LISA_ERR_STATE;
  }
}
extern "C" void lisa_host_set_obj_cache(int enabled) { parse_obj_set_cache(enabled); }
