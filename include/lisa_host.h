/* include/lisa_host.h — C ABI of the host front-end (liblisa_host.so).
 *
 * The reference's front-end is C++ used directly by its main():
 *   SceneParser(char*) / get_params()   src/LiSA/include/scene_parser.hh:10-13
 *   parse_obj(...)                       src/LiSA/include/parse_obj.hh:7-11
 *   render() / display()                 src/LiSA/include/render.hh:4-5
 * lisa_b200/host re-implements those classes/functions under the same names (C++), and this header
 * flattens them for non-C++ callers.  Where the reference prints a message and exits
 * (scene_parser.cc:15-16,100-101,...; parse_obj.cc:66-67) these functions return an error, and
 * lisa_host_last_error()/lisa_host_last_exit_code() give the reference's message and exit code.
 */
#ifndef LISA_HOST_H
#define LISA_HOST_H
#include "lisa_rt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lisa_scene lisa_scene;

/* SceneParser(path).  load_meshes == 0 parses the grammar but does not open the OBJ files. */
int  lisa_scene_parse(const char* path, int load_meshes, lisa_scene** out);
void lisa_scene_free(lisa_scene* scene);
/* get_params(): pointers borrow the scene object (like RendererParams borrows SceneParser). */
const lisa_scene_desc* lisa_scene_get_desc(const lisa_scene* scene);
int         lisa_scene_num_meshes(const lisa_scene* scene);
const char* lisa_scene_mesh_file(const lisa_scene* scene, int i, int* mat_idx);
int         lisa_scene_material_index(const lisa_scene* scene, const char* name);
/* render() (progressive == 0) or display() (progressive != 0) on an existing context. */
int         lisa_host_render(lisa_ctx* ctx, const lisa_scene* scene, int progressive);
/* Binary soup cache for parse_obj (no counterpart in parse_obj.cc:24-69, which re-parses the text
 * on every run): 1 = write/read <file>.obj.lisasoup next to each OBJ, 0 = off, -1 = as the environment variable
 * LISA_OBJ_CACHE says (the default; unset = off). */
void        lisa_host_set_obj_cache(int enabled);
const char* lisa_host_last_error(void);
int         lisa_host_last_exit_code(void);

#ifdef __cplusplus
}
#endif
#endif
