// lisa_b200/host/scene_parser.hh — `.rto` scene front-end (seam B1).
//
// Same surface as the reference's SceneParser (src/LiSA/include/scene_parser.hh:10-45):
// construct from a path, get_params() hands out a RendererParams-shaped POD (lisa_scene_desc,
// include/lisa_rt.h) whose pointers borrow this object's storage.  The grammar accepted is the one the
// reference's std::regex patterns define (src/LiSA/src/scene_parser.cc:35-245), matched here by a small
// hand-written scanner (no std::regex: the reference's comment pattern `(.|\n)*?` recurses per
// character and overflows the stack on large files).
// Errors: the reference prints to stderr and calls exit(1) (scene_parser.cc:15-16,100-101,107-111,
// 166-170,201-205) or exit(-1) for a missing OBJ (parse_obj.cc:66-67).  Here a SceneError carrying the
// same message and exit code is thrown; main.cc turns it back into print + exit.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "lisa_rt.h"

struct SceneError : std::runtime_error {
  int exit_code;
  SceneError(const std::string& msg, int code = 1) : std::runtime_error(msg), exit_code(code) {}
};

class SceneParser {
 public:
  explicit SceneParser(const char* path, bool load_meshes = true);
  lisa_scene_desc get_params() const { return params; }

  // exposed for tests
  static std::string remove_comments(const std::string& file);
  const std::vector<std::pair<std::string, int>>& mesh_files() const { return meshes; }
  const std::map<std::string, int>& material_names() const { return mat_name_idx; }

 private:
  void build_materials(const std::string& file);
  void build_meshes(const std::string& file, bool load);
  void build_camera(const std::string& file);
  void mk_params();

  std::vector<float>         vertices;  // xyz packed, 3 vertices per triangle
  std::vector<float>         normals;
  std::vector<lisa_material> materials;
  std::vector<int32_t>       mat_indices;  // one per triangle
  std::vector<std::pair<std::string, int>> meshes;
  unsigned int width = 0, height = 0, num_samples = 0, num_bounces = 0;
  lisa_camera  camera{};
  std::string  output_image;
  std::map<std::string, int> mat_name_idx;
  lisa_scene_desc params{};
};
