// lisa_b200/host/parse_obj.hh — minimal OBJ loader (same contract as src/LiSA/include/parse_obj.hh:7-11).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

// Appends the de-indexed triangle soup of `obj_file_path` (3 vertices + 3 normals per face, one material
// index per face) and prints the reference's two progress lines.  Throws SceneError (exit code -1) when
// the file cannot be opened, like parse_obj.cc:66-67.
void parse_obj(const std::string& obj_file_path, std::vector<float>& vertices, std::vector<float>& normals,
               std::vector<int32_t>& mat_indices, int mat_idx);

// Binary soup cache next to the OBJ (<file>.lisasoup, see parse_obj.cc): 1 on, 0 off, -1 = as LISA_OBJ_CACHE says (default).
void parse_obj_set_cache(int enabled);
