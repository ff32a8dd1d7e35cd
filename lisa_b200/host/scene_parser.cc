// lisa_b200/host/scene_parser.cc — grammar-compatible `.rto` parser (see scene_parser.hh).
// Each matcher below states the reference pattern it reproduces (src/LiSA/src/scene_parser.cc).
#include "scene_parser.hh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "parse_obj.hh"

namespace {

using std::string;
typedef size_t pos_t;
const pos_t NPOS = string::npos;

// ECMAScript \s
inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
// var_rgx "([a-zA-Z0-9]|_)+"  (scene_parser.hh:43)
inline bool is_var(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || is_digit(c) || c == '_'; }
// '.' of ECMAScript: anything but a line terminator
inline bool is_dot(char c) { return c != '\n' && c != '\r'; }

inline pos_t skip_ws(const string& s, pos_t p) { while (p < s.size() && is_ws(s[p])) p++; return p; }

// [+-]?([0-9]*[.])?[0-9]+   -> end of the longest match starting at p, or NPOS
pos_t match_float(const string& s, pos_t p) {
  pos_t q = p;
  if (q < s.size() && (s[q] == '+' || s[q] == '-')) q++;
  pos_t d0 = q;
  while (q < s.size() && is_digit(s[q])) q++;
  if (q < s.size() && s[q] == '.' && q + 1 < s.size() && is_digit(s[q + 1])) {
    q++;
    while (q < s.size() && is_digit(s[q])) q++;
    return q;
  }
  return q > d0 ? q : NPOS;
}

float to_float(const string& s, pos_t b, pos_t e) { return std::stof(s.substr(b, e - b)); }

// begin \s*=\s*  -> position after it, or NPOS   (p is just after `begin`)
pos_t match_eq(const string& s, pos_t p) {
  p = skip_ws(s, p);
  if (p >= s.size() || s[p] != '=') return NPOS;
  return skip_ws(s, p + 1);
}

// search_vector (scene_parser.cc:66-79):  begin\s*=\s*\( F , F , ... \)  with F = \s*float\s*
bool search_vector(const string& s, const string& begin, int n, float* out) {
  for (pos_t at = s.find(begin); at != NPOS; at = s.find(begin, at + 1)) {
    pos_t p = match_eq(s, at + begin.size());
    if (p == NPOS || p >= s.size() || s[p] != '(') continue;
    p++;
    bool ok = true;
    for (int i = 0; i < n && ok; i++) {
      p = skip_ws(s, p);
      pos_t e = match_float(s, p);
      if (e == NPOS) { ok = false; break; }
      out[i] = to_float(s, p, e);
      p = skip_ws(s, e);
      const char want = (i + 1 < n) ? ',' : ')';
      if (p >= s.size() || s[p] != want) ok = false;
      p++;
    }
    if (ok) return true;
  }
  return false;
}

// search_float (scene_parser.cc:81-83) for a list of literal alternatives tried at every position in
// order; the first (leftmost) position where one of them matches wins.
bool search_float(const string& s, const std::vector<string>& begins, float* out) {
  for (pos_t at = 0; at < s.size(); at++) {
    for (const string& b : begins) {
      if (s.compare(at, b.size(), b) != 0) continue;
      pos_t p = match_eq(s, at + b.size());
      if (p == NPOS) continue;
      pos_t e = match_float(s, p);
      if (e == NPOS) continue;
      *out = to_float(s, p, e);
      return true;
    }
  }
  return false;
}

// search_int (scene_parser.cc:85-87,89-103):  name\s*=\s*([0-9]+)
bool search_int(const string& s, const string& name, unsigned int* out) {
  for (pos_t at = s.find(name); at != NPOS; at = s.find(name, at + 1)) {
    pos_t p = match_eq(s, at + name.size());
    if (p == NPOS) continue;
    pos_t e = p;
    while (e < s.size() && is_digit(s[e])) e++;
    if (e == p) continue;
    *out = (unsigned int)std::stoi(s.substr(p, e - p));  // stoi like the reference (throws on overflow)
    return true;
  }
  return false;
}

// output_image\s*=\s*(([a-zA-Z0-9]|_|\/|.)+)   — '.' unescaped: the rest of the line, trailing blanks included.
// The greedy \s* may swallow newlines; if nothing capturable follows, the regex backtracks into the blanks.
bool search_path(const string& s, const string& name, string* out) {
  for (pos_t at = s.find(name); at != NPOS; at = s.find(name, at + 1)) {
    pos_t p = skip_ws(s, at + name.size());
    if (p >= s.size() || s[p] != '=') continue;
    const pos_t ws0 = p + 1;
    pos_t       q   = skip_ws(s, ws0);
    while (true) {
      if (q < s.size() && is_dot(s[q])) {
        pos_t e = q;
        while (e < s.size() && is_dot(s[e])) e++;
        *out = s.substr(q, e - q);
        return true;
      }
      if (q == ws0) break;
      q--;
    }
  }
  return false;
}

// all non-overlapping matches of   keyword <sep> \{ \n* [^}]*   where <sep> is
//   named == true :  \s+ ([a-zA-Z0-9]|_)+ \s*     (material)
//   named == false:  \s*                           (mesh, camera)
std::vector<string> match_blocks(const string& s, const string& keyword, bool named) {
  std::vector<string> res;
  pos_t from = 0;
  while (true) {
    pos_t at = s.find(keyword, from);
    if (at == NPOS) break;
    pos_t p = at + keyword.size();
    bool  ok = true;
    if (named) {
      pos_t w = skip_ws(s, p);
      if (w == p) ok = false;
      p = w;
      pos_t v = p;
      while (v < s.size() && is_var(s[v])) v++;
      if (v == p) ok = false;
      p = v;
    }
    if (ok) {
      p = skip_ws(s, p);
      if (p >= s.size() || s[p] != '{') ok = false;
    }
    if (!ok) { from = at + 1; continue; }
    pos_t e = s.find('}', p + 1);
    if (e == NPOS) e = s.size();
    res.push_back(s.substr(at, e - at));
    from = e;
  }
  return res;
}

string read_file(const char* path) {  // scene_parser.cc:5-19: getline + "\n"
  std::ifstream f(path);
  if (!f.is_open()) throw SceneError(string(path) + " not found", 1);
  string res, line;
  while (std::getline(f, line)) { res += line; res += "\n"; }
  return res;
}

}  // namespace

// scene_parser.cc:59-64:  /\*((.|\n)*?)\*/  -> ""   (non-greedy, multi-line; an unterminated "/*" stays)
std::string SceneParser::remove_comments(const std::string& file) {
  string out;
  out.reserve(file.size());
  pos_t p = 0;
  while (true) {
    pos_t a = file.find("/*", p);
    if (a == NPOS) break;
    pos_t b = file.find("*/", a + 2);
    if (b == NPOS) break;
    out.append(file, p, a - p);
    p = b + 2;
  }
  out.append(file, p, NPOS);
  return out;
}

SceneParser::SceneParser(const char* path, bool load_meshes) {
  string file = remove_comments(read_file(path));
  auto need_int = [&](const char* name, unsigned int* v) {
    if (!search_int(file, name, v)) throw SceneError(string("Param ") + name + " not found.\n", 1);
  };
  // same order as scene_parser.cc:45-55
  need_int("width", &width);
  need_int("height", &height);
  need_int("num_samples", &num_samples);
  need_int("num_bounces", &num_bounces);
  if (!search_path(file, "output_image", &output_image)) throw SceneError("Param output_image not found.\n", 1);
  build_materials(file);
  build_meshes(file, load_meshes);
  build_camera(file);
  mk_params();
}

// scene_parser.cc:106-163
void SceneParser::build_materials(const std::string& file) {
  auto err = [](const string& e) { return SceneError("Error in materials declaration: " + e + "\n", 1); };
  std::vector<string> blocks = match_blocks(file, "material", true);
  if (blocks.empty()) throw err("no valid materials declared.");
  for (const string& s : blocks) {
    // name:  \s+(VAR)\s*\{   — first match in the block
    string name;
    for (pos_t at = 0; at < s.size() && name.empty(); at++) {
      if (!is_ws(s[at])) continue;
      pos_t p = skip_ws(s, at), v = p;
      while (v < s.size() && is_var(s[v])) v++;
      if (v == p) continue;
      pos_t b = skip_ws(s, v);
      if (b < s.size() && s[b] == '{') name = s.substr(p, v - p);
    }
    // emit\s*=\s*true
    bool is_light = false;
    for (pos_t at = s.find("emit"); at != NPOS && !is_light; at = s.find("emit", at + 1)) {
      pos_t p = match_eq(s, at + 4);
      if (p != NPOS && s.compare(p, 4, "true") == 0) is_light = true;
    }
    float col[4];
    if (!search_vector(s, "color", 4, col)) throw err("no valid color provided in material " + name + ".");
    float r = col[0], g = col[1], b = col[2], alpha = col[3];
    if (r > 1 || g > 1 || b > 1) {  // scene_parser.cc:139-143 (double division, Q13)
      r = (float)(r / 255.0); g = (float)(g / 255.0); b = (float)(b / 255.0);
    }
    lisa_material m;
    memset(&m, 0, sizeof(m));  // the reference leaves unused fields uninitialised (Q12)
    if (is_light) {            // structs.hh:25-31
      m.emit = 1;
      m.emission_color[0] = r; m.emission_color[1] = g; m.emission_color[2] = b;
      m.alpha = 1.0f;
    } else {
      float param;
      // literal alternatives of the reference pattern "(roughness | n)": "roughness " and " n"
      if (!search_float(s, {"roughness ", " n"}, &param)) throw err("no valid roughness|refractive index in material " + name + ".");
      if (alpha < 1.0f) m.n = param; else m.roughness = param;  // structs.hh:41-52
      m.alpha = alpha;
      m.diffuse_color[0] = r; m.diffuse_color[1] = g; m.diffuse_color[2] = b;
      m.emit = 0;
    }
    if (mat_name_idx.find(name) == mat_name_idx.end()) {  // first declaration wins
      materials.push_back(m);
      mat_name_idx[name] = (int)materials.size() - 1;
    }
  }
}

// scene_parser.cc:165-198
void SceneParser::build_meshes(const std::string& file, bool load) {
  auto err = [](const string& e) { return SceneError("Error in mesh declaration: " + e + "\n", 1); };
  for (const string& s : match_blocks(file, "mesh", false)) {
    // material\s*=\s*(VAR)
    string mname;
    bool   found = false;
    for (pos_t at = s.find("material"); at != NPOS && !found; at = s.find("material", at + 1)) {
      pos_t p = match_eq(s, at + 8);
      if (p == NPOS) continue;
      pos_t v = p;
      while (v < s.size() && is_var(s[v])) v++;
      if (v == p) continue;
      mname = s.substr(p, v - p);
      found = true;
    }
    if (!found) throw err("no valid material provided");
    auto it = mat_name_idx.find(mname);
    if (it == mat_name_idx.end()) throw err(mname + " material not found");
    // obj_file\s*=\s*(.+\.obj)  — greedy: up to the LAST ".obj" on the line
    string path;
    found = false;
    for (pos_t at = s.find("obj_file"); at != NPOS && !found; at = s.find("obj_file", at + 1)) {
      pos_t p = skip_ws(s, at + 8);
      if (p >= s.size() || s[p] != '=') continue;
      const pos_t ws0 = p + 1;
      pos_t       q   = skip_ws(s, ws0);
      while (!found) {
        pos_t e = q;
        while (e < s.size() && is_dot(s[e])) e++;
        // last ".obj" in [q+1, e) preceded by at least one character
        pos_t k = string(s, q, e - q).rfind(".obj");
        if (k != NPOS && k >= 1) { path = s.substr(q, k + 4); found = true; break; }
        if (q == ws0) break;
        q--;
      }
    }
    if (!found) throw err("no valid obj_file provided");
    meshes.emplace_back(path, it->second);
    if (load) parse_obj(path, vertices, normals, mat_indices, it->second);
  }
}

// scene_parser.cc:200-245
void SceneParser::build_camera(const std::string& file) {
  auto err = [](const string& e) { return SceneError("Error in camera declaration: " + e + "\n", 1); };
  std::vector<string> cams = match_blocks(file, "camera", false);
  if (cams.size() != 1) throw err("no valid camera declared.");
  const string& s = cams[0];
  if (!search_vector(s, "position", 3, camera.eye)) throw err("no valid position provided.");
  if (!search_vector(s, "look_at", 3, camera.look_at)) throw err("no valid look_at provided.");
  if (!search_float(s, {"fov"}, &camera.fov)) throw err("no valid fov provided.");
}

// scene_parser.cc:247-262
void SceneParser::mk_params() {
  params.vertices      = vertices.data();
  params.normals       = normals.data();
  params.materials     = materials.data();
  params.mat_indices   = mat_indices.data();
  params.num_vertices  = (int32_t)(vertices.size() / 3);
  params.num_materials = (int32_t)materials.size();
  params.width         = width;
  params.height        = height;
  params.camera        = camera;
  params.num_samples   = num_samples;
  params.num_bounces   = num_bounces;
  params.output_image  = output_image.c_str();
}
