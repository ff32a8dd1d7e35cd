// lisa_b200/host/parse_obj.cc — OBJ subset of the reference (src/LiSA/src/parse_obj.cc:24-69):
//   "v x y z", "vn x y z", "f a/b/c a/b/c a/b/c" (also a//c); tokens separated by single spaces;
//   only the first three vertices of a face are used (quads are truncated, not triangulated);
//   the normal index (third field) is mandatory; indices are 1-based and local to the file;
//   every other line is ignored.
// Differences, all on inputs for which the reference has undefined behaviour: runs of spaces are
// tolerated, and short lines / out-of-range indices raise an error instead of reading out of bounds.
// The whole file is read once and scanned in place (no per-line vector<string>), which is what makes
// multi-million-triangle soups loadable in seconds.
#include "parse_obj.hh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "scene_parser.hh"

namespace {

struct Cursor {
  const char* p;
  const char* end;  // end of the current line
  void skip_spaces() { while (p < end && *p == ' ') p++; }
  bool token(const char*& b, const char*& e) {
    skip_spaces();
    if (p >= end) return false;
    b = p;
    while (p < end && *p != ' ') p++;
    e = p;
    return true;
  }
};

float to_float(const char* b, const char* e, const std::string& path, long line) {
  char  buf[64];
  size_t n = (size_t)(e - b) < sizeof(buf) - 1 ? (size_t)(e - b) : sizeof(buf) - 1;
  memcpy(buf, b, n);
  buf[n] = 0;
  char* q;
  float v = strtof(buf, &q);  // like std::stof: leading float, trailing characters ignored
  if (q == buf) throw SceneError(path + ":" + std::to_string(line) + ": not a number: '" + buf + "'", 1);
  return v;
}

}  // namespace

void parse_obj(const std::string& path, std::vector<float>& vertices, std::vector<float>& normals,
               std::vector<int32_t>& mat_indices, int mat_idx) {
  printf("Importing %s...\n", path.c_str());
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw SceneError(path + " not found.", -1);
  fseek(f, 0, SEEK_END);
  long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> data((size_t)size + 1);
  if (size > 0 && fread(data.data(), 1, (size_t)size, f) != (size_t)size) {
    fclose(f);
    throw SceneError(path + ": read error", -1);
  }
  fclose(f);
  data[(size_t)size] = '\n';

  std::vector<float> vt, nt;
  int  nb_triangles = 0;
  long line_no = 0;
  const char* p = data.data();
  const char* const stop = p + size;
  while (p < stop) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(stop - p) + 1);
    line_no++;
    Cursor c{p, eol};
    const char *b, *e;
    if (*p != ' ' && c.token(b, e)) {  // a leading space makes the first token empty in the reference: line ignored
      const size_t len = (size_t)(e - b);
      if ((len == 1 && b[0] == 'v') || (len == 2 && b[0] == 'v' && b[1] == 'n')) {
        std::vector<float>& dst = len == 1 ? vt : nt;
        for (int k = 0; k < 3; k++) {
          if (!c.token(b, e)) throw SceneError(path + ":" + std::to_string(line_no) + ": expected 3 components", 1);
          dst.push_back(to_float(b, e, path, line_no));
        }
      } else if (len == 1 && b[0] == 'f') {
        for (int k = 0; k < 3; k++) {
          if (!c.token(b, e)) throw SceneError(path + ":" + std::to_string(line_no) + ": face needs 3 vertices", 1);
          // a/b/c : fields 0 and 2
          const char* s1 = (const char*)memchr(b, '/', (size_t)(e - b));
          const char* s2 = s1 ? (const char*)memchr(s1 + 1, '/', (size_t)(e - s1 - 1)) : nullptr;
          if (!s2) throw SceneError(path + ":" + std::to_string(line_no) + ": face vertex without normal index", 1);
          long vi = strtol(b, nullptr, 10), ni = strtol(s2 + 1, nullptr, 10);
          if (vi < 1 || (size_t)vi * 3 > vt.size() || ni < 1 || (size_t)ni * 3 > nt.size())
            throw SceneError(path + ":" + std::to_string(line_no) + ": index out of range", 1);
          vertices.insert(vertices.end(), vt.begin() + 3 * (vi - 1), vt.begin() + 3 * vi);
          normals.insert(normals.end(), nt.begin() + 3 * (ni - 1), nt.begin() + 3 * ni);
        }
        mat_indices.push_back(mat_idx);
        nb_triangles++;
      }
    }
    p = eol + 1;
  }
  printf("Done. Imported %d triangles.\n", nb_triangles);
}
