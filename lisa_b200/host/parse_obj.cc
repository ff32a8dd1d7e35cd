// lisa_b200/host/parse_obj.cc — OBJ subset of the reference (src/LiSA/src/parse_obj.cc:24-69):
//   "v x y z", "vn x y z", "f a/b/c a/b/c a/b/c" (also a//c); tokens separated by single spaces;
//   only the first three vertices of a face are used (quads are truncated, not triangulated);
//   the normal index (third field) is mandatory; indices are 1-based and local to the file;
//   every other line is ignored.
// Differences, all on inputs for which the reference has undefined behaviour: runs of spaces are
// tolerated, and short lines / out-of-range indices raise an error instead of reading out of bounds.
// The whole file is read once and scanned in place (no per-line vector<string>) by one worker thread per chunk of
// whole lines, in three passes (count, positions/normals, faces): what makes multi-million-triangle soups load in
// seconds.  LISA_OBJ_THREADS overrides the worker count.
//
// Binary soup cache (opt-in: LISA_OBJ_CACHE=1, or `lisa --obj-cache`, or lisa_host_set_obj_cache(1)): after a text
// parse the de-indexed soup is written next to the OBJ as <file>.lisasoup; the next load of the same OBJ (same size and
// modification time) reads it back instead of parsing text — two freads of 36 bytes per triangle each.  A stale,
// truncated or foreign file is ignored and rewritten; a directory that cannot be written to just means no cache.
// Off by default: the reference never writes next to its inputs.
#include "parse_obj.hh"

#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

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

float to_float_slow(const char* b, const char* e, const std::string& path, long line) {
  char  buf[64];
  size_t n = (size_t)(e - b) < sizeof(buf) - 1 ? (size_t)(e - b) : sizeof(buf) - 1;
  memcpy(buf, b, n);
  buf[n] = 0;
  char* q;
  float v = strtof(buf, &q);  // like std::stof: leading float, trailing characters ignored
  if (q == buf) throw SceneError(path + ":" + std::to_string(line) + ": not a number: '" + buf + "'", 1);
  return v;
}

// The same value as strtof, without strtof, for the tokens OBJ files are made of: [+-]digits[.digits][e[+-]digits] with at
// most 15 significant digits and a decimal exponent within +-22.  Then the decimal is w * 10^k with w < 2^53 and 10^|k|
// exact in double, so ONE double multiplication or division gives the correctly rounded double (Clinger 1990), and
// rounding that to float equals rounding the decimal itself to float unless the double sits exactly on the midpoint of two
// floats (no float midpoint can lie strictly between a decimal and its nearest double: both are doubles).  Midpoints,
// overflow, subnormal floats and every token of another shape (inf, nan, hex, trailing characters, too long) take the
// strtof path, so the result is bit-identical to the reference's std::stof for every input.
float to_float(const char* b, const char* e, const std::string& path, long line) {
  static const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                                 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char* p = b;
  if (e - b > 40 || p >= e) return to_float_slow(b, e, path, line);
  bool neg = false;
  if (*p == '-' || *p == '+') { neg = *p == '-'; p++; }
  uint64_t w = 0;
  int      sig = 0, k = 0, ndig = 0;  // significant digits taken, decimal exponent, digits seen
  for (; p < e && *p >= '0' && *p <= '9'; p++, ndig++) {
    if (sig || *p != '0') { if (sig < 15) { w = w * 10 + (uint64_t)(*p - '0'); sig++; } else return to_float_slow(b, e, path, line); }
  }
  if (p < e && *p == '.') {
    p++;
    for (; p < e && *p >= '0' && *p <= '9'; p++, ndig++) {
      if (sig || *p != '0') { if (sig < 15) { w = w * 10 + (uint64_t)(*p - '0'); sig++; } else return to_float_slow(b, e, path, line); }
      k--;
    }
  }
  if (ndig == 0) return to_float_slow(b, e, path, line);
  if (p < e && (*p == 'e' || *p == 'E')) {
    p++;
    bool eneg = false;
    if (p < e && (*p == '-' || *p == '+')) { eneg = *p == '-'; p++; }
    int ex = 0, nd = 0;
    for (; p < e && *p >= '0' && *p <= '9' && nd < 4; p++, nd++) ex = ex * 10 + (*p - '0');
    if (nd == 0 || nd == 4) return to_float_slow(b, e, path, line);
    k += eneg ? -ex : ex;
  }
  if (p != e || k < -22 || k > 22) return to_float_slow(b, e, path, line);
  double d = (double)w;
  d = k >= 0 ? d * p10[k] : d / p10[-k];
  if (d != 0.0) {
    uint64_t bits;
    memcpy(&bits, &d, 8);
    if (d > 3.0e38 || d < 2.0e-38 || (bits & 0x1fffffffull) == 0x10000000ull) return to_float_slow(b, e, path, line);
  }
  const float f = (float)d;
  return neg ? -f : f;
}

// 1-based index held by the field [b, e)
inline long to_index(const char* b, const char* e) {
  // the field is [b, e): nothing outside it is ever read (the text may end right behind the last token).  An empty
  // field ("f 1// 2//2 3//3") yields 0, which the caller reports as an index out of range — the reference's
  // std::stoi throws on it (parse_obj.cc:52-54).  A sign is accepted like stoi does; negative indices are out of range.
  const char* p = b;
  bool neg = false;
  if (p < e && (*p == '+' || *p == '-')) { neg = *p == '-'; p++; }
  long v = 0;
  int  n = 0;
  for (; p < e && *p >= '0' && *p <= '9'; p++, n++) {
    if (n >= 18) return neg ? -1 : (long)0x7fffffffffffffffL;  // more digits than any index can have: out of range either way
    v = v * 10 + (*p - '0');
  }
  if (n == 0) return 0;
  return neg ? -v : v;
}

// ---- binary soup cache ---------------------------------------------------------------------------
struct SoupHeader {           // 48 bytes, little endian
  char     magic[8];          // "LISASOUP"
  uint32_t version;           // 1
  uint32_t header_bytes;      // sizeof(SoupHeader)
  uint64_t source_size;       // the OBJ this was made from: size ...
  int64_t  source_mtime_ns;   // ... and modification time
  uint64_t num_triangles;     // then 9 T floats (vertices) and 9 T floats (normals)
  uint64_t reserved;
};
static_assert(sizeof(SoupHeader) == 48, "SoupHeader layout");

int g_obj_cache = -1;  // -1: ask the environment

bool obj_cache_enabled() {
  if (g_obj_cache >= 0) return g_obj_cache != 0;
  const char* e = getenv("LISA_OBJ_CACHE");
  return e && atoi(e) != 0;
}

bool source_stamp(const std::string& path, uint64_t& size, int64_t& mtime_ns) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return false;
  size = (uint64_t)st.st_size;
  mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + (int64_t)st.st_mtim.tv_nsec;
  return true;
}

// true: the soup of `path` was appended from its cache file
bool soup_cache_load(const std::string& path, std::vector<float>& vertices, std::vector<float>& normals, size_t& num_tris) {
  uint64_t size; int64_t mtime;
  if (!source_stamp(path, size, mtime)) return false;
  FILE* f = fopen((path + ".lisasoup").c_str(), "rb");
  if (!f) return false;
  SoupHeader h;
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "LISASOUP", 8) == 0 && h.version == 1 && h.header_bytes == sizeof(h) &&
            h.source_size == size && h.source_mtime_ns == mtime;
  if (ok) {  // the file must hold exactly what the header promises
    struct stat st;
    ok = fstat(fileno(f), &st) == 0 && (uint64_t)st.st_size == sizeof(h) + h.num_triangles * 72ull;
  }
  if (ok) {
    const size_t n = (size_t)h.num_triangles * 9, v0 = vertices.size(), n0 = normals.size();
    vertices.resize(v0 + n);
    normals.resize(n0 + n);
    ok = (n == 0) || (fread(&vertices[v0], 4, n, f) == n && fread(&normals[n0], 4, n, f) == n);
    if (!ok) { vertices.resize(v0); normals.resize(n0); }
    else num_tris = (size_t)h.num_triangles;
  }
  fclose(f);
  return ok;
}

void soup_cache_store(const std::string& path, const float* v, const float* n, size_t num_tris) {
  SoupHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "LISASOUP", 8);
  h.version = 1; h.header_bytes = sizeof(h); h.num_triangles = num_tris;
  if (!source_stamp(path, h.source_size, h.source_mtime_ns)) return;
  const std::string dst = path + ".lisasoup", tmp = dst + ".tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const size_t cnt = num_tris * 9;
  const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && (cnt == 0 || (fwrite(v, 4, cnt, f) == cnt && fwrite(n, 4, cnt, f) == cnt));
  if (fclose(f) != 0 || !ok || rename(tmp.c_str(), dst.c_str()) != 0) remove(tmp.c_str());  // readers never see a partial file
}

}  // namespace

void parse_obj_set_cache(int enabled) { g_obj_cache = enabled < 0 ? -1 : (enabled != 0); }

namespace {

struct Chunk {
  const char* b;
  const char* e;          // [b, e): whole lines
  long        first_line; // 1-based number of the first line (for messages)
  size_t      nv = 0, nn = 0, nf = 0;
  std::string error;
};

// calls fn(kind, cursor-after-the-keyword, line_no) for every "v", "vn", "f" line of the chunk
template <typename F>
void for_each_record(const Chunk& c, F&& fn) {
  const char* p = c.b;
  long line_no = c.first_line - 1;
  while (p < c.e) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(c.e - p));
    if (!eol) eol = c.e;
    line_no++;
    if (*p != ' ') {  // a leading space makes the first token empty in the reference: line ignored
      Cursor cur{p, eol};
      const char *b, *e;
      if (cur.token(b, e)) {
        const size_t len = (size_t)(e - b);
        if (len == 1 && b[0] == 'v') fn(0, cur, line_no);
        else if (len == 2 && b[0] == 'v' && b[1] == 'n') fn(1, cur, line_no);
        else if (len == 1 && b[0] == 'f') fn(2, cur, line_no);
      }
    }
    p = eol + 1;
  }
}

}  // namespace

void parse_obj(const std::string& path, std::vector<float>& vertices, std::vector<float>& normals,
               std::vector<int32_t>& mat_indices, int mat_idx) {
  printf("Importing %s...\n", path.c_str());
  const bool dbg = getenv("LISA_DEBUG_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::steady_clock::now(); };
  auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  auto T0 = tnow();
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw SceneError(path + " not found.", -1);
  const bool use_cache = obj_cache_enabled();
  if (use_cache) {
    size_t nt_cached = 0;
    if (soup_cache_load(path, vertices, normals, nt_cached)) {
      fclose(f);
      mat_indices.resize(mat_indices.size() + nt_cached, mat_idx);
      if (dbg) fprintf(stderr, "parse_obj: %zu triangles from %s.lisasoup in %.1f ms\n", nt_cached, path.c_str(), tms(T0, tnow()));
      printf("Done. Imported %d triangles.\n", (int)nt_cached);
      return;
    }
  }
  fseek(f, 0, SEEK_END);
  long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  // The text is scanned in place.  Mapped straight from the page cache when the file does not end on a page boundary (the
  // zero tail of its last page then terminates a number that runs to the end of the file); read into a buffer with a
  // newline sentinel otherwise.
  struct Mapping {
    void* p = MAP_FAILED; size_t n = 0;
    ~Mapping() { if (p != MAP_FAILED) munmap(p, n); }
  } map;
  std::vector<char> buffer;
  const char*       text = nullptr;
  const long        page = sysconf(_SC_PAGESIZE);
  if (size > 0 && page > 0 && size % page != 0 && !getenv("LISA_OBJ_NO_MMAP")) {
    map.p = mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fileno(f), 0);
    map.n = (size_t)size;
    if (map.p != MAP_FAILED) text = static_cast<const char*>(map.p);
  }
  if (!text) {
    buffer.resize((size_t)size + 1);
    if (size > 0 && fread(buffer.data(), 1, (size_t)size, f) != (size_t)size) {
      fclose(f);
      throw SceneError(path + ": read error", -1);
    }
    buffer[(size_t)size] = '\n';
    text = buffer.data();
  }
  fclose(f);

  auto T1 = tnow();
  // chunks of whole lines, one per worker (a single chunk below 4 MB)
  unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
  if (size < (4 << 20)) nthreads = 1;
  if (const char* e = getenv("LISA_OBJ_THREADS")) nthreads = (unsigned)std::max(1, atoi(e));
  std::vector<Chunk> chunks;
  {
    const char* base = text;
    const char* stop = base + size;
    const char* p = base;
    for (unsigned k = 0; k < nthreads && p < stop; k++) {
      const char* want = k + 1 == nthreads ? stop : base + (size_t)size * (k + 1) / nthreads;
      if (want < p) want = p;
      const char* e = want >= stop ? stop : (const char*)memchr(want, '\n', (size_t)(stop - want));
      e = e ? std::min(e + 1, stop) : stop;
      chunks.push_back(Chunk{p, e, 0});
      p = e;
    }
  }
  auto run = [&](auto&& body) {
    if (chunks.size() == 1) { body(chunks[0]); return; }
    std::vector<std::thread> th;
    for (Chunk& c : chunks) th.emplace_back([&body, &c] { try { body(c); } catch (const std::exception& e) { c.error = e.what(); } });
    for (auto& t : th) t.join();
    for (Chunk& c : chunks) if (!c.error.empty()) throw SceneError(c.error, 1);
  };
  auto T2 = tnow();
  // pass 1: count records and lines per chunk
  std::vector<long> lines(chunks.size(), 0);
  run([&](Chunk& c) {
    long n = 0;
    for (const char* p = c.b; p < c.e;) { const char* eol = (const char*)memchr(p, '\n', (size_t)(c.e - p)); n++; p = eol ? eol + 1 : c.e; }
    lines[&c - chunks.data()] = n;
    c.first_line = 1;
    for_each_record(c, [&](int kind, Cursor&, long) { (kind == 0 ? c.nv : kind == 1 ? c.nn : c.nf)++; });
  });
  size_t tv = 0, tn = 0, tf = 0;
  std::vector<size_t> ov(chunks.size()), on(chunks.size()), of(chunks.size());
  long line0 = 1;
  for (size_t k = 0; k < chunks.size(); k++) {
    ov[k] = tv; on[k] = tn; of[k] = tf;
    tv += chunks[k].nv; tn += chunks[k].nn; tf += chunks[k].nf;
    chunks[k].first_line = line0;
    line0 += lines[k];
  }
  auto T3 = tnow();
  // pass 2: positions and normals
  std::vector<float> vt(3 * tv), nt(3 * tn);
  run([&](Chunk& c) {
    const size_t k = (size_t)(&c - chunks.data());
    size_t iv = ov[k], in = on[k];
    for_each_record(c, [&](int kind, Cursor& cur, long line_no) {
      if (kind == 2) return;
      float* dst = kind == 0 ? &vt[3 * iv++] : &nt[3 * in++];
      for (int j = 0; j < 3; j++) {
        const char *b, *e;
        if (!cur.token(b, e)) throw SceneError(path + ":" + std::to_string(line_no) + ": expected 3 components", 1);
        dst[j] = to_float(b, e, path, line_no);
      }
    });
  });
  auto T4 = tnow();
  // pass 3: faces -> de-indexed soup.  In the reference a face may only use vertices declared BEFORE it
  // (it indexes the arrays as they grow); files that respect that give the same result here.
  const size_t v0 = vertices.size(), n0 = normals.size(), m0 = mat_indices.size();
  vertices.resize(v0 + 9 * tf);
  normals.resize(n0 + 9 * tf);
  mat_indices.resize(m0 + tf, mat_idx);
  run([&](Chunk& c) {
    const size_t k = (size_t)(&c - chunks.data());
    size_t it = of[k];
    for_each_record(c, [&](int kind, Cursor& cur, long line_no) {
      if (kind != 2) return;
      for (int j = 0; j < 3; j++) {
        const char *b, *e;
        if (!cur.token(b, e)) throw SceneError(path + ":" + std::to_string(line_no) + ": face needs 3 vertices", 1);
        const char* s1 = (const char*)memchr(b, '/', (size_t)(e - b));
        const char* s2 = s1 ? (const char*)memchr(s1 + 1, '/', (size_t)(e - s1 - 1)) : nullptr;
        if (!s2) throw SceneError(path + ":" + std::to_string(line_no) + ": face vertex without normal index", 1);
        const long vi = to_index(b, s1), ni = to_index(s2 + 1, e);
        if (vi < 1 || (size_t)vi > tv || ni < 1 || (size_t)ni > tn)
          throw SceneError(path + ":" + std::to_string(line_no) + ": index out of range", 1);
        memcpy(&vertices[v0 + 9 * it + 3 * j], &vt[3 * (vi - 1)], 12);
        memcpy(&normals[n0 + 9 * it + 3 * j], &nt[3 * (ni - 1)], 12);
      }
      it++;
    });
  });
  auto T5 = tnow();
  if (use_cache) soup_cache_store(path, vertices.data() + v0, normals.data() + n0, tf);
  if (dbg) fprintf(stderr, "parse_obj: read %.1f chunk %.1f count %.1f v/vn %.1f faces %.1f ms (%zu chunks)\n", tms(T0, T1), tms(T1, T2), tms(T2, T3), tms(T3, T4), tms(T4, T5), chunks.size());
  printf("Done. Imported %d triangles.\n", (int)tf);
}
