/* oracle/cpu_ref.c — CPU restatement of the reference's render path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (lisa_b200/, include/lisa_rt.h) never links, imports or executes it.
 *
 * What it restates (gaetanserre/LiSA, paths relative to the reference root):
 *   integrator ............ src/LiSA/src/shader.cu:102-167 (raygen, trace),
 *                           :172-255 (miss / closest-hit programs, 30-try light sampling)
 *   maths ................. src/LiSA/src/maths.cu:6-71
 *   BSDF .................. src/LiSA/src/bsdfs/lambertian.cu:7-27
 *   RNG ................... src/cuda/random.h:31-67 (tea<16>, lcg, rnd)
 *   tonemap ............... src/cuda/helpers.h:107-138
 *   camera ................ src/sutil/Camera.cpp:34-45, src/LiSA/src/optix_wrapper.cc:430-442
 *   PPM ................... src/sutil/sutil.cpp:523-554, 97-117 (vertical flip, P6)
 * Third-party arithmetic NOT in /root/reference: NVIDIA OptiX 7.4.0 (ABI 55,
 * src/optix/include/optix.h:37; implementation in the driver's
 * libnvoptix.so.1) owns BVH build, traversal order and the ray/triangle test.
 * It is restated here as an exact closest-hit query (double-precision
 * Moller-Trumbore over a median-split BVH).  Two behaviours are decided by
 * that closed-source code and are therefore POLICY here (DESIGN.md "Parity"):
 *   Q2 shadow rays terminate on the first-FOUND hit in OptiX; here the
 *      CLOSEST hit decides (the only traversal-order-independent rule);
 *   Q7 a null direction (refract() under total internal reflection) is a miss.
 *
 * Pinning: every helper is checked against tests/golden/ref_kat.json (values
 * produced by compiling the reference's own headers, oracle/ref_kat_main.cc),
 * and whole images are checked against accumulators dumped by the unmodified
 * reference OptiX renderer on a B200 (oracle/optix_ref_main.cc,
 * tests/golden/optix_*.npz).
 *
 * Random draws are assigned x, y, z in call order, as the reference's device
 * code does (PTX of shader.cu: first lcg step -> .x); see ref_kat_main.cc.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } f3;

/* src/LiSA/include/structs.hh:16-23 — 40 bytes */
typedef struct {
  float   roughness;
  float   alpha;
  float   n;
  float   diffuse[3];
  uint8_t emit;
  uint8_t pad[3];
  float   emission[3];
} orc_material;

static inline f3    mk(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3    add(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3    sub(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3    mul(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3    scale(f3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3    cross(f3 a, f3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* src/sutil/vec_math.h:539-543 */
static inline f3 normalize(f3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return scale(v, inv); }
/* vec_math.h:494 lerp = a + t*(b-a) ; :552 reflect = i - 2 n dot(n,i) ; :561 faceforward = n * copysign(1, dot(i, nref)) */
static inline f3 lerp3(f3 a, f3 b, float t) { return add(a, scale(sub(b, a), t)); }
static inline f3 reflect3(f3 i, f3 n) { return sub(i, scale(n, 2.0f * dot(n, i))); }
static inline float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }

/* ---------------------------------------------------------------- RNG: src/cuda/random.h:31-67 */
uint32_t orc_tea16(uint32_t val0, uint32_t val1) {
  uint32_t v0 = val0, v1 = val1, s0 = 0;
  for (int n = 0; n < 16; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
uint32_t orc_lcg(uint32_t* prev) {
  *prev = 1664525u * *prev + 1013904223u;
  return *prev & 0x00FFFFFFu;
}
float orc_rnd(uint32_t* prev) { return (float)orc_lcg(prev) / (float)0x01000000; }
/* maths.cu:6-8 */
static inline float rng(uint32_t* s) { return orc_rnd(s) * 2.0f - 1.0f; }

/* ---------------------------------------------------------------- maths.cu:10-15 */
static f3 hemisphere(f3 n, uint32_t* seed) {
  float a = rng(seed), b = rng(seed), c = rng(seed);
  f3    d = normalize(mk(a, b, c));
  float s = copysignf(1.0f, dot(d, n)); /* faceforward(d, n, d) */
  return scale(d, s);
}
void orc_hemisphere(const float* N, uint32_t* seed, float* out) {
  f3 h = hemisphere(mk(N[0], N[1], N[2]), seed);
  out[0] = h.x; out[1] = h.y; out[2] = h.z;
}
/* maths.cu:17-20 — Schlick, evaluated in double like the reference's pow() calls */
float orc_fresnel(float cosT, float eta) {
  double r  = (double)((1.0f - eta) / (1.0f + eta));
  double R0 = r * r;
  return (float)(R0 + (1.0 - R0) * pow((double)(1.0f - cosT), 5.0));
}
/* maths.cu:22-30 */
static f3 refract3(float cosI, f3 d, f3 N, float eta) {
  float cost2 = 1.0f - eta * eta * (1.0f - cosI * cosI);
  f3    t     = add(scale(d, eta), scale(N, eta * cosI - sqrtf(fabsf(cost2))));
  return scale(t, cost2 > 0 ? 1.0f : 0.0f);
}
void orc_refract(float cosI, const float* d, const float* N, float eta, float* out) {
  f3 r = refract3(cosI, mk(d[0], d[1], d[2]), mk(N[0], N[1], N[2]), eta);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* maths.cu:33-57 */
static f3 barycentric_normal(f3 P, f3 n1, f3 n2, f3 n3, f3 v1, f3 v2, f3 v3) {
  f3    e1 = sub(v2, v1), e2 = sub(v3, v1), i = sub(P, v1);
  float d00 = dot(e1, e1), d01 = dot(e1, e2), d11 = dot(e2, e2), d20 = dot(i, e1), d21 = dot(i, e2);
  float denom = d00 * d11 - d01 * d01;
  float w = (d00 * d21 - d01 * d20) / denom;
  float v = (d11 * d20 - d01 * d21) / denom;
  float u = 1 - v - w;
  return normalize(add(add(scale(n1, u), scale(n2, v)), scale(n3, w)));
}
void orc_barycentric_normal(const float* P, const float* n, const float* v, float* out) {
  f3 r = barycentric_normal(mk(P[0], P[1], P[2]), mk(n[0], n[1], n[2]), mk(n[3], n[4], n[5]), mk(n[6], n[7], n[8]),
                            mk(v[0], v[1], v[2]), mk(v[3], v[4], v[5]), mk(v[6], v[7], v[8]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* The interchangeable BSDF (seam B4, shader.cu:4 `#include "bsdfs/lambertian.cu"`): 0 = the reference's Lambertian
 * (lambertian.cu:7-27), 1 = the GGX variant the product can be built with (lisa_b200/csrc/bsdf/ggx.cuh; GGX is on the
 * reference's TODO list, README.md:183-184, so there is no reference code to follow for it: the oracle restates the
 * variant's own definition so that the GGX kernel is checked against something, same seeds, same estimator). */
static int g_bsdf = 0;
void orc_set_bsdf(int which) { g_bsdf = which; }

/* ggx.cuh: half vector from the GGX normal distribution of width a = roughness^2 (two rnd() draws), the incoming direction
 * mirrored about it; a sample below the surface falls back to the plain mirror direction */
static f3 ggx_bounce(f3 d, f3 N, uint32_t* seed, float roughness) {
  float a  = fmaxf(roughness * roughness, 1e-3f);
  float u1 = orc_rnd(seed), u2 = orc_rnd(seed);
  float ct = sqrtf((1.0f - u1) / (1.0f + (a * a - 1.0f) * u1)), st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f));
  float ph = 6.283185307179586f * u2;
  f3    up = fabsf(N.z) < 0.999f ? mk(0, 0, 1) : mk(1, 0, 0);
  f3    T = normalize(cross(up, N)), B = cross(N, T);
  f3    h = add(add(scale(T, st * cosf(ph)), scale(B, st * sinf(ph))), scale(N, ct));
  f3    out = reflect3(d, h);
  if (dot(out, N) * dot(d, N) > 0.0f) out = reflect3(d, N);
  return out;
}
/* ggx.cuh: the GGX lobe D for the half vector between N and L, times clamp(N.L)^2 like the Lambertian term */
static float ggx_brdf(f3 N, f3 L, float roughness) {
  float a   = fmaxf(roughness * roughness, 1e-3f);
  float ndl = clampf(dot(N, L), 0.0f, 1.0f);
  f3    h   = normalize(add(N, L));
  float ndh = clampf(dot(N, h), 0.0f, 1.0f);
  float d   = ndh * ndh * (a * a - 1.0f) + 1.0f;
  float D   = (a * a) / (3.14159265358979f * d * d);
  return ndl * D * ndl;
}

/* lambertian.cu:7-13 */
static f3 bounce(f3 d, f3 N, uint32_t* seed, float roughness) {
  if (g_bsdf == 1) return ggx_bounce(d, N, seed, roughness);
  f3 refl = reflect3(d, N);
  f3 h    = hemisphere(N, seed);
  return lerp3(refl, h, roughness);
}
void orc_bounce(const float* d, const float* N, uint32_t* seed, float roughness, float* out) {
  f3 r = bounce(mk(d[0], d[1], d[2]), mk(N[0], N[1], N[2]), seed, roughness);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
/* lambertian.cu:15-22 — NdotL * float(NdotL / M_PI) */
static float brdf(f3 N, f3 L) {
  float ndl  = clampf(dot(N, L), 0.0f, 1.0f);
  float prob = (float)((double)ndl / M_PI);
  return ndl * prob;
}
float orc_brdf(const float* N, const float* L) { return brdf(mk(N[0], N[1], N[2]), mk(L[0], L[1], L[2])); }

/* ---------------------------------------------------------------- helpers.h:107-138 */
static float to_srgb1(float c) {
  float invGamma = 1.0f / 2.4f;
  float p        = powf(c, invGamma);
  return c < 0.0031308f ? 12.92f * c : 1.055f * p - 0.055f;
}
static uint8_t quant8(float x) {
  x = clampf(x, 0.0f, 1.0f);
  enum { N = (1 << 8) - 1, Np1 = (1 << 8) };
  uint32_t v = (uint32_t)(x * (float)Np1);
  return (uint8_t)(v < (uint32_t)N ? v : (uint32_t)N);
}
void orc_make_color(const float* rgb, uint8_t* out) {
  for (int i = 0; i < 3; i++) out[i] = quant8(to_srgb1(clampf(rgb[i], 0.0f, 1.0f)));
  out[3] = 255;
}

/* ---------------------------------------------------------------- Camera.cpp:34-45 */
void orc_uvw(const float* eye, const float* look, float fov_deg, float aspect, float* U, float* V, float* W) {
  f3    w    = sub(mk(look[0], look[1], look[2]), mk(eye[0], eye[1], eye[2]));
  float wlen = sqrtf(dot(w, w));
  f3    up   = mk(0, 1, 0); /* optix_wrapper.cc:437 */
  f3    u    = normalize(cross(w, up));
  f3    v    = normalize(cross(u, w));
  float vlen = wlen * tanf(0.5f * fov_deg * (float)M_PI / 180.0f);
  v          = scale(v, vlen);
  float ulen = vlen * aspect;
  u          = scale(u, ulen);
  U[0] = u.x; U[1] = u.y; U[2] = u.z;
  V[0] = v.x; V[1] = v.y; V[2] = v.z;
  W[0] = w.x; W[1] = w.y; W[2] = w.z;
}

/* ---------------------------------------------------------------- scene + exact closest hit */
typedef struct {
  float   lo[3], hi[3];
  int32_t left, right; /* children, or (first, -count) for a leaf when right < 0 */
} orc_node;

typedef struct {
  int           ntris, nmat;
  f3*           v;  /* 3 per triangle */
  f3*           n;  /* 3 per triangle */
  int32_t*      mat;
  orc_material* mats;
  int32_t*      order; /* triangle ids in leaf order */
  orc_node*     nodes;
  int           nnodes;
} orc_scene;

static void tri_bounds(const orc_scene* s, int t, float* lo, float* hi) {
  for (int a = 0; a < 3; a++) { lo[a] = 1e30f; hi[a] = -1e30f; }
  for (int k = 0; k < 3; k++) {
    const float* p = &s->v[3 * t + k].x;
    for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], p[a]); hi[a] = fmaxf(hi[a], p[a]); }
  }
}
static float tri_centroid(const orc_scene* s, int t, int a) {
  return ((&s->v[3 * t].x)[a] + (&s->v[3 * t + 1].x)[a] + (&s->v[3 * t + 2].x)[a]) * (1.0f / 3.0f);
}
static int build_rec(orc_scene* s, int first, int count) {
  int       id = s->nnodes++;
  orc_node* nd = &s->nodes[id];
  for (int a = 0; a < 3; a++) { nd->lo[a] = 1e30f; nd->hi[a] = -1e30f; }
  float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = first; i < first + count; i++) {
    float lo[3], hi[3];
    tri_bounds(s, s->order[i], lo, hi);
    for (int a = 0; a < 3; a++) {
      nd->lo[a] = fminf(nd->lo[a], lo[a]); nd->hi[a] = fmaxf(nd->hi[a], hi[a]);
      float c = tri_centroid(s, s->order[i], a);
      clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c);
    }
  }
  if (count <= 4) { nd->left = first; nd->right = -count; return id; }
  int ax = 0;
  if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
  if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
  float mid = 0.5f * (clo[ax] + chi[ax]);
  int   i = first, j = first + count - 1;
  while (i <= j) {
    if (tri_centroid(s, s->order[i], ax) < mid) i++;
    else { int t = s->order[i]; s->order[i] = s->order[j]; s->order[j] = t; j--; }
  }
  int nl = i - first;
  if (nl == 0 || nl == count) nl = count / 2; /* degenerate: split by index */
  int l = build_rec(s, first, nl);
  int r = build_rec(s, first + nl, count - nl);
  nd = &s->nodes[id];
  nd->left = l; nd->right = r;
  return id;
}

orc_scene* orc_scene_create(const float* verts, const float* normals, const int32_t* mat_idx, int ntris,
                            const void* materials, int nmat) {
  orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
  s->ntris = ntris; s->nmat = nmat;
  s->v    = (f3*)malloc(sizeof(f3) * 3 * (size_t)(ntris ? ntris : 1));
  s->n    = (f3*)malloc(sizeof(f3) * 3 * (size_t)(ntris ? ntris : 1));
  s->mat  = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntris ? ntris : 1));
  s->mats = (orc_material*)malloc(sizeof(orc_material) * (size_t)(nmat ? nmat : 1));
  memcpy(s->v, verts, sizeof(f3) * 3 * (size_t)ntris);
  memcpy(s->n, normals, sizeof(f3) * 3 * (size_t)ntris);
  memcpy(s->mat, mat_idx, sizeof(int32_t) * (size_t)ntris);
  memcpy(s->mats, materials, sizeof(orc_material) * (size_t)nmat);
  s->order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ntris ? ntris : 1));
  for (int i = 0; i < ntris; i++) s->order[i] = i;
  s->nodes  = (orc_node*)malloc(sizeof(orc_node) * (size_t)(2 * ntris + 2));
  s->nnodes = 0;
  if (ntris > 0) build_rec(s, 0, ntris);
  return s;
}
void orc_scene_destroy(orc_scene* s) {
  if (!s) return;
  free(s->v); free(s->n); free(s->mat); free(s->mats); free(s->order); free(s->nodes); free(s);
}

/* Closest hit with t in (tmin, tmax), t in units of |dir| (dir may be non-unit, Q5).
 * Double-precision Moller-Trumbore; no culling.  Returns the primitive index or -1. */
int orc_closest_hit(const orc_scene* s, const float* o, const float* d, float tmin, float tmax, float* t_out) {
  if (s->ntris == 0) return -1;
  if (d[0] == 0.0f && d[1] == 0.0f && d[2] == 0.0f) return -1; /* Q7 */
  double inv[3];
  for (int a = 0; a < 3; a++) inv[a] = 1.0 / (double)d[a];
  double best = tmax;
  int    prim = -1;
  int    stack[128], sp = 0;
  stack[sp++] = 0;
  while (sp) {
    const orc_node* nd = &s->nodes[stack[--sp]];
    double t0 = tmin, t1 = best;
    int    ok = 1;
    for (int a = 0; a < 3 && ok; a++) {
      if (d[a] == 0.0f) {
        if (o[a] < nd->lo[a] || o[a] > nd->hi[a]) ok = 0;
        continue;
      }
      double ta = ((double)nd->lo[a] - o[a]) * inv[a], tb = ((double)nd->hi[a] - o[a]) * inv[a];
      if (ta > tb) { double t = ta; ta = tb; tb = t; }
      /* pad so that boxes never reject a triangle the exact test would accept */
      ta -= 1e-9 * fabs(ta) + 1e-12; tb += 1e-9 * fabs(tb) + 1e-12;
      if (ta > t0) t0 = ta;
      if (tb < t1) t1 = tb;
      if (t0 > t1) ok = 0;
    }
    if (!ok) continue;
    if (nd->right < 0) {
      for (int i = nd->left; i < nd->left - nd->right; i++) {
        int       t  = s->order[i];
        const f3 *a = &s->v[3 * t], *b = &s->v[3 * t + 1], *c = &s->v[3 * t + 2];
        double e1[3] = {(double)b->x - a->x, (double)b->y - a->y, (double)b->z - a->z};
        double e2[3] = {(double)c->x - a->x, (double)c->y - a->y, (double)c->z - a->z};
        double p[3]  = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
        double det   = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
        if (det == 0.0) continue;
        double id   = 1.0 / det;
        double tv[3] = {(double)o[0] - a->x, (double)o[1] - a->y, (double)o[2] - a->z};
        double u     = (tv[0] * p[0] + tv[1] * p[1] + tv[2] * p[2]) * id;
        if (u < 0.0 || u > 1.0) continue;
        double q[3] = {tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0]};
        double v    = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * id;
        if (v < 0.0 || u + v > 1.0) continue;
        double tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * id;
        if (tt > tmin && tt < best) { best = tt; prim = t; }
      }
    } else {
      stack[sp++] = nd->left;
      stack[sp++] = nd->right;
    }
  }
  if (prim >= 0 && t_out) *t_out = (float)best;
  return prim;
}

/* ---------------------------------------------------------------- integrator */
typedef struct {
  uint64_t radiance_rays, shadow_rays, samples, null_dirs;
} orc_counters;

typedef struct {
  const orc_scene* s;
  int              bounces;
  int              shadow_tries; /* 30, shader.cu:199 */
} orc_ctx;

/* Debug aid for parity triage: per-pixel event bits (1 glass event, 2 null direction, 4 emitter hit by a
 * bounce ray, 8 escaped, 16 light sample found, 32 all 30 tries failed at least once). */
static uint8_t*              g_flagbuf = 0;
static _Thread_local uint8_t t_flags;
void orc_set_flag_buffer(uint8_t* buf) { g_flagbuf = buf; }

/* shader.cu:102-124 + the programs it reaches.  Returns the radiance of one sample. */
static f3 trace_path(const orc_ctx* c, f3 org, f3 dir, uint32_t* seed, orc_counters* cnt) {
  const orc_scene* s = c->s;
  /* RayState, shader.cu:30-42 */
  f3  atten = mk(1, 1, 1), color = mk(0, 0, 0);
  int hit = 0, light = -1; /* Q1: sticky across bounces of one sample */
  for (int b = 0; b < c->bounces; b++) {
    float t;
    int   null_dir = (dir.x == 0.0f && dir.y == 0.0f && dir.z == 0.0f);
    if (null_dir) { cnt->null_dirs++; t_flags |= 2; break; } /* Q7: treated as a miss, not counted as a ray */
    cnt->radiance_rays++;
    int prim = orc_closest_hit(s, &org.x, &dir.x, 1e-4f, 1e16f, &t);
    if (prim < 0) { t_flags |= 8; break; } /* __miss__radiance: bg = 0, done */
    const orc_material* m = &s->mats[s->mat[prim]];
    if (m->emit) { /* shader.cu:216-218 */
      color = add(color, mul(mk(m->emission[0], m->emission[1], m->emission[2]), atten));
      t_flags |= 4;
      break;
    }
    f3 P = add(org, scale(dir, t)); /* :221 */
    f3 N = barycentric_normal(P, s->n[3 * prim], s->n[3 * prim + 1], s->n[3 * prim + 2], s->v[3 * prim],
                              s->v[3 * prim + 1], s->v[3 * prim + 2]);
    f3 ndir = dir;
    if (m->alpha < 1.0f) { /* :226-246 */
      t_flags |= 1;
      float cosI = dot(dir, N), eta;
      f3    Nn;
      if (cosI < 0.0f) { cosI = -cosI; eta = 1 / m->n; Nn = N; }
      else { atten = mul(atten, mk(m->diffuse[0], m->diffuse[1], m->diffuse[2])); eta = m->n; Nn = scale(N, -1.0f); }
      if (eta == 1.0f) ndir = dir;
      else if (orc_rnd(seed) <= orc_fresnel(cosI, eta)) ndir = reflect3(dir, Nn);
      else ndir = refract3(cosI, dir, Nn, eta);
    } else { /* :248-253 */
      atten = mul(atten, mk(m->diffuse[0], m->diffuse[1], m->diffuse[2]));
      f3 L  = mk(0, 0, 0);
      for (int i = 0; i < c->shadow_tries; i++) { /* :196-209 */
        f3    w = hemisphere(N, seed);
        float ts;
        cnt->shadow_rays++;
        int sp = orc_closest_hit(s, &P.x, &w.x, 1e-4f, 1e16f, &ts);
        if (sp < 0) hit = 0;                                      /* __miss__occlusion */
        else if (s->mats[s->mat[sp]].emit) { hit = 1; light = s->mat[sp]; } /* __closesthit__occlusion */
        /* else: unchanged (Q1) */
        if (hit) {
          const orc_material* lm = &s->mats[light];
          L = scale(mk(lm->emission[0], lm->emission[1], lm->emission[2]), g_bsdf == 1 ? ggx_brdf(N, w, m->roughness) : brdf(N, w));
          t_flags |= 16;
          break;
        }
      }
      if (!hit) t_flags |= 32;
      color = add(color, mul(L, atten));
      ndir  = bounce(dir, N, seed, m->roughness);
    }
    org = P;
    dir = ndir;
  }
  return color;
}

typedef struct {
  float    eye[3], U[3], V[3], W[3];
  uint32_t width, height;
} orc_camera;

/* One pixel of __raygen__rg (shader.cu:126-167) for subframe f with S samples; returns the MEAN. */
static f3 render_pixel(const orc_ctx* c, const orc_camera* cam, uint32_t x, uint32_t y, uint32_t f, uint32_t S,
                       orc_counters* cnt) {
  float    sx = (float)cam->width, sy = (float)cam->height;
  uint32_t seed = orc_tea16((uint32_t)((float)y * sx + (float)x), f);
  f3       U = mk(cam->U[0], cam->U[1], cam->U[2]), V = mk(cam->V[0], cam->V[1], cam->V[2]);
  f3       W = mk(cam->W[0], cam->W[1], cam->W[2]), eye = mk(cam->eye[0], cam->eye[1], cam->eye[2]);
  f3       acc = mk(0, 0, 0);
  for (uint32_t i = 0; i < S; i++) {
    float jx = rng(&seed), jy = rng(&seed);
    float dx = (2.0f * (float)x + jx) / sx - 1.0f;
    float dy = (2.0f * (float)y + jy) / sy - 1.0f;
    f3    dir = normalize(add(add(scale(U, dx), scale(V, dy)), W));
    acc       = add(acc, trace_path(c, eye, dir, &seed, cnt));
    cnt->samples++;
  }
  return scale(acc, 1.0f / (float)S) /* accum / S */;
}

/* Renders subframes first..first+count-1 of S samples each and merges them with the reference's running
 * mean (shader.cu:160-164).  accum: W*H*4 floats, row 0 = image bottom, alpha = 1.  If pixels != NULL only
 * the listed pixel indices (y*W+x) are rendered into accum[4*i] (i = position in the list).
 * counters: 4 uint64 {radiance rays, shadow rays, samples, null directions}.  Returns threads used. */
int orc_render(const orc_scene* s, const float* eye, const float* look, float fov, uint32_t width, uint32_t height,
               uint32_t bounces, uint32_t first, uint32_t count, uint32_t S, const uint32_t* pixels, uint64_t npixels,
               float* accum, uint64_t* counters, int nthreads) {
  orc_ctx    c = {s, (int)bounces, 30};
  orc_camera cam;
  memcpy(cam.eye, eye, 12);
  orc_uvw(eye, look, fov, (float)width / (float)height, cam.U, cam.V, cam.W);
  cam.width = width; cam.height = height;
  uint64_t     n = pixels ? npixels : (uint64_t)width * height;
  orc_counters tot = {0, 0, 0, 0};
  int          used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    orc_counters cnt = {0, 0, 0, 0};
#ifdef _OPENMP
#pragma omp single
    used = omp_get_num_threads();
#endif
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)n; i++) {
      uint32_t p = pixels ? pixels[i] : (uint32_t)i;
      uint32_t x = p % width, y = p / width;
      f3       mean = mk(0, 0, 0);
      t_flags = 0;
      for (uint32_t f = first; f < first + count; f++) {
        f3 cur = render_pixel(&c, &cam, x, y, f, S, &cnt);
        if (f > 0 && f > first) mean = lerp3(mean, cur, 1.0f / (float)(f + 1));
        else mean = cur; /* subframe 0 (or the first of a partial range) overwrites */
      }
      if (g_flagbuf) g_flagbuf[i] = t_flags;
      accum[4 * i + 0] = mean.x; accum[4 * i + 1] = mean.y; accum[4 * i + 2] = mean.z; accum[4 * i + 3] = 1.0f;
    }
#pragma omp critical
    {
      tot.radiance_rays += cnt.radiance_rays; tot.shadow_rays += cnt.shadow_rays;
      tot.samples += cnt.samples; tot.null_dirs += cnt.null_dirs;
    }
  }
  if (counters) {
    counters[0] = tot.radiance_rays; counters[1] = tot.shadow_rays; counters[2] = tot.samples; counters[3] = tot.null_dirs;
  }
  return used;
}

/* One camera ray exactly as raygen builds it; for primary-ray parity tests. */
void orc_primary_ray(const float* eye, const float* look, float fov, uint32_t width, uint32_t height, uint32_t x,
                     uint32_t y, uint32_t f, float* dir, uint32_t* seed_after) {
  float U[3], V[3], W[3];
  orc_uvw(eye, look, fov, (float)width / (float)height, U, V, W);
  uint32_t seed = orc_tea16((uint32_t)((float)y * (float)width + (float)x), f);
  float    jx = rng(&seed), jy = rng(&seed);
  float    dx = (2.0f * (float)x + jx) / (float)width - 1.0f, dy = (2.0f * (float)y + jy) / (float)height - 1.0f;
  f3       d = normalize(mk(U[0] * dx + V[0] * dy + W[0], U[1] * dx + V[1] * dy + W[1], U[2] * dx + V[2] * dy + W[2]));
  dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
  *seed_after = seed;
}

/* sutil.cpp:523-554 + 97-117: sRGB-quantise, flip vertically, "P6\nW H\n255\n". accum is the linear mean. */
int orc_write_ppm(const char* path, const float* accum, uint32_t width, uint32_t height) {
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  fprintf(f, "P6\n%u %u\n255\n", width, height);
  for (int y = (int)height - 1; y >= 0; y--)
    for (uint32_t x = 0; x < width; x++) {
      uint8_t c[4];
      orc_make_color(&accum[4 * ((size_t)y * width + x)], c);
      fwrite(c, 1, 3, f);
    }
  fclose(f);
  return 0;
}
