/* oracle/ref_kat_main.cc — known-answer generator.  TEST INFRASTRUCTURE.
 *
 * Compiles the reference's OWN headers on the host (they are included from
 * /root/reference by oracle/Makefile, never copied): src/cuda/random.h (tea,
 * lcg, rnd), src/cuda/helpers.h (make_color), src/LiSA/src/maths.cu,
 * src/LiSA/src/bsdfs/lambertian.cu, src/sutil/Camera.cpp.  Prints one JSON
 * document with inputs and outputs; tests/golden/ref_kat.json is this
 * program's output (see oracle/make_golden.py), and pins oracle/cpu_ref.c.
 *
 * Host-side evaluation order of make_float3(rng(s), rng(s), rng(s)) is
 * unspecified (g++ evaluates right-to-left, nvcc device code left-to-right:
 * the reference's PTX assigns draw 1 -> x, 2 -> y, 3 -> z); consumers of the
 * "hemisphere"/"bounce" tables compare as multisets of |component|.
 */
#include <cmath>
#include <cstdio>
static unsigned int optixGetPrimitiveIndex() { return 0; }
#include "bsdfs/lambertian.cu"
#include <sutil/Camera.h>

static unsigned g_s = 12345u;
static float    frand() { return rnd(g_s); }              /* harness-only input generator */
static float3   fdir() {
  float3 d;
  do { d = make_float3(frand() * 2 - 1, frand() * 2 - 1, frand() * 2 - 1); } while (dot(d, d) < 1e-3f);
  return normalize(d);
}
#define F3(v) (v).x, (v).y, (v).z

int main() {
  printf("{\n");
  /* tea<16> + rnd */
  const unsigned px[][2] = {{0, 0}, {0, 1}, {0, 7}, {1, 0}, {1999, 0}, {2000, 0}, {3999999, 0}, {8294399, 0},
                            {262143, 3}, {12345, 124}, {4194303, 255}, {77, 4095}};
  printf("\"tea16\": [\n");
  for (size_t i = 0; i < sizeof(px) / sizeof(px[0]); i++) {
    unsigned s  = tea<16>(px[i][0], px[i][1]);
    unsigned s0 = s;
    float a = rnd(s), b = rnd(s), c = rnd(s);
    printf("  {\"pixel\": %u, \"subframe\": %u, \"seed\": %u, \"rnd\": [%.9g, %.9g, %.9g], \"seed_after\": %u}%s\n",
           px[i][0], px[i][1], s0, a, b, c, s, i + 1 < sizeof(px) / sizeof(px[0]) ? "," : "");
  }
  printf("],\n");
  {
    unsigned s = 0;
    printf("\"lcg_from_0\": [");
    for (int i = 0; i < 8; i++) { lcg(s); printf("%u%s", s, i < 7 ? ", " : ""); }
    printf("],\n");
  }
  /* camera */
  printf("\"uvw\": [\n");
  const float cams[][8] = {{-0.01f, 0.015f, 0.6f, -0.01f, 0.015f, 0.f, 40.f, 1.f},
                           {-0.01f, 0.015f, 0.6f, -0.01f, 0.015f, 0.f, 40.f, 16.f / 9.f},
                           {1.f, 2.f, 3.f, -0.5f, 0.25f, -4.f, 63.5f, 1.5f},
                           {0.f, 0.f, 5.f, 1.f, 1.f, 0.f, 90.f, 0.75f}};
  for (int i = 0; i < 4; i++) {
    sutil::Camera c;
    c.setEye(make_float3(cams[i][0], cams[i][1], cams[i][2]));
    c.setLookat(make_float3(cams[i][3], cams[i][4], cams[i][5]));
    c.setUp(make_float3(0, 1, 0));
    c.setFovY(cams[i][6]);
    c.setAspectRatio(cams[i][7]);
    float3 U, V, W;
    c.UVWFrame(U, V, W);
    printf("  {\"eye\": [%.9g, %.9g, %.9g], \"look_at\": [%.9g, %.9g, %.9g], \"fov\": %.9g, \"aspect\": %.9g, "
           "\"U\": [%.9g, %.9g, %.9g], \"V\": [%.9g, %.9g, %.9g], \"W\": [%.9g, %.9g, %.9g]}%s\n",
           cams[i][0], cams[i][1], cams[i][2], cams[i][3], cams[i][4], cams[i][5], cams[i][6], cams[i][7], F3(U),
           F3(V), F3(W), i < 3 ? "," : "");
  }
  printf("],\n");
  /* make_color */
  printf("\"make_color\": [\n");
  const float cols[] = {0.f, 0.001f, 0.0031308f, 0.0031309f, 0.01f, 0.05f, 0.18f, 0.25f, 0.5f, 0.75f,
                        0.9f, 0.999f, 1.f, 2.f, -1.f, 0.0005f, 0.003f, 0.33333f, 0.66667f, 0.12345f};
  for (size_t i = 0; i < sizeof(cols) / sizeof(float); i++) {
    uchar4 c = make_color(make_float3(cols[i], cols[i], cols[i]));
    printf("  {\"in\": %.9g, \"out\": %d, \"alpha\": %d}%s\n", cols[i], c.x, c.w,
           i + 1 < sizeof(cols) / sizeof(float) ? "," : "");
  }
  printf("],\n");
  /* fresnel (Schlick) */
  printf("\"fresnel\": [\n");
  const float etas[] = {1.f / 1.5f, 1.5f, 1.f / 1.33f, 1.33f, 2.4f, 1.f};
  for (int e = 0; e < 6; e++)
    for (int k = 0; k <= 10; k++) {
      float c = k / 10.f;
      printf("  {\"cos\": %.9g, \"eta\": %.9g, \"out\": %.9g}%s\n", c, etas[e], BTDF(c, etas[e]),
             (e == 5 && k == 10) ? "" : ",");
    }
  printf("],\n");
  /* refract */
  printf("\"refract\": [\n");
  for (int i = 0; i < 48; i++) {
    float3 N   = fdir();
    float3 d   = fdir() * (0.5f + frand());          /* non-unit directions occur (bounce, Q5) */
    float  eta = etas[i % 5];
    float  cosI = dot(d, N);
    float3 Nn = N;
    if (cosI < 0) cosI = -cosI; else Nn = -N;
    float3 r = refract(cosI, d, Nn, eta);
    printf("  {\"cosI\": %.9g, \"dir\": [%.9g, %.9g, %.9g], \"N\": [%.9g, %.9g, %.9g], \"eta\": %.9g, "
           "\"out\": [%.9g, %.9g, %.9g]}%s\n", cosI, F3(d), F3(Nn), eta, F3(r), i < 47 ? "," : "");
  }
  printf("],\n");
  /* shoot_ray_hemisphere / BRDF / bounce */
  printf("\"hemisphere\": [\n");
  for (int i = 0; i < 48; i++) {
    float3   N  = i == 0 ? make_float3(0, 1, 0) : fdir();
    unsigned s  = i == 0 ? 0x741c187du : tea<16>(1000u + i, i);
    unsigned s0 = s;
    float3   h  = shoot_ray_hemisphere(N, s);
    Material m{};
    printf("  {\"seed\": %u, \"N\": [%.9g, %.9g, %.9g], \"out\": [%.9g, %.9g, %.9g], \"seed_after\": %u, "
           "\"brdf\": %.9g}%s\n", s0, F3(N), F3(h), s, BRDF(N, h, &m), i < 47 ? "," : "");
  }
  printf("],\n");
  printf("\"bounce\": [\n");
  const float roughs[] = {0.f, 0.25f, 0.5f, 1.f};
  for (int i = 0; i < 32; i++) {
    float3   N  = i == 0 ? make_float3(0, 1, 0) : fdir();
    float3   d  = i == 0 ? normalize(make_float3(1, -1, 0)) : fdir();
    unsigned s  = i == 0 ? 0x741c187du : tea<16>(2000u + i, i);
    unsigned s0 = s;
    Material m{};
    m.roughness = roughs[i % 4];
    float3 b    = bounce(d, N, s, &m);
    /* reflect part is order independent: report it separately */
    float3 refl = reflect(d, N);
    printf("  {\"seed\": %u, \"dir\": [%.9g, %.9g, %.9g], \"N\": [%.9g, %.9g, %.9g], \"roughness\": %.9g, "
           "\"out\": [%.9g, %.9g, %.9g], \"reflect\": [%.9g, %.9g, %.9g], \"length\": %.9g, \"seed_after\": %u}%s\n",
           s0, F3(d), F3(N), m.roughness, F3(b), F3(refl), length(b), s, i < 31 ? "," : "");
  }
  printf("],\n");
  /* barycentric_normal */
  printf("\"barycentric_normal\": [\n");
  for (int i = 0; i < 32; i++) {
    float3 v1, v2, v3, n1, n2, n3, P;
    if (i == 0) {
      v1 = make_float3(0, 0, 0); v2 = make_float3(1, 0, 0); v3 = make_float3(0, 1, 0);
      n1 = make_float3(1, 0, 0); n2 = make_float3(0, 1, 0); n3 = make_float3(0, 0, 1);
      P  = make_float3(.25f, .25f, 0);
    } else {
      v1 = fdir() * frand(); v2 = fdir() * frand(); v3 = fdir() * frand();
      n1 = fdir(); n2 = fdir(); n3 = fdir();
      float a = frand(), b = frand() * (1 - a);
      P = v1 * (1 - a - b) + v2 * a + v3 * b;
    }
    float3 n = barycentric_normal(P, n1, n2, n3, v1, v2, v3);
    printf("  {\"P\": [%.9g, %.9g, %.9g], \"n\": [[%.9g, %.9g, %.9g], [%.9g, %.9g, %.9g], [%.9g, %.9g, %.9g]], "
           "\"v\": [[%.9g, %.9g, %.9g], [%.9g, %.9g, %.9g], [%.9g, %.9g, %.9g]], \"out\": [%.9g, %.9g, %.9g]}%s\n",
           F3(P), F3(n1), F3(n2), F3(n3), F3(v1), F3(v2), F3(v3), F3(n), i < 31 ? "," : "");
  }
  printf("],\n");
  printf("\"sizeof_material\": %zu\n}\n", sizeof(Material));
  return 0;
}
