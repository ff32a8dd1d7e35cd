/* oracle/ref_parse_main.cc — dumps what the reference's own SceneParser
 * (src/LiSA/src/scene_parser.cc, parse_obj.cc; compiled in place from
 * /root/reference by oracle/Makefile) makes of a .rto file, as JSON.
 * TEST INFRASTRUCTURE: differential oracle for lisa_b200's parser.
 * The reference prints progress on stdout ("Importing ...") — the JSON is
 * therefore written to the file named by argv[2]. */
#include <cstdio>
#include "scene_parser.hh"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s scene.rto out.json [max_tris_dumped]\n", argv[0]); return 1; }
  SceneParser    parser(argv[1]);
  RendererParams p = parser.get_params();
  int            T = p.num_vertices / 3;
  int            maxd = argc > 3 ? atoi(argv[3]) : 64;
  FILE*          f = fopen(argv[2], "w");
  if (!f) { perror(argv[2]); return 2; }
  fprintf(f, "{\"width\": %u, \"height\": %u, \"num_samples\": %u, \"num_bounces\": %u, \"output_image\": \"",
          p.width, p.height, p.num_samples, p.num_bounces);
  for (const char* c = p.output_image; *c; c++) {
    if ((unsigned char)*c < 0x20) { fprintf(f, "\\u%04x", (unsigned)(unsigned char)*c); continue; }
    if (*c == '"' || *c == '\\') fputc('\\', f);
    fputc(*c, f);
  }
  fprintf(f, "\",\n \"camera\": {\"eye\": [%.9g, %.9g, %.9g], \"look_at\": [%.9g, %.9g, %.9g], \"fov\": %.9g},\n",
          p.camera.eye.x, p.camera.eye.y, p.camera.eye.z, p.camera.look_at.x, p.camera.look_at.y, p.camera.look_at.z,
          p.camera.fov);
  fprintf(f, " \"num_vertices\": %d, \"num_materials\": %d, \"sizeof_material\": %zu,\n \"materials\": [\n",
          p.num_vertices, p.num_materials, sizeof(Material));
  for (int i = 0; i < p.num_materials; i++) {
    const Material& m = p.materials[i];
    /* fields the reference leaves uninitialised (Q12) are reported as null */
    fprintf(f, "  {\"emit\": %s, \"alpha\": %.9g, ", m.emit ? "true" : "false", m.alpha);
    if (m.emit)
      fprintf(f, "\"emission\": [%.9g, %.9g, %.9g]}", m.emission_color.x, m.emission_color.y, m.emission_color.z);
    else {
      fprintf(f, "\"diffuse\": [%.9g, %.9g, %.9g], ", m.diffuse_color.x, m.diffuse_color.y, m.diffuse_color.z);
      if (m.alpha < 1.0f) fprintf(f, "\"n\": %.9g}", m.n);
      else fprintf(f, "\"roughness\": %.9g}", m.roughness);
    }
    fprintf(f, "%s\n", i + 1 < p.num_materials ? "," : "");
  }
  fprintf(f, " ],\n \"mat_indices_rle\": [");
  for (int i = 0; i < T;) {
    int j = i;
    while (j < T && p.mat_indices[j] == p.mat_indices[i]) j++;
    fprintf(f, "%s[%d, %d]", i ? ", " : "", p.mat_indices[i], j - i);
    i = j;
  }
  /* geometry: checksum over all floats + first triangles verbatim */
  double sv = 0, sn = 0, wv = 0, wn = 0;
  const float* v = reinterpret_cast<const float*>(p.vertices);
  const float* n = reinterpret_cast<const float*>(p.normals);
  for (int i = 0; i < p.num_vertices * 3; i++) {
    sv += v[i]; sn += n[i];
    wv += v[i] * (double)((i % 251) + 1); wn += n[i] * (double)((i % 241) + 1);
  }
  fprintf(f, "],\n \"vertex_sum\": %.12g, \"normal_sum\": %.12g, \"vertex_wsum\": %.12g, \"normal_wsum\": %.12g,\n", sv, sn,
          wv, wn);
  fprintf(f, " \"first_vertices\": [");
  for (int i = 0; i < p.num_vertices * 3 && i < maxd * 9; i++) fprintf(f, "%s%.9g", i ? ", " : "", v[i]);
  fprintf(f, "],\n \"first_normals\": [");
  for (int i = 0; i < p.num_vertices * 3 && i < maxd * 9; i++) fprintf(f, "%s%.9g", i ? ", " : "", n[i]);
  fprintf(f, "]}\n");
  fclose(f);
  return 0;
}
