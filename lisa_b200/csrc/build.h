// lisa_b200/csrc/build.h — host-side interface of the device BVH builder (bvh_build.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace lisa {

struct BuildInput {
  const float*         d_verts;    // 9 floats per triangle (the soup as uploaded)
  const float*         d_normals;  // 9 floats per triangle
  const int*           d_mat_idx;  // one per triangle
  const unsigned char* d_mat_emit; // one per material: 1 = emitter
  int                  num_tris;
  int                  num_mats;
  int                  wide;       // 1: compressed 8-wide nodes, 0: binary nodes
  int                  lbvh;       // 1: plain LBVH hierarchy (fast build), 0: PLOC (default, near-SAH quality)
  int                  ploc_radius;
  int                  rotate_passes;  // SAH tree-rotation passes over the binary hierarchy before the collapse (0 = none)
  float                leaf_sah;   // >= 0: leaves by the surface-area heuristic, the cost of one more child slot in triangle tests (k_collapse8); < 0: every subtree of <= 3 triangles is one leaf
  // Triangle splitting (split.cu): when set, the builder's primitives are REFERENCES — num_tris of them — to the triangles
  // of the soup: reference r is triangle ref_tri[r] with the box [ref_lo[r], ref_hi[r]].  nullptr: one primitive per triangle.
  const int*           d_ref_tri;
  const float4*        d_ref_lo;
  const float4*        d_ref_hi;
};

struct SplitOutput {
  int*    d_ref_tri;
  float4* d_ref_lo;
  float4* d_ref_hi;
  int     num_refs;
  float   cell;  // the grid cell size that was used
};
// References for the triangles of a soup (split.cu): at most budget * T + 64 of them.  Returns 0 or a negative lisa_status.
int  split_triangles(const float* d_verts, int num_tris, float budget, SplitOutput* out, cudaStream_t st, char* err, size_t errlen);
void split_free(SplitOutput* s);

struct BuildOutput {
  float4* d_nodes;          // node array shared by both BVHs
  float4* d_tri_v;          // 3 float4 per triangle, final order
  float4* d_tri_n;          // 3 float4 per triangle, final order
  int*    d_final_to_orig;  // final index -> caller's triangle index
  int     num_tris;        // primitives of the BVH = entries of d_tri_v / d_tri_n / d_final_to_orig (references when the triangles were split)
  int     num_input_tris;  // triangles of the caller's soup (set by the caller of build_bvh)
  int     num_nodes, nodes_other, nodes_emit;
  int     root_other, root_emit;
  int     num_emit_tris;
  size_t  node_bytes;
  float   box_other[6], box_emit[6];  // root bounds (lo.xyz, hi.xyz) of the two partitions
  float   sah_nodes_per_ray;          // surface-area estimate of the node visits of a ray that crosses the scene: sum of the wide nodes' areas / their root's (0 for the binary BVH)
  float   stage_ms[4];                // CUDA-event time of the build stages: Morton + sort, hierarchy (PLOC / LBVH, rotations), collapse, pack
};

// Returns 0 or a negative lisa_status; on failure err holds a message.  Synchronises the stream.
int build_bvh(const BuildInput& in, BuildOutput* out, cudaStream_t st, char* err, size_t errlen);

}  // namespace lisa
