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
};

struct BuildOutput {
  float4* d_nodes;          // node array shared by both BVHs
  float4* d_tri_v;          // 3 float4 per triangle, final order
  float4* d_tri_n;          // 3 float4 per triangle, final order
  int*    d_final_to_orig;  // final index -> caller's triangle index
  int     num_tris;
  int     num_nodes, nodes_other, nodes_emit;
  int     root_other, root_emit;
  int     num_emit_tris;
  size_t  node_bytes;
  float   box_other[6], box_emit[6];  // root bounds (lo.xyz, hi.xyz) of the two partitions
  float   stage_ms[4];                // CUDA-event time of the build stages: Morton + sort, hierarchy (PLOC / LBVH, rotations), collapse, pack
};

// Returns 0 or a negative lisa_status; on failure err holds a message.  Synchronises the stream.
int build_bvh(const BuildInput& in, BuildOutput* out, cudaStream_t st, char* err, size_t errlen);

}  // namespace lisa
