// lisa_b200/csrc/scene.cuh — device-resident scene: packed triangles, materials, BVH(s), camera.
//
// Data layout in HBM (all arrays 16-byte aligned, read with LDG.128 through the read-only path):
//   tri_v   3 x float4 per triangle, FINAL (BVH leaf) order: (v0.xyz, material id), (v1.xyz, -), (v2.xyz, -)
//   tri_n   3 x float4 per triangle, same order:             (n0.xyz, original triangle index), (n1.xyz,-), (n2.xyz,-)
//   mats    3 x float4 per material (material.cuh)
//   bvh     binary nodes (4 x float4) or compressed 8-wide nodes (5 x float4), see traverse.cuh
// Emitters and non-emitters are kept in two BVHs over the same arrays (root_emit / root_other): a
// shadow ray asks "is the closest hit an emitter?" (the reference's __closesthit__occlusion,
// shader.cu:177-184), which becomes "closest emitter hit, then ANY occluder in front of it".
#pragma once
#include "common.cuh"
#include "material.cuh"

namespace lisa {

struct DCamera {
  float3   eye, U, V, W;
  uint32_t width, height;
};

struct DScene {
  const float4*    tri_v;
  const float4*    tri_n;
  const DMaterial* mats;
  const float4*    bvh;        // nodes of both BVHs
  int              root_other; // node index of the non-emitter BVH root, -1 if there is none
  int              root_emit;  // node index of the emitter BVH root, -1 if there is none
  int              wide;       // 1: compressed 8-wide nodes, 0: binary nodes
  int              num_tris;
  int              num_mats;
  int              single_light; // 1 when all emitter triangles share one material
  int              shadow_first_found; // LISA_SHADOW_FIRST_FOUND
  float3           emit_lo, emit_hi;   // bounds of all emitter triangles (padded): cheap reject before the emitter BVH
  float3           emit_c;             // centre and squared radius of the sphere around those bounds:
  float            emit_r2;            //   a shadow ray outside the cone (P, sphere) cannot reach an emitter
  int              cull;               // 1: resolve shadow tries that provably cannot change RayState::hit without traversal
  int              emit_flat;          // axis (0, 1, 2) along which the emitter bounds are flat (a quad light: all emitter
                                       //   triangles in one axis-aligned plane), -1 otherwise; flat bounds get an (almost)
                                       //   exact, division-free filter for shadow tries instead of the cone test
  float            emit_plane;         // the plane's coordinate on that axis
  float            emit_hh;            // half thickness of the padded bounds on that axis, plus slack
};

}  // namespace lisa
