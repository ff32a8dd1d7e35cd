// lisa_b200/csrc/material.cuh — device material record.
// The host hands over the reference's 40-byte Material (src/LiSA/include/structs.hh:16-23,
// include/lisa_rt.h lisa_material); on the device it is repacked into three 16-byte words so a
// material is fetched with three LDG.128 instead of ten scalar loads.
#pragma once
#include <cuda_runtime.h>

namespace lisa {

struct DMaterial {
  float4 a;  // diffuse.xyz, roughness
  float4 b;  // emission.xyz, n (IOR)
  float4 c;  // alpha, emit (0/1), unused, unused
  __device__ __forceinline__ float3 diffuse() const { return make_float3(a.x, a.y, a.z); }
  __device__ __forceinline__ float3 emission() const { return make_float3(b.x, b.y, b.z); }
  __device__ __forceinline__ float  roughness() const { return a.w; }
  __device__ __forceinline__ float  ior() const { return b.w; }
  __device__ __forceinline__ float  alpha() const { return c.x; }
  __device__ __forceinline__ bool   emit() const { return c.y != 0.0f; }
};

__device__ __forceinline__ DMaterial load_material(const DMaterial* mats, int i) {
  const float4* p = reinterpret_cast<const float4*>(mats + i);
  DMaterial     m;
  m.a = __ldg(p);
  m.b = __ldg(p + 1);
  m.c = __ldg(p + 2);
  return m;
}

// Handle a BSDF receives instead of the reference's `const Material*`: the material table and an index, so that an
// implementation loads only the words it needs (the Lambertian BRDF needs none, its bounce only the roughness).
struct MatRef {
  const DMaterial* mats;
  int              id;
  __device__ __forceinline__ float roughness() const { return __ldg(reinterpret_cast<const float4*>(mats + id)).w; }
  __device__ __forceinline__ DMaterial load() const { return load_material(mats, id); }
};

}  // namespace lisa
