// lisa_b200/csrc/bsdf/ggx.cuh — a second BSDF behind the same three-function seam (GGX is on
// the reference's TODO list, README.md:183).  Selected at compile time, like the reference selects its BSDF by
// `#include "bsdfs/lambertian.cu"` (shader.cu:4):   make -C lisa_b200 BSDF=ggx   (-> liblisa_rt_ggx.so)
//
// The reference's interface carries no view vector (BRDF(N, L, mat), lambertian.cu:15), so this is the part of a GGX
// microfacet model that fits it: bounce() samples a half vector from the GGX normal distribution of width
// alpha = roughness^2 and mirrors the incoming direction about it; BRDF() is the GGX lobe D evaluated for the half
// vector between N and L, energy-normalised like the Lambertian term (value 1/pi at alpha = 1, N.L = 1).
#pragma once
#include "../common.cuh"
#include "../material.cuh"

namespace lisa { namespace bsdf {

__device__ __forceinline__ float3 bounce(const float3& ray_dir, const float3& N, uint32_t& seed, const MatRef& mat) {
  const float r = mat.roughness(), a = fmaxf(r * r, 1e-3f);
  const float u1 = rnd(seed), u2 = rnd(seed);
  // GGX half vector around N: cos^2(theta) = (1 - u1) / (1 + (a^2 - 1) u1)
  const float ct = sqrtf((1.0f - u1) / (1.0f + (a * a - 1.0f) * u1)), st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f));
  const float ph = 6.283185307179586f * u2;
  const float3 up = fabsf(N.z) < 0.999f ? f3(0, 0, 1) : f3(1, 0, 0);
  const float3 T = normalize(cross(up, N)), B = cross(N, T);
  const float3 h = T * (st * __cosf(ph)) + B * (st * __sinf(ph)) + N * ct;
  float3 out = reflect(ray_dir, h);
  if (dot(out, N) * dot(ray_dir, N) > 0.0f) out = reflect(ray_dir, N);  // sampled below the surface: plain mirror
  return out;
}

__device__ __forceinline__ float BRDF(const float3& N, const float3& L, const MatRef& mat) {
  const float r = mat.roughness(), a = fmaxf(r * r, 1e-3f);
  const float ndl = fminf(fmaxf(dot(N, L), 0.0f), 1.0f);
  const float3 h = normalize(N + L);
  const float ndh = fminf(fmaxf(dot(N, h), 0.0f), 1.0f);
  const float d = ndh * ndh * (a * a - 1.0f) + 1.0f;
  const float D = (a * a) / (3.14159265358979f * d * d);  // GGX normal distribution
  return ndl * D * ndl;
}

__device__ __forceinline__ float BTDF(const float& cosI, const float& eta) { return fresnel(cosI, eta); }

}}  // namespace lisa::bsdf
