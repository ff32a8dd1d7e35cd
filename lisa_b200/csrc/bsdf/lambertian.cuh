// lisa_b200/csrc/bsdf/lambertian.cuh — the interchangeable BSDF (seam B4).
//
// Same three functions, same argument meaning as the reference's
// src/LiSA/src/bsdfs/lambertian.cu:7-27; the integrator (estimator.cuh and the sched_*.cuh kernels) only ever calls
// bsdf::bounce / bsdf::BRDF / bsdf::BTDF, and the implementation is chosen at compile time by
// which header LISA_BSDF_HEADER names (default: this file), like the reference's
// `#include "bsdfs/lambertian.cu"` (shader.cu:4).  All three are pure except for advancing `seed`.
#pragma once
#include "../common.cuh"
#include "../material.cuh"

namespace lisa { namespace bsdf {

// lambertian.cu:7-13 — lerp(mirror direction, hemisphere sample, roughness); NOT normalised (Q5).
__device__ __forceinline__ float3 bounce(const float3& ray_dir, const float3& N, uint32_t& seed, const MatRef& mat) {
  return lerp(reflect(ray_dir, N), shoot_ray_hemisphere(N, seed), mat.roughness());
}

// lambertian.cu:15-22 — clamp(N.L, 0, 1)^2 / pi (no pdf division, Q3).
__device__ __forceinline__ float BRDF(const float3& N, const float3& L, const MatRef& /*mat*/) {
  float NdotL = fminf(fmaxf(dot(N, L), 0.0f), 1.0f);
  return NdotL * (NdotL * 0.318309886183790672f);
}

// lambertian.cu:24-27 — reflection probability at a dielectric interface (Schlick).
__device__ __forceinline__ float BTDF(const float& cosI, const float& eta) { return fresnel(cosI, eta); }

}}  // namespace lisa::bsdf
