// lisa_b200/csrc/common.cuh — vector helpers, RNG and small device utilities shared by all kernels.
//
// Semantics follow the reference where it defines the image:
//   tea<16>/lcg/rnd ........ src/cuda/random.h:31-67
//   rng, shoot_ray_hemisphere, fresnel, refract .. src/LiSA/src/maths.cu:6-30
//   normalize/lerp/reflect/faceforward ............ src/sutil/vec_math.h:494-564
//   toSRGB/quantize/make_color ..................... src/cuda/helpers.h:107-138
// The reference is compiled with --use_fast_math (src/CMakeLists.txt:176) and so is this path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LISA_TMIN 1e-4f   // shader.cu:114,202
#define LISA_TMAX 1e16f   // shader.cu:115,202
#define LISA_SHADOW_TRIES 30  // shader.cu:199

namespace lisa {

__host__ __device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__host__ __device__ __forceinline__ float3 f3(const float4& v) { return make_float3(v.x, v.y, v.z); }
__host__ __device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
// Every multiply-add of the estimator is spelled out (fmaf) and the CUDA sources are compiled with -fmad=false, so that
// no kernel can contract an expression differently from another: the pipelines stay bit-identical by construction.
__host__ __device__ __forceinline__ float  dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
// a + s * b
__host__ __device__ __forceinline__ float3 madd(float3 a, float s, float3 b) { return f3(fmaf(s, b.x, a.x), fmaf(s, b.y, a.y), fmaf(s, b.z, a.z)); }
__host__ __device__ __forceinline__ float3 cross(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ float3 fmin3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__host__ __device__ __forceinline__ float3 fmax3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

// vec_math.h:539-543
__device__ __forceinline__ float3 normalize(float3 v) { return v * rsqrtf(dot(v, v)); }
// vec_math.h:494
__device__ __forceinline__ float3 lerp(float3 a, float3 b, float t) { return f3(fmaf(b.x - a.x, t, a.x), fmaf(b.y - a.y, t, a.y), fmaf(b.z - a.z, t, a.z)); }
// vec_math.h:552
__device__ __forceinline__ float3 reflect(float3 i, float3 n) { return madd(i, -2.0f * dot(n, i), n); }

// ---- RNG (src/cuda/random.h) ------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t tea16(uint32_t val0, uint32_t val1) {
  uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
__host__ __device__ __forceinline__ uint32_t lcg(uint32_t& prev) {
  prev = 1664525u * prev + 1013904223u;
  return prev & 0x00FFFFFFu;
}
// rnd = lcg / 2^24: the multiply by 2^-24 is exact, like the reference's div.approx by a power of two
__host__ __device__ __forceinline__ float rnd(uint32_t& prev) { return (float)lcg(prev) * (1.0f / 16777216.0f); }
// maths.cu:6-8
__device__ __forceinline__ float rng(uint32_t& seed) { return fmaf(rnd(seed), 2.0f, -1.0f); }

// Same value as rng() without the int->float conversion (which issues on the slow conversion pipe):
// 2*(u/2^24) - 1 = (u - 2^23) * 2^-23 is exactly representable, so building it from the low 23 bits with the
// 2^23 magic exponent and a constant chosen by bit 23 is bit-identical to the reference's fma(rnd, 2, -1).
__device__ __forceinline__ float rng_fast(uint32_t& seed) {
  seed = 1664525u * seed + 1013904223u;
  const float m = __uint_as_float(0x4B000000u | (seed & 0x007FFFFFu));     // 2^23 + low 23 bits, exact
  const float c = __uint_as_float(0xC0000000u - (seed & 0x00800000u));     // -2 if bit 23 is clear, -1 if set
  return fmaf(m, 1.1920928955078125e-07f, c);
}

// maths.cu:10-15 — draws land in x, y, z in call order (the reference's device code does the same)
__device__ __forceinline__ float3 shoot_ray_hemisphere(const float3& normal, uint32_t& seed) {
  float  a = rng_fast(seed), b = rng_fast(seed), c = rng_fast(seed);
  float3 d = normalize(f3(a, b, c));
  return d * copysignf(1.0f, dot(d, normal));  // faceforward(d, normal, d), vec_math.h:561
}

// maths.cu:17-20 — Schlick.  The reference evaluates pow() in double; x^2 and x^5 by multiplies in
// float differ from that by < 1e-7, far below the 2^-24 granularity of the rnd() it is compared with.
__device__ __forceinline__ float fresnel(float cosT, float eta) {
  float r  = (1.0f - eta) / (1.0f + eta);
  float R0 = r * r;
  float m  = 1.0f - cosT, m2 = m * m;
  return fmaf(1.0f - R0, m2 * m2 * m, R0);
}
// maths.cu:22-30 — zero vector under total internal reflection (Q7)
__device__ __forceinline__ float3 refract(float cosI, const float3& d, const float3& N, float eta) {
  float  cost2 = fmaf(-(eta * eta), fmaf(-cosI, cosI, 1.0f), 1.0f);
  float3 t     = madd(eta * d, fmaf(eta, cosI, -sqrtf(fabsf(cost2))), N);
  return cost2 > 0.0f ? t : f3(0.0f, 0.0f, 0.0f);
}

// ---- tonemap (src/cuda/helpers.h:107-138) ------------------------------------------------------
__device__ __forceinline__ float to_srgb(float c) {
  c = fminf(fmaxf(c, 0.0f), 1.0f);
  return c < 0.0031308f ? 12.92f * c : fmaf(1.055f, powf(c, 1.0f / 2.4f), -0.055f);
}
__device__ __forceinline__ uint32_t quantize8(float x) {
  x          = fminf(fmaxf(x, 0.0f), 1.0f);
  uint32_t v = (uint32_t)(x * 256.0f);
  return v < 255u ? v : 255u;
}
__device__ __forceinline__ uint32_t make_color(float3 c) {
  return quantize8(to_srgb(c.x)) | (quantize8(to_srgb(c.y)) << 8) | (quantize8(to_srgb(c.z)) << 16) | (255u << 24);
}

// ---- misc -----------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace lisa
