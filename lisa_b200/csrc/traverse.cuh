// lisa_b200/csrc/traverse.cuh — ray/triangle test and BVH traversal (device code).
//
// Replaces what the reference gets from the closed-source OptiX runtime through optixTrace
// (src/LiSA/src/shader.cu:62,86): closest-hit and shadow queries over one triangle soup.
//   * ray/triangle: watertight test of Woop, Benthin, Wald (JCGT 2013): shear the triangle into
//     ray space, three edge functions computed WITHOUT fma contraction so that the two triangles
//     sharing an edge see exactly opposite values, double-precision fallback on a zero edge value;
//     no culling (the reference never culls), t in units of |dir| (directions are not unit, Q5).
//   * BVH: binary LBVH nodes (64 B, both child boxes in one node) or the compressed 8-wide nodes
//     of Ylitie, Karras, Laine (HPG 2017; 80 B, 8-bit quantised child boxes, octant-ordered
//     traversal with a (node-group, hit-mask) stack).
//   * traversal stack: the first entries live in shared memory ([entry][thread] layout, so a warp
//     touches 32 consecutive banks), deeper levels spill to a per-thread local array.
#pragma once
#include "common.cuh"

namespace lisa {

struct Hit {
  float t;
  float u, v;  // barycentric weights of vertex 1 and vertex 2 (vertex 0 has 1-u-v)
  int   prim;  // index into the packed triangle arrays, -1 = miss
};

// ------------------------------------------------------------------------------------------------
// Watertight ray/triangle
// ------------------------------------------------------------------------------------------------
struct RayPre {
  float3 o;
  int    kz;
  float  Sx, Sy, Sz;
};

__device__ __forceinline__ float3 perm3(const float3& v, int kz) {
  // (v[kx], v[ky], v[kz]) with kx = (kz+1)%3, ky = (kz+2)%3 (a pure rotation: winding is irrelevant without culling)
  return f3(kz == 0 ? v.y : (kz == 1 ? v.z : v.x), kz == 0 ? v.z : (kz == 1 ? v.x : v.y),
            kz == 0 ? v.x : (kz == 1 ? v.y : v.z));
}
// the same rotation as six selects per vertex, never a branch: lanes of one warp hold rays of different kz, and the
// compiler's if-conversion gives up on the three-component form above (it emitted ~13 instructions and a divergent
// region per vertex).  k0 = (kz == 0), k1 = (kz == 1).
__device__ __forceinline__ float sel_f(bool p, float a, float b) {
  float r;
  asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.f32 %0, %1, %2, q; }" : "=f"(r) : "f"(a), "f"(b), "r"((uint32_t)p));
  return r;
}
__device__ __forceinline__ float3 perm3_sel(const float3& v, bool k0, bool k1) {
  return f3(sel_f(k0, v.y, sel_f(k1, v.z, v.x)), sel_f(k0, v.z, sel_f(k1, v.x, v.y)), sel_f(k0, v.x, sel_f(k1, v.y, v.z)));
}

__device__ __forceinline__ RayPre ray_precompute(const float3& o, const float3& d) {
  RayPre r;
  r.o      = o;
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  r.kz     = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
  float3 p = perm3(d, r.kz);
  r.Sz     = 1.0f / p.z;
  r.Sx     = p.x * r.Sz;
  r.Sy     = p.y * r.Sz;
  return r;
}

// Returns true when the triangle is hit with tmin < t < tmax; (t, u, v) are only meaningful then.
__device__ __forceinline__ bool intersect_tri(const RayPre& r, const float3& v0, const float3& v1, const float3& v2,
                                              float tmin, float tmax, float& t_out, float& u_out, float& v_out) {
  const bool   k0 = r.kz == 0, k1 = r.kz == 1;
  const float3 A = perm3_sel(v0 - r.o, k0, k1), B = perm3_sel(v1 - r.o, k0, k1), C = perm3_sel(v2 - r.o, k0, k1);
  const float  Ax = fmaf(-r.Sx, A.z, A.x), Ay = fmaf(-r.Sy, A.z, A.y);
  const float  Bx = fmaf(-r.Sx, B.z, B.x), By = fmaf(-r.Sy, B.z, B.y);
  const float  Cx = fmaf(-r.Sx, C.z, C.x), Cy = fmaf(-r.Sy, C.z, C.y);
  // edge functions: explicit round-to-nearest mul/sub, never contracted to fma
  float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
  float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
  float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
    V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
    W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
  }
  // From here on no branch: in nearly every warp some lane passes the edge test, so the instructions below are issued anyway,
  // and every early `return` cost a divergence / reconvergence pair on top (5 % of k_pool's issue slots, ncu source view).
  // det == 0 needs no test of its own: it means U = V = W = 0 (all of one sign), then T = 0 and t = 0 * inf = NaN fails
  // both comparisons; with mixed signs the edge test fails.
  const bool  inside = !(fminf(fminf(U, V), W) < 0.0f && fmaxf(fmaxf(U, V), W) > 0.0f);
  const float det = U + V + W;
  const float T   = fmaf(W, r.Sz * C.z, fmaf(V, r.Sz * B.z, U * (r.Sz * A.z)));
  const float rcp = 1.0f / det;
  const float t   = T * rcp;
  t_out = t;
  u_out = V * rcp;
  v_out = W * rcp;
  return inside & (t > tmin) & (t < tmax);
}

// ------------------------------------------------------------------------------------------------
// Traversal stack: NS entries in shared memory, NL more in local memory
// ------------------------------------------------------------------------------------------------
// Set when a push finds the stack full (NS + NL entries): that subtree is then NOT traversed, so the host fails the
// call that ran the kernel (lisa_rt.cu: LISA_ERR_STATE) instead of returning an image with hits missing.  The builders do
// not bound the depth of the tree; a full stack takes a chain-shaped hierarchy of depth > NS + NL 8-wide levels.
static __device__ unsigned int g_trav_overflow = 0u;

template <typename T, int NS, int NL>
struct TravStack {
  T*  sm;      // this thread's column of the block's shared array
  int stride;  // = blockDim.x
  T   loc[NL];
  int sp;
  __device__ __forceinline__ TravStack(T* smem_base) : sm(smem_base + threadIdx.x), stride(blockDim.x), sp(0) {}
  __device__ __forceinline__ void push(const T& v) {
    if (sp < NS) sm[sp * stride] = v;
    else if (sp < NS + NL) loc[sp - NS] = v;
    else { g_trav_overflow = 1u; return; }  // full: the entry is dropped (pop stays in bounds) and the call fails
    sp++;
  }
  __device__ __forceinline__ T pop() {
    sp--;
    return sp < NS ? sm[sp * stride] : loc[sp - NS];
  }
  __device__ __forceinline__ bool empty() const { return sp == 0; }
  __device__ __forceinline__ void clear() { sp = 0; }
};

// 4 entries x 8 B per thread in shared memory (3 KB per 128-thread CTA).  Measured on B200: 12 -> 4 entries gives +2 %
// on the Cornell box (smaller carve-out, more L1 for nodes and triangles) and changes nothing on the 1M-triangle
// soup (29 nodes per ray), whose deeper levels spill to the local array.
#ifndef LISA_STACK_SM
#define LISA_STACK_SM 4
#endif
#ifndef LISA_STACK_TOTAL
#define LISA_STACK_TOTAL 64  // entries before a push overflows (the test build liblisa_rt_tinystack.so shrinks it to provoke that)
#endif
#define LISA_STACK_LOC (LISA_STACK_TOTAL - LISA_STACK_SM)
typedef TravStack<uint2, LISA_STACK_SM, LISA_STACK_LOC> Stack;
// bytes of dynamic shared memory a traversal kernel needs per thread
#define LISA_STACK_SMEM_PER_THREAD (LISA_STACK_SM * 8)

// |component| clamped to 1e-20: the wide node test scales the reciprocal by 2^(e+15), which has to stay finite for any sane scene
__device__ __forceinline__ float3 safe_rcp_dir(const float3& d) {
  const float eps = 1e-20f;
  return f3(1.0f / (fabsf(d.x) > eps ? d.x : copysignf(eps, d.x)), 1.0f / (fabsf(d.y) > eps ? d.y : copysignf(eps, d.y)),
            1.0f / (fabsf(d.z) > eps ? d.z : copysignf(eps, d.z)));
}

// ------------------------------------------------------------------------------------------------
// Node layouts
//
// Binary (ablation path).  Node = 4 x float4:
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   n3 = (child0, child1, -, -) as int bits
// child >= 0: node index; child < 0: leaf holding the single triangle ~child.
//
// Compressed 8-wide (Ylitie, Karras, Laine 2017).  Node = 5 x float4 (80 B):
//   w0 = (p.x, p.y, p.z, [ex | ey<<8 | ez<<16 | imask<<24])
//   w1 = (child_base, tri_base, meta[0..3], meta[4..7])
//   w2 = (qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7])
//   w3 = (qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7])
//   w4 = (qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7])
// meta[i]: 0 empty | internal: 0b001_11sss (sss = slot i) | leaf: unary triangle count in the top
// 3 bits (1: 001, 2: 011, 3: 111) and the offset of its first triangle from tri_base in the low 5.
// Child boxes: lo = p + qlo * 2^e, hi = p + qhi * 2^e per axis.  Traversal order: children are stored in octant
// slots; hit internal children get the bit 24 + (slot ^ octant of the ray), popped highest first.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
  // each byte with bit 7 set becomes 0xff, else 0x00
  uint32_t r;
  asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ uint32_t extract_byte(uint32_t x, uint32_t i) { return (x >> (i * 8)) & 0xffu; }

// ================================================================================================
// Step-wise traversal (used by every kernel: k_extend, k_rays and the diagnostic queries).
//
// A warp's lanes are at different points of different rays, so the kernel runs a state machine whose
// loop body is ONE work quantum per lane: at most one node visit followed by at most LISA_TRI_PER_STEP
// triangle tests.  Lanes therefore reconverge after every quantum instead of after every ray, which is
// what keeps SIMT efficiency up when rays need between 1 and 10 node visits.
// ================================================================================================
#ifndef LISA_TRI_PER_STEP
#define LISA_TRI_PER_STEP 2
#endif

struct StepRay {        // per-ray constants kept in registers
  float3   idir;        // safe reciprocal direction
  float    Sx, Sy, Sz;  // watertight shear (ray_precompute)
  int      kz;
  uint32_t oct_inv4;    // wide BVH: octant byte replicated x4; bit 2/1/0 set = dir.x/y/z >= 0
};

__device__ __forceinline__ StepRay step_ray(const float3& d) {
  StepRay r;
  r.idir = safe_rcp_dir(d);
  const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  r.kz = (ax > ay) ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
  const float3 p = perm3(d, r.kz);
  r.Sz = 1.0f / p.z;
  r.Sx = p.x * r.Sz;
  r.Sy = p.y * r.Sz;
  r.oct_inv4 = ((d.x < 0.0f ? 0u : 4u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 1u)) * 0x01010101u;
  return r;
}

__device__ __forceinline__ bool step_tri(const float3& o, const StepRay& r, const float4* __restrict__ tri_v, int ti, float tmin,
                                         float tmax, float& t_out) {
  const float4 a = __ldg(tri_v + 3 * ti), b = __ldg(tri_v + 3 * ti + 1), c = __ldg(tri_v + 3 * ti + 2);
  RayPre pre;
  pre.o = o; pre.kz = r.kz; pre.Sx = r.Sx; pre.Sy = r.Sy; pre.Sz = r.Sz;
  float u, v;
  return intersect_tri(pre, f3(a), f3(b), f3(c), tmin, tmax, t_out, u, v);
}

__device__ __forceinline__ bool step_tri_uv(const float3& o, const StepRay& r, const float4* __restrict__ tri_v, int ti, float tmin,
                                            float tmax, float& t_out, float& u_out, float& v_out) {
  const float4 a = __ldg(tri_v + 3 * ti), b = __ldg(tri_v + 3 * ti + 1), c = __ldg(tri_v + 3 * ti + 2);
  RayPre pre;
  pre.o = o; pre.kz = r.kz; pre.Sx = r.Sx; pre.Sy = r.Sy; pre.Sz = r.Sz;
  return intersect_tri(pre, f3(a), f3(b), f3(c), tmin, tmax, t_out, u_out, v_out);
}

// byte j of q -> the float 1 + q * 2^-15 in ONE instruction: PRMT drops the byte into bits 8..15 of the mantissa of 1.0f.
// The plane equation q * a + o is then evaluated as fma(that, 2^15 a, o - 2^15 a): no int->float conversion (slow
// pipe) and no subtraction of the magic.  The 1.0f comes from constant memory so that the SELECTOR is the immediate
// operand of PRMT; with both operands literal, ptxas keeps the magic as the immediate and spends a MOV per PRMT on the
// selector (48 per node visit).
static __constant__ uint32_t c_unit_magic = 0x3F800000u;
__device__ __forceinline__ float byte_to_unit(uint32_t q, int j, uint32_t magic) {
  return __uint_as_float(__byte_perm(q, magic, 0x7604u | ((uint32_t)j << 4)));
}

// ---- compressed 8-wide --------------------------------------------------------------------------
struct WideState {
  uint2 ng, tg;  // node group (child base | hit bits + imask), triangle group (tri base | hit bits)
  __device__ __forceinline__ void begin(int root) { ng = make_uint2((uint32_t)root, root >= 0 ? 0x80000000u : 0u); tg = make_uint2(0u, 0u); }
  __device__ __forceinline__ bool has_nodes() const { return (ng.y & 0xff000000u) != 0u; }
  __device__ __forceinline__ bool has_tris() const { return tg.y != 0u; }
};

// one node visit; precondition st.has_nodes() && !st.has_tris()
template <class STK>
__device__ __forceinline__ void wide_node_step(const float4* __restrict__ nodes, const float3& o, const StepRay& r, float tmin,
                                               float tlimit, WideState& st, STK& stack) {
  const uint32_t hits_imask      = st.ng.y;
  const uint32_t child_bit_index = 31u - __clz(hits_imask);
  const uint32_t child_base      = st.ng.x;
  st.ng.y &= ~(1u << child_bit_index);
  if (st.ng.y & 0xff000000u) stack.push(st.ng);
  const uint32_t slot_index     = (child_bit_index - 24u) ^ (r.oct_inv4 & 0xffu);
  const uint32_t relative_index = __popc(hits_imask & ~(0xffffffffu << slot_index) & 0xffu);
  const uint32_t ni             = child_base + relative_index;
  const float4 w0 = __ldg(nodes + 5 * ni), w1 = __ldg(nodes + 5 * ni + 1), w2 = __ldg(nodes + 5 * ni + 2),
               w3 = __ldg(nodes + 5 * ni + 3), w4 = __ldg(nodes + 5 * ni + 4);
  const uint32_t eimask = __float_as_uint(w0.w);
  const float3 adir = f3(__uint_as_float((eimask & 0xffu) << 23) * r.idir.x, __uint_as_float(((eimask >> 8) & 0xffu) << 23) * r.idir.y,
                         __uint_as_float(((eimask >> 16) & 0xffu) << 23) * r.idir.z);
  const float3 org  = f3((w0.x - o.x) * r.idir.x, (w0.y - o.y) * r.idir.y, (w0.z - o.z) * r.idir.z);
  // conservative planes.  With u = 1 + q 2^-15 (byte_to_unit) the plane is t = u * A + (org - A), A = 2^15 adir: the fma
  // is exact up to its final rounding, org - A carries an error of 2^-24 max(|org|, 2^15 |adir|), and org itself two
  // roundings; |t| <= |org| + 255 |adir|.  One absolute pad of 1.2e-6 (|org| + 3072 |adir|) covers all of it
  // (3072 * 1.2e-6 = 3.7e-3 of a quantisation step, 2^-9 = 1.95e-3 needed).
  const float  kpad = 1.2e-6f;
  const float3 pad  = f3(kpad * fmaf(3072.0f, fabsf(adir.x), fabsf(org.x)), kpad * fmaf(3072.0f, fabsf(adir.y), fabsf(org.y)),
                         kpad * fmaf(3072.0f, fabsf(adir.z), fabsf(org.z)));
  const float3 an = f3(adir.x * 32768.0f, adir.y * 32768.0f, adir.z * 32768.0f), af = an;
  const float3 ob = org - an;
  const float3 on = ob - pad, of = ob + pad;
  const uint32_t magic = c_unit_magic;
  st.ng.x = __float_as_uint(w1.x);
  st.tg.x = __float_as_uint(w1.y);
  const bool negx = !(r.oct_inv4 & 4u), negy = !(r.oct_inv4 & 2u), negz = !(r.oct_inv4 & 1u);
  uint32_t hitmask = 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const uint32_t meta4 = __float_as_uint(half == 0 ? w1.z : w1.w);
    if (meta4 == 0u) continue;  // four empty slots
    const uint32_t is_inner4   = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4  = (meta4 ^ (r.oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t qlox = __float_as_uint(half == 0 ? w2.x : w2.y), qloy = __float_as_uint(half == 0 ? w2.z : w2.w);
    const uint32_t qloz = __float_as_uint(half == 0 ? w3.x : w3.y), qhix = __float_as_uint(half == 0 ? w3.z : w3.w);
    const uint32_t qhiy = __float_as_uint(half == 0 ? w4.x : w4.y), qhiz = __float_as_uint(half == 0 ? w4.z : w4.w);
    const uint32_t nx = negx ? qhix : qlox, fx = negx ? qlox : qhix;
    const uint32_t ny = negy ? qhiy : qloy, fy = negy ? qloy : qhiy;
    const uint32_t nz = negz ? qhiz : qloz, fz = negz ? qloz : qhiz;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float tnx = fmaf(byte_to_unit(nx, j, magic), an.x, on.x), tfx = fmaf(byte_to_unit(fx, j, magic), af.x, of.x);
      const float tny = fmaf(byte_to_unit(ny, j, magic), an.y, on.y), tfy = fmaf(byte_to_unit(fy, j, magic), af.y, of.y);
      const float tnz = fmaf(byte_to_unit(nz, j, magic), an.z, on.z), tfz = fmaf(byte_to_unit(fz, j, magic), af.z, of.z);
      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
      const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tlimit));
      if (tn <= tf) hitmask |= extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
    }
  }
  st.ng.y = (hitmask & 0xff000000u) | (eimask >> 24);
  st.tg.y = hitmask & 0x00ffffffu;
}

// ---- binary ---------------------------------------------------------------------------------------
#define LISA_BIN_NONE 0x7fffffff
struct BinState {
  int cur;  // node index >= 0, leaf ~tri < 0, LISA_BIN_NONE when nothing is pending
  __device__ __forceinline__ void begin(int root) { cur = root >= 0 ? root : LISA_BIN_NONE; }
  __device__ __forceinline__ bool has_nodes() const { return cur >= 0 && cur != LISA_BIN_NONE; }
  __device__ __forceinline__ bool has_tris() const { return cur < 0; }
};

// one node visit; precondition st.has_nodes().  Leaves st.cur = next node / leaf / NONE (after a pop).
template <class STK>
__device__ __forceinline__ void bin_node_step(const float4* __restrict__ nodes, const float3& o, const StepRay& r, float tmin,
                                              float tlimit, BinState& st, STK& stack) {
  const float4 n0 = __ldg(nodes + 4 * st.cur), n1 = __ldg(nodes + 4 * st.cur + 1), n2 = __ldg(nodes + 4 * st.cur + 2),
               n3 = __ldg(nodes + 4 * st.cur + 3);
  // (plane - o) * idir: one rounding each, so a relative pad is enough
  const float a0x = (n0.x - o.x) * r.idir.x, b0x = (n0.y - o.x) * r.idir.x, a0y = (n0.z - o.y) * r.idir.y, b0y = (n0.w - o.y) * r.idir.y;
  const float a0z = (n2.x - o.z) * r.idir.z, b0z = (n2.y - o.z) * r.idir.z;
  const float a1x = (n1.x - o.x) * r.idir.x, b1x = (n1.y - o.x) * r.idir.x, a1y = (n1.z - o.y) * r.idir.y, b1y = (n1.w - o.y) * r.idir.y;
  const float a1z = (n2.z - o.z) * r.idir.z, b1z = (n2.w - o.z) * r.idir.z;
  const float t0n = fmaxf(fmaxf(fminf(a0x, b0x), fminf(a0y, b0y)), fminf(a0z, b0z)) * 0.999999f;
  const float t0f = fminf(fminf(fmaxf(a0x, b0x), fmaxf(a0y, b0y)), fmaxf(a0z, b0z)) * 1.000001f;
  const float t1n = fmaxf(fmaxf(fminf(a1x, b1x), fminf(a1y, b1y)), fminf(a1z, b1z)) * 0.999999f;
  const float t1f = fminf(fminf(fmaxf(a1x, b1x), fmaxf(a1y, b1y)), fmaxf(a1z, b1z)) * 1.000001f;
  const bool  h0 = fmaxf(t0n, tmin) <= fminf(t0f, tlimit), h1 = fmaxf(t1n, tmin) <= fminf(t1f, tlimit);
  int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
  if (h0 && h1) {
    if (t1n < t0n) { int t = c0; c0 = c1; c1 = t; }
    stack.push(make_uint2((uint32_t)c1, 0u));
    st.cur = c0;
  } else if (h0) st.cur = c0;
  else if (h1) st.cur = c1;
  else st.cur = stack.empty() ? LISA_BIN_NONE : (int)stack.pop().x;
}


// One whole traversal of one ray with the step functions above (the diagnostic queries use this; the render kernels
// interleave the same steps across lanes).  ANY = stop at the first hit.
template <bool WIDE, bool ANY>
__device__ __forceinline__ bool trace_steps(const float4* __restrict__ nodes, const float4* __restrict__ tri_v, int root,
                                            const float3& o, const float3& d, float tmin, float tmax, Hit& hit, Stack& stack,
                                            uint32_t& n_nodes, uint32_t& n_tris) {
  hit.prim = -1;
  hit.t    = tmax;
  if (root < 0) return false;
  const StepRay r = step_ray(d);
  stack.clear();
  if (WIDE) {
    WideState st;
    st.begin(root);
    while (true) {
      if (st.has_nodes() && !st.has_tris()) { n_nodes++; wide_node_step(nodes, o, r, tmin, hit.t, st, stack); }
      while (st.tg.y) {
        const uint32_t b = __ffs(st.tg.y) - 1u;
        st.tg.y &= st.tg.y - 1u;
        const int ti = (int)(st.tg.x + b);
        float t, u, v;
        n_tris++;
        if (step_tri_uv(o, r, tri_v, ti, tmin, hit.t, t, u, v)) { hit.t = t; hit.u = u; hit.v = v; hit.prim = ti; if (ANY) return true; }
      }
      if (!st.has_nodes()) {
        if (stack.empty()) break;
        st.ng = stack.pop();
      }
    }
  } else {
    BinState st;
    st.begin(root);
    while (st.cur != LISA_BIN_NONE) {
      if (st.has_nodes()) { n_nodes++; bin_node_step(nodes, o, r, tmin, hit.t, st, stack); }
      if (st.has_tris()) {
        const int ti = ~st.cur;
        float t, u, v;
        n_tris++;
        if (step_tri_uv(o, r, tri_v, ti, tmin, hit.t, t, u, v)) { hit.t = t; hit.u = u; hit.v = v; hit.prim = ti; if (ANY) return true; }
        st.cur = stack.empty() ? LISA_BIN_NONE : (int)stack.pop().x;
      }
    }
  }
  return hit.prim >= 0;
}

}  // namespace lisa
