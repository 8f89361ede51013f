// pg_sphere_exact.cuh -- exact sphere-sphere / sphere-plane visits on scalar state, shared by the
// batched-world kernel (pg_spheres.cu) and the large-world pipeline (pg_large.cu).
#pragma once

#include "pg_math.cuh"

namespace pgs {

using pgm::V3;
using pgm::v3;

struct Sphere {          // one body's hot state as scalars (register resident)
  float px, py, pz, vx, vy, vz, r;
  int rest;
};

__device__ __forceinline__ float reach(float r, float vx, float vy, float vz, float dt) {
  // >= r + |v| dt with a 1e-4 relative cushion; FMA is fine, this is only a filter
  float sp = sqrtf(__fmaf_rn(vx, vx, __fmaf_rn(vy, vy, vz * vz)));
  return __fmaf_rn(sp, dt, r) * 1.0001f + 1.0e-6f;
}

// Conservative gate in front of the exact sphere-sphere predicate.  Closest approach of the two
// linear paths over [0,dt]: a leaf interval of interval_collision (len < 1e-6 s) passes only if the
// COMPUTED distance is <= (nA+nB)*len < 1e-6 (nA+nB) at both ends, and computed and real distances
// differ by less than 4e-6 x the coordinate magnitude (see pgm::ss_interval).  If the real minimum
// gap exceeds delta, no leaf passes and the predicate is false.  `vsum` >= |vA| + |vB|.
static __device__ __forceinline__ bool pair_gate_rejects(V3 cA, V3 vA, V3 cB, V3 vB, float rs, float vsum, float dt) {
  const V3 s = pgm::sub(cA, cB), v = pgm::sub(vA, vB);
  const float vv = __fmaf_rn(v.x, v.x, __fmaf_rn(v.y, v.y, v.z * v.z));
  const float sv = __fmaf_rn(s.x, v.x, __fmaf_rn(s.y, v.y, s.z * v.z));
  const float t = vv > 0.0f ? fminf(fmaxf(-sv / vv, 0.0f), dt) : 0.0f;
  const float qx = __fmaf_rn(v.x, t, s.x), qy = __fmaf_rn(v.y, t, s.y), qz = __fmaf_rn(v.z, t, s.z);
  const float dmin2 = __fmaf_rn(qx, qx, __fmaf_rn(qy, qy, qz * qz));
  const float cmax = fmaxf(fmaxf(fmaxf(fabsf(cA.x), fabsf(cA.y)), fabsf(cA.z)),
                           fmaxf(fmaxf(fabsf(cB.x), fabsf(cB.y)), fabsf(cB.z)));
  const float delta = 1.1e-6f * vsum + 4.0e-6f * (cmax + vsum * dt) + 1.0e-6f;
  const float lim = (rs + delta) * 1.00001f;
  return dmin2 > lim * lim && dmin2 < 3.0e38f;
}

// Reach filter + closest-approach gate of bodies r and c of a shared-memory world image
// (P4 = position + reach, V4 = velocity + radius); `integrate_c`: body c as it will be once its own
// row has integrated it.  True = the pair cannot fire.  One out-of-line copy for every caller.
static __device__ __noinline__ bool pair_gate_far(const float4 *P4, const float4 *V4, int r, int c, bool integrate_c, float dt) {
  const float4 ap = P4[r], av = V4[r];
  const float4 bp = P4[c], bv = V4[c];
  V3 p = v3(bp.x, bp.y, bp.z), vel = v3(bv.x, bv.y, bv.z);
  float rho = bp.w;
  const float na2 = __fmaf_rn(av.x, av.x, __fmaf_rn(av.y, av.y, av.z * av.z));
  float nb2 = __fmaf_rn(vel.x, vel.x, __fmaf_rn(vel.y, vel.y, vel.z * vel.z));
  if (integrate_c) {
    pgm::integrate(p, vel, dt);
    nb2 = __fmaf_rn(vel.x, vel.x, __fmaf_rn(vel.y, vel.y, vel.z * vel.z));
    rho = __fmaf_rn(sqrtf(nb2), dt, bv.w) * 1.0001f + 1.0e-6f;         // == reach()
  }
  const float dx = ap.x - p.x, dy = ap.y - p.y, dz = ap.z - p.z;
  const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz));
  const float T = ap.w + rho;
  if (!(d2 <= T * T)) return true;                                     // root test of the interval search fails
  if (dt <= 1.0e-9f) return false;
  const float vsum = sqrtf(2.0f * (na2 + nb2)) * 1.0001f;              // >= |vA| + |vB|
  return pair_gate_rejects(v3(ap.x, ap.y, ap.z), v3(av.x, av.y, av.z), p, vel, av.w + bv.w, vsum, dt);
}

// Exact visit of two spheres (world.c:481-541 with the sphere-sphere table entries).
// All lanes call it with identical arguments.  Returns true if an event fires; A and B updated.
// Gate = false: the caller already applied pair_gate_rejects to this very state.
template <bool Gate, int Coop>
static __device__ __forceinline__ bool sphere_pair_exact_impl(Sphere &A, Sphere &B, float dt, pgm::LeafLen leaf) {
  V3 cA = v3(0.0f + A.px, 0.0f + A.py, 0.0f + A.pz);     // centre (0,0,0) + position, distance.c:283
  V3 cB = v3(0.0f + B.px, 0.0f + B.py, 0.0f + B.pz);
  V3 vA = v3(A.vx, A.vy, A.vz), vB = v3(B.vx, B.vy, B.vz);
  float rs = A.r + B.r;
  float nA = pgm::norm(vA), nB = pgm::norm(vB);
  if (Gate && pair_gate_rejects(cA, vA, cB, vB, rs, nA + nB, dt)) return false;
  if (!pgm::ss_interval<Coop>(cA, vA, nA, cB, vB, nB, rs, 0.0f, dt, leaf)) return false;
  pgm::Contact c = pgm::ss_narrow(cA, vA, cB, vB, rs);
  if (!(c.hit && c.t >= 0)) return false;
  V3 pA = v3(A.px, A.py, A.pz), pB = v3(B.px, B.py, B.pz);
  pgm::ss_resolve(pA, vA, v3(0.0f, 0.0f, 0.0f), pB, vB, v3(0.0f, 0.0f, 0.0f), rs, c.t);
  A.px = pA.x; A.py = pA.y; A.pz = pA.z; A.vx = vA.x; A.vy = vA.y; A.vz = vA.z;
  B.px = pB.x; B.py = pB.y; B.pz = pB.z; B.vx = vB.x; B.vy = vB.y; B.vz = vB.z;
  return true;
}
static __device__ __noinline__ bool sphere_pair_exact(Sphere &A, Sphere &B, float dt, pgm::LeafLen leaf) {
  return sphere_pair_exact_impl<true, 0>(A, B, dt, leaf);
}
// Warp-uniform variant (all 32 lanes, identical arguments, already gated): the lanes share the leaf scan.
static __device__ __noinline__ bool sphere_pair_exact_gated(Sphere &A, Sphere &B, float dt, pgm::LeafLen leaf) {
  return sphere_pair_exact_impl<false, 1>(A, B, dt, leaf);
}

// One thread per pair with a bounded walk: 0 = no event, 1 = event (A, B updated), 2 = the interval
// search is one of the rare long ones (a grazing pair): nothing changed, evaluate the visit with a
// whole block instead (sphere_pair_exact_block).
static __device__ __noinline__ int sphere_pair_exact_try(Sphere &A, Sphere &B, float dt, pgm::LeafLen leaf) {
  V3 cA = v3(0.0f + A.px, 0.0f + A.py, 0.0f + A.pz);
  V3 cB = v3(0.0f + B.px, 0.0f + B.py, 0.0f + B.pz);
  V3 vA = v3(A.vx, A.vy, A.vz), vB = v3(B.vx, B.vy, B.vz);
  float rs = A.r + B.r;
  float nA = pgm::norm(vA), nB = pgm::norm(vB);
  if (pair_gate_rejects(cA, vA, cB, vB, rs, nA + nB, dt)) return 0;
  const int r = pgm::ss_interval_impl<0, true>(cA, vA, nA, cB, vB, nB, rs, 0.0f, dt, leaf);
  if (r != pgm::kWalkTrue) return r == pgm::kWalkGaveUp ? 2 : 0;
  pgm::Contact c = pgm::ss_narrow(cA, vA, cB, vB, rs);
  if (!(c.hit && c.t >= 0)) return 0;
  V3 pA = v3(A.px, A.py, A.pz), pB = v3(B.px, B.py, B.pz);
  pgm::ss_resolve(pA, vA, v3(0.0f, 0.0f, 0.0f), pB, vB, v3(0.0f, 0.0f, 0.0f), rs, c.t);
  A.px = pA.x; A.py = pA.y; A.pz = pA.z; A.vx = vA.x; A.vy = vA.y; A.vz = vA.z;
  B.px = pB.x; B.py = pB.y; B.pz = pB.z; B.vx = vB.x; B.vy = vB.y; B.vz = vB.z;
  return 1;
}
// All threads of the block, identical arguments (gate included): they share the leaf scan.
static __device__ __noinline__ bool sphere_pair_exact_block(Sphere &A, Sphere &B, float dt, pgm::LeafLen leaf) {
  return sphere_pair_exact_impl<true, 2>(A, B, dt, leaf);
}

// Exact visit sphere x plane (per lane, divergent).  Returns true if an event fires.
static __device__ __noinline__ bool sphere_plane_exact(Sphere &A, float4 pl, float restitution, float dt, pgm::LeafLen leaf) {
  V3 c = v3(0.0f + A.px, 0.0f + A.py, 0.0f + A.pz);
  V3 v = v3(A.vx, A.vy, A.vz);
  V3 n = v3(pl.x, pl.y, pl.z);
  float nv = pgm::norm(v);
  if (!pgm::sp_interval(c, v, nv, n, pl.w, A.r, 0.0f, dt, leaf)) return false;
  pgm::Contact k = pgm::sp_narrow(c, v, n, pl.w, A.r, dt);
  if (!(k.hit && k.t >= 0)) return false;
  V3 p = v3(A.px, A.py, A.pz);
  bool rest = pgm::sp_resolve(p, v, v3(0.0f, 0.0f, 0.0f), A.r, restitution, n, pl.w, k.t);
  A.px = p.x; A.py = p.y; A.pz = p.z; A.vx = v.x; A.vy = v.y; A.vz = v.z;
  if (rest) A.rest = 1;
  return true;
}


}  // namespace pgs
