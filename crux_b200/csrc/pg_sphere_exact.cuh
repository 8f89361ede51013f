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

// Exact visit of two spheres (world.c:481-541 with the sphere-sphere table entries).
// All lanes call it with identical arguments.  Returns true if an event fires; A and B updated.
static __device__ __noinline__ bool sphere_pair_exact(Sphere &A, Sphere &B, float dt, pgm::LeafLen leaf) {
  V3 cA = v3(0.0f + A.px, 0.0f + A.py, 0.0f + A.pz);     // centre (0,0,0) + position, distance.c:283
  V3 cB = v3(0.0f + B.px, 0.0f + B.py, 0.0f + B.pz);
  V3 vA = v3(A.vx, A.vy, A.vz), vB = v3(B.vx, B.vy, B.vz);
  float rs = A.r + B.r;
  float nA = pgm::norm(vA), nB = pgm::norm(vB);
  if (pair_gate_rejects(cA, vA, cB, vB, rs, nA + nB, dt)) return false;
  if (!pgm::ss_interval(cA, vA, nA, cB, vB, nB, rs, 0.0f, dt, leaf)) return false;
  pgm::Contact c = pgm::ss_narrow(cA, vA, cB, vB, rs);
  if (!(c.hit && c.t >= 0)) return false;
  V3 pA = v3(A.px, A.py, A.pz), pB = v3(B.px, B.py, B.pz);
  pgm::ss_resolve(pA, vA, v3(0.0f, 0.0f, 0.0f), pB, vB, v3(0.0f, 0.0f, 0.0f), rs, c.t);
  A.px = pA.x; A.py = pA.y; A.pz = pA.z; A.vx = vA.x; A.vy = vA.y; A.vz = vA.z;
  B.px = pB.x; B.py = pB.y; B.pz = pB.z; B.vx = vB.x; B.vy = vB.y; B.vz = vB.z;
  return true;
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
