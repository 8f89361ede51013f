// pg_math.cuh -- exact per-pair arithmetic of the Crux physics step, device side.
//
// Every function here must produce the SAME BITS as the reference's C built by gcc for x86-64
// (SSE2 scalar math, no FMA contraction).  Rules that make that true on sm_100a:
//   * this header is only ever compiled with  -fmad=false  (no FFMA/DFMA contraction), default
//     -prec-div=true -prec-sqrt=true -ftz=false, never --use_fast_math;
//   * products are summed left to right exactly as cglm's scalar vec3/mat4 routines do;
//   * wherever the reference's C promotes to double (a double literal such as 0.001 / 0.0001 / 0.5,
//     or a <math.h> double function: sqrt, fabs) the same operation is done in double here [f64];
//   * sinf/cosf are never evaluated on the device: the two Euler matrices the path needs are
//     computed by the host with libm (PgBody.euler_raw / euler_node, include/physics_gpu.h).
// Explicit __fmaf_rn / float-only shortcuts appear ONLY in conservative filters whose outcome is
// proven to equal the exact predicate (see interval tests below); anything inside the proven band
// falls through to the exact [f64] evaluation.
//
// Reference citations are relative to /root/reference/src/physics/.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/physics_gpu.h"

namespace pgm {

constexpr float kGravity = 9.8f;            // world.c:363,550; resolution.c:28,159,314,421,472,583,633
constexpr double kMaxDt = 0.01666;          // world.c:18 (double)
constexpr float kIntervalEps = 0.000001f;   // world.c:21
constexpr double kNpEps = 0.0001;           // narrow_phase.c:6 (double)

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 v3p(const float *p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void store(float *p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ float get(const V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__device__ __forceinline__ void set(V3 &a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

// ---- cglm scalar forms (vec3.h) ----
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float norm(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 muladds(V3 a, float s, V3 d) { return v3(d.x + a.x * s, d.y + a.y * s, d.z + a.z * s); }
__device__ __forceinline__ V3 mulsubs(V3 a, float s, V3 d) { return v3(d.x - a.x * s, d.y - a.y * s, d.z - a.z * s); }
__device__ __forceinline__ V3 normalize(V3 a) {
  float n = norm(a);
  if (n < FLT_EPSILON) return v3(0.0f, 0.0f, 0.0f);
  return scale(a, 1.0f / n);
}
__device__ __forceinline__ float glm_max(float a, float b) { if (a > b) return a; return b; }
__device__ __forceinline__ float glm_min(float a, float b) { if (a < b) return a; return b; }
__device__ __forceinline__ float glm_clamp(float v, float lo, float hi) { return glm_min(glm_max(v, lo), hi); }

// glm_mat4_mulv3(m, v, 1.0f): column-major m[c*4+r]
__device__ __forceinline__ V3 xform_point(const float *m, V3 v) {
  V3 r;
  r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * 1.0f;
  r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * 1.0f;
  r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * 1.0f;
  return r;
}
// glm_mat3_mulv, column-major m[c*3+r]
__device__ __forceinline__ V3 mat3_mulv(const float *m, V3 v) {
  V3 r;
  r.x = m[0] * v.x + m[3] * v.y + m[6] * v.z;
  r.y = m[1] * v.x + m[4] * v.y + m[7] * v.z;
  r.z = m[2] * v.x + m[5] * v.y + m[8] * v.z;
  return r;
}

// World-space decomposition every AABB path repeats (distance.c:27-35): translation, per-column
// norms, upper 3x3 divided by scale[0].
struct NodeFrame { V3 pos; V3 scl; float rot[9]; };
__device__ inline void node_decompose(const float *m, NodeFrame &f) {
  f.pos = xform_point(m, v3(0.0f, 0.0f, 0.0f));
  f.scl.x = norm(v3(m[0], m[1], m[2]));
  f.scl.y = norm(v3(m[4], m[5], m[6]));
  f.scl.z = norm(v3(m[8], m[9], m[10]));
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) f.rot[c * 3 + r] = m[c * 4 + r];
  if (f.scl.x != 0.0f) {
    float s = 1.0f / f.scl.x;
#pragma unroll
    for (int k = 0; k < 9; k++) f.rot[k] *= s;
  }
}

struct Box { V3 c, e; };
struct Ball { V3 c; float r; };
struct Pill { V3 a, b; float r; };
struct Flat { V3 n; float d; };

// AABB_update, aabb.c:59-78.  [f64]: fabs() products are double, accumulated into a float.
__device__ inline Box aabb_update(const float *shape, const float *rot, V3 tr, V3 scl) {
  Box o;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float c = get(tr, i);
    float e = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      c += rot[j * 3 + i] * shape[j];
      e = (float)((double)e + fabs((double)rot[j * 3 + i]) * (double)shape[3 + j] * fabs((double)get(scl, i)));
    }
    set(o.c, i, c);
    set(o.e, i, e);
  }
  return o;
}
__device__ __forceinline__ Box zero_box() { Box b; b.c = v3(0, 0, 0); b.e = v3(0, 0, 0); return b; }

// Box through the scene-node transform, advanced by adv*time (distance.c:26-37; resolution.c:41-52).
__device__ inline Box box_from_node(const PgBody &b, bool advance, V3 adv, float time) {
  if (!b.has_node) return zero_box();
  NodeFrame f; node_decompose(b.node_xform, f);
  if (advance) f.pos = muladds(adv, time, f.pos);
  return aabb_update(b.shape, f.rot, f.pos, f.scl);
}
// Box from body rotation (degrees read as radians) / position / scale: distance.c:253-264.
__device__ inline Box box_from_body(const PgBody &b, V3 pos) {
  return aabb_update(b.shape, b.euler_raw, pos, v3p(b.scale));
}
__device__ __forceinline__ Ball ball_from_body(const PgBody &b, V3 pos) {   // distance.c:282-285
  Ball s; s.c = add(v3p(b.shape), pos); s.r = b.shape[3] * b.scale[0]; return s;
}
__device__ inline Ball ball_from_node(const PgBody &b, bool advance, V3 adv, float time) {   // distance.c:113-125
  Ball s; s.c = v3(0, 0, 0); s.r = 0.0f;
  if (!b.has_node) return s;
  NodeFrame f; node_decompose(b.node_xform, f);
  if (advance) f.pos = muladds(adv, time, f.pos);
  s.c = add(v3p(b.shape), f.pos);
  s.r = b.shape[3] * f.scl.x;
  return s;
}
__device__ inline Pill pill_from_node(const PgBody &b) {   // distance.c:185-189,209
  Pill p; p.a = v3(0, 0, 0); p.b = v3(0, 0, 0);
  if (b.has_node) {
    p.a = xform_point(b.node_xform, v3p(b.shape));
    p.b = xform_point(b.node_xform, v3p(b.shape + 3));
  }
  p.r = b.shape[6] * b.scale[0];
  return p;
}
__device__ inline Pill pill_from_body(const PgBody &b) {   // distance.c:366-379
  Pill p;
  p.a = scale(v3p(b.shape), b.scale[0]);
  p.b = scale(v3p(b.shape + 3), b.scale[0]);
  p.a = mat3_mulv(b.euler_raw, p.a);
  p.b = mat3_mulv(b.euler_raw, p.b);
  p.a = add(p.a, v3p(b.position));
  p.b = add(p.b, v3p(b.position));
  p.r = b.shape[6] * b.scale[0];
  return p;
}
// Plane as a capsule sees it: rotated by the PARENT node (distance.c:342-364).
__device__ inline Flat flat_for_pill(const PgBody &b) {
  Flat f; f.n = v3(0, 0, 0); f.d = 0.0f;
  if (!b.has_node) return f;
  NodeFrame nf; node_decompose(b.parent_xform, nf);
  f.n = normalize(mat3_mulv(nf.rot, v3p(b.shape)));
  f.d = b.shape[3] + dot(nf.pos, f.n);
  return f;
}
__device__ inline V3 pill_closest_to_flat(const Pill &p, const Flat &f) {   // distance.c:391-400
  float n_dot_a = dot(f.n, p.a);
  V3 seg = sub(p.b, p.a);
  float n_dot_seg = dot(f.n, seg);
  float t = (f.d - n_dot_a) / n_dot_seg;
  if (t <= 0) return p.a;
  if (t >= 1) return p.b;
  V3 v = sub(p.b, p.a);                       // glm_vec3_lerp: from + t*(to-from)
  v = v3(t * v.x, t * v.y, t * v.z);
  return add(p.a, v);
}
__device__ inline V3 pill_box_pq(const Pill &p, const Box &bx) {   // distance.c:211-238
  V3 seg = sub(p.b, p.a), a2c = sub(bx.c, p.a);
  float proj = dot(seg, a2c);
  float t = proj / dot(seg, seg);
  t = glm_clamp(t, 0.0f, 1.0f);
  V3 cp = muladds(seg, t, p.a);
  V3 q;
  q.x = glm_clamp(cp.x, bx.c.x - bx.e.x, bx.c.x + bx.e.x);
  q.y = glm_clamp(cp.y, bx.c.y - bx.e.y, bx.c.y + bx.e.y);
  q.z = glm_clamp(cp.z, bx.c.z - bx.e.z, bx.c.z + bx.e.z);
  return sub(q, cp);
}
// (float)sqrt((double)x) for a float x, as distance.c:91 computes it.  A square root evaluated with q >= 2p + 2 bits and
// then rounded to p bits equals the correctly rounded p-bit square root (double rounding is innocuous for sqrt: Figueroa,
// "When is double rounding innocuous?", 1995); here p = 24, q = 53 >= 50, and sqrtf is the correctly rounded IEEE root
// (-prec-sqrt=true, -ftz=false).  Same bits, a fraction of the instructions of the binary64 root.
__device__ __forceinline__ float sqrt_f64_as_f32(float x) { return sqrtf(x); }
__device__ __forceinline__ float proj_radius_f64(V3 e, V3 n) {   // distance.c:267-270 [f64]
  return (float)((double)e.x * fabs((double)n.x) + (double)e.y * fabs((double)n.y) + (double)e.z * fabs((double)n.z));
}

// =====================================================================================
// Sphere-only scalar kernels: the hot path of configs C2-C4.  The general table below calls the
// same functions, so both paths share one definition.
// =====================================================================================

// min_dist_at_time_sphere_sphere tail, distance.c:293-298.  d2 = |cA - cB|^2, rs = rA + rB.
__device__ __forceinline__ float ss_dist_from_d2(float d2, float rs) {
  return d2 < (rs * rs) ? 0.0f : (float)(sqrt((double)d2) - (double)rs);   // [f64]
}
// Does `ss_dist_from_d2(d2, rs) > mm` hold?  Exact, but decided in float whenever d2 is outside a
// 1e-6-relative band around (rs+mm)^2:
//   d2 < T2(1-4e-7)  =>  sqrt(d2)-rs < mm               => fl32(.) <= mm  => false
//   d2 > T2(1+1e-6)  =>  sqrt(d2)-rs > mm(1+4e-7)+...   => fl32(.) >  mm  => true
// with T2 = fl(fl(rs+mm)^2), whose relative error is < 1.8e-7.  Non-finite inputs take the exact path.
__device__ __forceinline__ bool ss_dist_exceeds(float d2, float rs, float mm) {
  float T = rs + mm;
  float T2 = T * T;
  if (d2 < T2 * 0.9999996f) return false;
  if (d2 > T2 * 1.000001f && d2 < 3.0e38f) return true;
  return ss_dist_from_d2(d2, rs) > mm;
}
__device__ __forceinline__ float ss_d2_at(V3 cA, V3 vA, V3 cB, V3 vB, float t) {   // distance.c:283-295
  V3 a = muladds(vA, t, cA), b = muladds(vB, t, cB);
  V3 d = sub(a, b);
  return dot(d, d);
}


// Depth-first walk of interval_collision's recursion tree (world.c:557-582) WITHOUT a stack.
// A node is (depth, path): bit k of `path` (k = depth-1 .. 0, root side first) says whether the
// walk went to the right child at that level.  End points of any node are re-derived from the
// root by the very same sequence of (t0+t1)*0.5f operations the recursion performs, so they are
// bit-identical to the recursive ones.  `probe(t0, t1)` evaluates a node and answers
//   kDead  : an end-point test fails, or the subtree provably holds no surviving leaf
//   kOpen  : both end-point tests pass, descend (a leaf that is kOpen survives)
//   kSure  : every node of the subtree provably passes, so a surviving leaf exists
// The walk visits children left first and returns true at the first surviving leaf; since only the
// boolean is consumed by physics_step (hit_time is never read, world.c:196-198), proofs that skip
// or short-cut a subtree leave the result unchanged.
enum { kDead = 0, kOpen = 1, kSure = 2 };

template <typename Probe>
__device__ __forceinline__ bool interval_walk(float t_begin, float t_end, float *hit, Probe probe) {
  unsigned long long path = 0;
  int depth = 0;
  float t0 = t_begin, t1 = t_end;
  for (;;) {
    const int st = probe(t0, t1);
    if (st == kSure) { if (hit) *hit = t0; return true; }
    if (st == kOpen) {
      if (t1 - t0 < kIntervalEps) { if (hit) *hit = t0; return true; }
      if (depth >= 62) return false;                 // unreachable for finite end points
      path <<= 1; depth++;
      t1 = (t0 + t1) * 0.5f;
      continue;
    }
    while (depth > 0 && (path & 1ull)) { path >>= 1; depth--; }   // right children are finished
    if (depth == 0) return false;
    path |= 1ull;                                    // left child failed: take its right sibling
    t0 = t_begin; t1 = t_end;
    for (int k = depth - 1; k >= 0; k--) {
      const float mid = (t0 + t1) * 0.5f;
      if ((path >> k) & 1ull) t0 = mid; else t1 = mid;
    }
  }
}

// Guided variant for the sphere paths.  Because only the boolean "some leaf survives" is consumed,
// the ORDER in which children are explored is free: at every node the child that contains t_pref
// (where the caller expects the bodies to be closest) is explored first, so a real contact is
// confirmed in O(depth) nodes with no backtracking.  End-point samples are carried down the tree:
// a child shares one end point with its parent, and eval(t) at the same float t gives the same
// bits, so each level costs one new sample.  path bit = 1 means "the non-preferred child".
//   eval(t)                       -> Sample   (the computed quantity at time t)
//   probe(t0, t1, Sample, Sample) -> kDead / kOpen / kSure
#ifdef PG_WALK_STATS
// debug build only (crux_b200/lib/alt): nodes visited per walk kind -- [kind][calls, nodes, max nodes, walks > 256 nodes]
static __device__ unsigned long long pg_walk_stats[4][4];
struct WalkCount {
  int kind; unsigned long long n; unsigned long long *out;
  __device__ WalkCount(int k, unsigned long long *o) : kind(k), n(0), out(o) {}
  __device__ ~WalkCount() {
    if (out) *out = n;
    atomicAdd(&pg_walk_stats[kind][0], 1ull); atomicAdd(&pg_walk_stats[kind][1], n); atomicMax(&pg_walk_stats[kind][2], n);
    if (n > 256) atomicAdd(&pg_walk_stats[kind][3], 1ull);
  }
};
#define PG_WALK_BEGIN(kind) WalkCount wc_(kind, dbg_nodes)
#define PG_WALK_NODE() (wc_.n++)
#else
#define PG_WALK_BEGIN(kind)
#define PG_WALK_NODE()
#endif

// `budget`: give up (return kWalkGaveUp) after that many nodes; the caller then decides the pair
// another way (ss_interval: leaf scan).
enum { kWalkFalse = 0, kWalkTrue = 1, kWalkGaveUp = 2 };
template <typename Sample, int Kind = 0, typename Eval, typename Probe>
__device__ __forceinline__ int interval_walk_guided(float t_begin, float t_end, float t_pref, Eval eval, Probe probe,
                                                    int budget = 0x7fffffff, unsigned long long *dbg_nodes = nullptr) {
  unsigned long long path = 0;
  int depth = 0;
  float t0 = t_begin, t1 = t_end;
  Sample Sa = eval(t0), Sb = eval(t1);
  PG_WALK_BEGIN(Kind);
  for (;;) {
    PG_WALK_NODE();
    if (--budget < 0) return kWalkGaveUp;
    const int st = probe(t0, t1, Sa, Sb);
    if (st == kSure) return kWalkTrue;
    if (st == kOpen) {
      if (t1 - t0 < kIntervalEps) return kWalkTrue;
      if (depth >= 62) return kWalkFalse;
      const float mid = (t0 + t1) * 0.5f;
      const Sample Sm = eval(mid);
      path <<= 1; depth++;
      if (t_pref >= mid) { t0 = mid; Sa = Sm; } else { t1 = mid; Sb = Sm; }
      continue;
    }
    while (depth > 0 && (path & 1ull)) { path >>= 1; depth--; }
    if (depth == 0) return kWalkFalse;
    path |= 1ull;
    t0 = t_begin; t1 = t_end;
    for (int k = depth - 1; k >= 0; k--) {
      const float mid = (t0 + t1) * 0.5f;
      const bool right = (t_pref >= mid) != (((path >> k) & 1ull) != 0ull);
      if (right) t0 = mid; else t1 = mid;
    }
    Sa = eval(t0); Sb = eval(t1);
  }
}

// Shortest and longest leaf interval (len < 1e-6) the walk can reach from a given root interval.
// The defaults hold for ANY root: a leaf's parent has len >= 1e-6, a midpoint splits it in halves
// up to a few ulps of t.  pg_leaf_lengths() (host) tabulates the exact extremes for a given root.
struct LeafLen { float lo, hi; };
__host__ __device__ __forceinline__ LeafLen default_leaf_len() { LeafLen l; l.lo = 0.4990e-6f; l.hi = 1.0e-6f; return l; }
constexpr float kU = 5.9604645e-8f;                  // 2^-24, unit roundoff of binary32
constexpr int kWalkBudget = 48;                      // nodes before a sphere-sphere walk switches to the leaf scan
constexpr int kScanMax = 8192;                       // widest leaf window the scan takes on

// Leaf scan behind ss_interval (see there); out of line and self-contained so that the common path
// keeps its registers.  Returns kWalkTrue / kWalkFalse, or kWalkGaveUp if the window is too wide or
// the halving tree is not uniform there (the caller then finishes the plain walk).
// Coop: 0 = the calling thread scans alone; 1 = all 32 lanes of the warp call with identical arguments
// and take every 32nd candidate; 2 = all threads of the BLOCK do (every blockDim.x-th candidate, one
// __syncthreads_or at the end).
template <int Coop>
__device__ __noinline__ int ss_leaf_scan(V3 cA, V3 vA, float nA, V3 cB, V3 vB, float nB, float rs,
                                         float t_begin, float t_end, LeafLen leaf) {
  const float tmax = fmaxf(fabsf(t_begin), fabsf(t_end));
  const float c1 = fabsf(cA.x) + fabsf(cA.y) + fabsf(cA.z) + fabsf(cB.x) + fabsf(cB.y) + fabsf(cB.z);
  const float v1 = fabsf(vA.x) + fabsf(vA.y) + fabsf(vA.z) + fabsf(vB.x) + fabsf(vB.y) + fabsf(vB.z);
  const float E0 = 1.25f * kU * (c1 + 2.0f * v1 * tmax) + 1.0e-30f;
  const float mm_hi = (nA * leaf.hi + nB * leaf.hi) * 1.000001f;
  if (!((nA + nB < 1.0e30f) && (c1 < 1.0e30f) && (rs < 1.0e30f) && (v1 < 1.0e30f))) return kWalkGaveUp;
  const double sx = (double)cA.x - (double)cB.x, sy = (double)cA.y - (double)cB.y, sz = (double)cA.z - (double)cB.z;
  const double ux = (double)vA.x - (double)vB.x, uy = (double)vA.y - (double)vB.y, uz = (double)vA.z - (double)vB.z;
  const double ss_d = sx * sx + sy * sy + sz * sz, su_d = sx * ux + sy * uy + sz * uz, uu_d = ux * ux + uy * uy + uz * uz;
  const double d2_slack = 1.0e-13 * (ss_d + 2.0 * fabs(su_d) * (double)tmax + uu_d * (double)tmax * (double)tmax);
  const double dead_thr = ((double)mm_hi + (double)rs + (double)E0) * 1.00001;
  const double dead_thr2 = dead_thr * dead_thr;
  auto eval = [&](float t) { return ss_d2_at(cA, vA, cB, vB, t); };
  {
    const double cq = ss_d - d2_slack - dead_thr2;            // uu t^2 + 2 su t + cq <= 0 inside the window
    double tL = (double)t_begin, tR = (double)t_end;
    if (uu_d > 0.0) {
      const double disc = su_d * su_d - uu_d * cq;
      if (disc < 0.0) return kWalkFalse;
      const double sq = sqrt(disc) * (1.0 + 1.0e-12);
      tL = (-su_d - sq) / uu_d; tR = (-su_d + sq) / uu_d;
    } else if (cq > 0.0) {
      return kWalkFalse;
    }
    int D = 0;                                                // depth of the leaves along the leftmost path
    for (float a = t_begin, b = t_end; b - a >= kIntervalEps && D < 40; D++) b = (a + b) * 0.5f;
    const double per_t = ldexp(1.0, D) / ((double)t_end - (double)t_begin);
    const double last = ldexp(1.0, D) - 1.0;
    const double k0d = fmax(floor((tL - (double)t_begin) * per_t) - 1.0, 0.0), k1d = fmin(floor((tR - (double)t_begin) * per_t) + 1.0, last);
    if (k0d > k1d) return kWalkFalse;                         // the window misses the interval
    if (k1d - k0d >= (double)kScanMax || D >= 40) return kWalkGaveUp;   // too wide to scan
    const long long k0 = (long long)k0d, k1 = (long long)k1d;
#ifdef PG_WALK_STATS
    atomicAdd(&pg_walk_stats[3][0], 1ull); atomicAdd(&pg_walk_stats[3][1], (unsigned long long)(k1 - k0 + 1)); atomicMax(&pg_walk_stats[3][2], (unsigned long long)(k1 - k0 + 1));
#endif
    bool found = false, unknown = false;
    long long k = k0;
    if (Coop == 1) k += threadIdx.x & 31;
    if (Coop == 2) k += threadIdx.x;
    const int k_step = Coop == 1 ? 32 : (Coop == 2 ? (int)blockDim.x : 1);
    for (; k <= k1 && !found; k += k_step) {
      float t0 = t_begin, t1 = t_end;
      float Sa = eval(t0), Sb = eval(t1);
      for (int level = D - 1;; level--) {
        const float len = t1 - t0;
        const float mm = nA * len + nB * len;
        if (ss_dist_exceeds(Sa, rs, mm) || ss_dist_exceeds(Sb, rs, mm)) break;        // this ancestor (or leaf) fails
        if (len < kIntervalEps) { found = true; break; }
        if (level < 0) { unknown = true; break; }             // deeper than the leftmost path: not a uniform tree
        const float mid = (t0 + t1) * 0.5f;
        const float Sm = eval(mid);
        if ((k >> level) & 1ll) { t0 = mid; Sa = Sm; } else { t1 = mid; Sb = Sm; }
      }
    }
    if (Coop == 1) { found = __any_sync(0xffffffffu, found); unknown = __any_sync(0xffffffffu, unknown); }
    if (Coop == 2) { found = __syncthreads_or(found) != 0; unknown = __syncthreads_or(unknown) != 0; }
    if (found) return kWalkTrue;
    return unknown ? kWalkGaveUp : kWalkFalse;
  }
}

// interval_collision for two spheres, world.c:557-582 with distance.c:278-299 inlined.
// cA/cB = centre + position at t=0, nA/nB = |v| (glm_vec3_norm), rs = rA + rB.
// Depth-first, left child first, like the recursion; an explicit stack keeps the right siblings.
// Enclosures (exactness-preserving).  The walk only needs to know whether SOME leaf survives; a
// leaf survives iff the COMPUTED distance is <= mm(len) = fl(fl(nA len) + fl(nB len)) at both end
// points, and mm is monotone in len.  For any float t in a node [a,b] the computed centre distance
// Dc(t) differs from the real D(t) = |s + u t| (real arithmetic on the float inputs) by at most
//   E = 1.25 u (|cA|_1 + |cB|_1 + 2 (|vA|_1 + |vB|_1) t_max) + 8 u D     (u = 2^-24)
// (roundings of v t, c + v t, the difference, the dot product and the square root).  D is convex,
// so on the node  D <= max(D(a), D(b)),  and D^2 is a parabola whose minimum over the node is known
// in closed form (vertex clamped into the node, evaluated in binary64).
//   lower bound of dist > mm(longest leaf)   -> no leaf below can survive          -> kDead
//   upper bound of dist <= mm(shortest leaf) -> every node below passes            -> kSure
// Fast bodies moving almost in parallel (|u| << nA+nB), and near misses of bodies moving head-on
// (|u| ~ nA+nB, where the reach shrinks as fast as the gap), would otherwise make the recursion
// visit hundreds to O(2^depth) nodes; with the enclosures it is O(depth) unless the closest
// approach lies inside a band of a few 1e-6 around mm(leaf).
// Coop = 1: all 32 lanes of the warp call with identical arguments (pg_spheres.cu) and share the leaf scan;
// Coop = 2: all threads of the block do (pg_large.cu, k_eval_hard).
// Defer (one thread per pair, pg_large.cu): a walk still going after kWalkBudget nodes returns
// kWalkGaveUp instead of being finished alone; the caller hands the pair to a warp (Coop).
template <int Coop, bool Defer>
__device__ inline int ss_interval_impl(V3 cA, V3 vA, float nA, V3 cB, V3 vB, float nB, float rs,
                                       float t_begin, float t_end, LeafLen leaf) {
  const V3 u = sub(vA, vB);
  const float tmax = fmaxf(fabsf(t_begin), fabsf(t_end));
  const float c1 = fabsf(cA.x) + fabsf(cA.y) + fabsf(cA.z) + fabsf(cB.x) + fabsf(cB.y) + fabsf(cB.z);
  const float v1 = fabsf(vA.x) + fabsf(vA.y) + fabsf(vA.z) + fabsf(vB.x) + fabsf(vB.y) + fabsf(vB.z);
  const float E0 = 1.25f * kU * (c1 + 2.0f * v1 * tmax) + 1.0e-30f;
  const float mm_hi = (nA * leaf.hi + nB * leaf.hi) * 1.000001f;   // >= mm of any leaf
  const float mm_lo = (nA * leaf.lo + nB * leaf.lo) * 0.999999f;   // <= mm of any node
  const bool finite_ok = (nA + nB < 1.0e30f) && (c1 < 1.0e30f) && (rs < 1.0e30f) && (v1 < 1.0e30f);
  // closest approach of the two linear paths: where a contact, if any, is
  const V3 s0 = sub(cA, cB);
  const float uu = __fmaf_rn(u.x, u.x, __fmaf_rn(u.y, u.y, u.z * u.z));
  const float su = __fmaf_rn(s0.x, u.x, __fmaf_rn(s0.y, u.y, s0.z * u.z));
  const float t_pref = uu > 0.0f ? fminf(fmaxf(-su / uu, t_begin), t_end) : t_begin;
#ifdef PG_WALK_STATS
  unsigned long long dbg_n = 0;
  struct Dump { const V3 &cA, &vA, &cB, &vB; float rs, t0, t1; unsigned long long &n;
    __device__ ~Dump() { if (n > 256) printf("SS %llu nodes cA %.9g %.9g %.9g vA %.9g %.9g %.9g cB %.9g %.9g %.9g vB %.9g %.9g %.9g rs %.9g t %.9g %.9g\n",
      n, cA.x, cA.y, cA.z, vA.x, vA.y, vA.z, cB.x, cB.y, cB.z, vB.x, vB.y, vB.z, rs, t0, t1); } } dump_{cA, vA, cB, vB, rs, t_begin, t_end, dbg_n};
#define PG_DBG_ARG , &dbg_n
#else
#define PG_DBG_ARG
#endif
  // The same parabola in binary64 (inputs are binary32, so s and u are exact and the three
  // coefficients carry a relative error of a few 2^-53): D^2(t) = ss + 2 su t + uu t^2.  A node is
  // dead as soon as the real minimum over it clears the longest leaf's reach:
  //   computed dist(t) >= (D(t)(1 - 8u) - E0 - rs)(1 - u) > mm_hi   <=   D_min > thr.
  const double sx = (double)cA.x - (double)cB.x, sy = (double)cA.y - (double)cB.y, sz = (double)cA.z - (double)cB.z;
  const double ux = (double)vA.x - (double)vB.x, uy = (double)vA.y - (double)vB.y, uz = (double)vA.z - (double)vB.z;
  const double ss_d = sx * sx + sy * sy + sz * sz, su_d = sx * ux + sy * uy + sz * uz, uu_d = ux * ux + uy * uy + uz * uz;
  const double nsu_d = -su_d;
  const double d2_slack = 1.0e-13 * (ss_d + 2.0 * fabs(su_d) * (double)tmax + uu_d * (double)tmax * (double)tmax);
  const double dead_thr = ((double)mm_hi + (double)rs + (double)E0) * 1.00001;
  const double dead_c2 = dead_thr * dead_thr + d2_slack;      // D^2 > dead_c2  =>  dead
  const double dead_c = ss_d - dead_c2;
  auto eval = [&](float t) { return ss_d2_at(cA, vA, cB, vB, t); };
  auto probe = [&](float t0, float t1, float d2a, float d2b) -> int {
    const float len = t1 - t0;
    const float mm = nA * len + nB * len;
    if (ss_dist_exceeds(d2a, rs, mm) || ss_dist_exceeds(d2b, rs, mm)) return kDead;
    if (len < kIntervalEps || !finite_ok) return kOpen;       // leaves are decided by the tests above
    // real minimum of |s + u t|^2 over the node, in closed form (no division: the vertex -su/uu lies
    // left of, right of, or inside the node; inside, the minimum is ss - su^2/uu)
    const double a0 = (double)t0 * uu_d, a1 = (double)t1 * uu_d;
    if (nsu_d > a0 && nsu_d < a1) {
      if (dead_c * uu_d > su_d * su_d) return kDead;
    } else {
      const double tc = nsu_d <= a0 ? (double)t0 : (double)t1;
      if (fma(tc, fma(uu_d, tc, 2.0 * su_d), ss_d) > dead_c2) return kDead;
    }
    const float Da = sqrtf(d2a), Db = sqrtf(d2b);
    const float Dmax = fmaxf(Da, Db);
    const float E = E0 + 8.0f * kU * Dmax;       // |computed - real| at any t of the node
    const float upper = fmaxf((Dmax + 2.0f * E) * 1.0000001f - rs, 0.0f) * 1.000001f;     // >= dist anywhere in the node
    if (upper <= mm_lo) return kSure;
    return kOpen;
  };
  // A grazing pair (closest approach within a few 1e-6 of the leaf reach -- typically the step after
  // a contact left the two spheres touching) keeps thousands of nodes around the vertex open.  After
  // kWalkBudget nodes the walk is abandoned for a LEAF SCAN: only leaves inside the window where
  // the real distance is below the dead threshold can pass their own test; each of them is reached
  // from the root by its own halving sequence, testing every ancestor on the way exactly as the
  // recursion would.  The candidates are independent, so a warp scans 32 at a time (Coop callers
  // only: a single thread scanning a window of hundreds of leaves is slower than its walk).
  for (int budget = (Coop || Defer) ? kWalkBudget : 0x7fffffff;; budget = 0x7fffffff) {     // one thread alone is better off walking
    const int r = interval_walk_guided<float>(t_begin, t_end, t_pref, eval, probe, budget PG_DBG_ARG);
    if (r != kWalkGaveUp || Defer) return r;
    const int q = ss_leaf_scan<Coop>(cA, vA, nA, cB, vB, nB, rs, t_begin, t_end, leaf);
    if (q != kWalkGaveUp) return q;
  }
}
template <int Coop = 0>
__device__ inline bool ss_interval(V3 cA, V3 vA, float nA, V3 cB, V3 vB, float nB, float rs,
                                   float t_begin, float t_end, LeafLen leaf = default_leaf_len()) {
  return ss_interval_impl<Coop, false>(cA, vA, nA, cB, vB, nB, rs, t_begin, t_end, leaf) == kWalkTrue;
}

// min_dist_at_time_sphere_plane, distance.c:305-321: max(|s| - r, 0), |s|-r in double.
__device__ __forceinline__ float sp_signed(V3 c, V3 n, float d) { return dot(c, n) - d; }
__device__ __forceinline__ float sp_dist_from_s(float s, float r) {
  // (float)(fabs((double)s) - (double)r): the double difference of two floats is exact unless
  // their exponents are > 28 apart, so one float subtraction gives the same bits.
  const float a = fabsf(s);
  float dist = (a < 1.0e8f * fabsf(r) && fabsf(r) < 1.0e8f * a) || a == 0.0f || r == 0.0f
                   ? a - r : (float)((double)a - (double)r);   // [f64]
  return glm_max(dist, 0);
}
__device__ __forceinline__ float sp_dist(V3 c, V3 n, float d, float r) { return sp_dist_from_s(sp_signed(c, n, d), r); }
// End points of the leaf at the very end (toward_end) or the very beginning of the halving tree of
// [t_begin, t_end]: the same (t0+t1)*0.5f sequence the recursion performs along that edge.
__device__ __forceinline__ void extreme_leaf(float t_begin, float t_end, bool toward_end, float &a, float &b) {
  a = t_begin; b = t_end;
  for (int k = 0; k < 64 && b - a >= kIntervalEps; k++) { const float m = (a + b) * 0.5f; if (toward_end) a = m; else b = m; }
}

// Same enclosure idea as ss_interval.  The real signed distance s(t) = n.(c + v t) - d is linear in
// t, the computed one differs from it by at most
//   Es = 1.5 u (sum_k |n_k| (4|c_k| + 4|v_k| t_max) + |s|_max)   (|n_k| <= 1; rigorous bound has factor 1.0)
// and is EXACTLY constant in t when every product v_k n_k vanishes (a body at rest that was
// pushed sideways along an axis-aligned plane: the case that otherwise costs 2^depth nodes every
// step, for ever, because a resting body is never integrated).
template <bool Hard>
__device__ __forceinline__ int sp_interval_impl(V3 c, V3 v, float nv, V3 n, float d, float r, float t_begin, float t_end, LeafLen leaf) {
  const float tmax = fmaxf(fabsf(t_begin), fabsf(t_end));
  const float c1 = fabsf(c.x) + fabsf(c.y) + fabsf(c.z), v1 = fabsf(v.x) + fabsf(v.y) + fabsf(v.z);
  const bool constant_s = (v.x == 0.0f || n.x == 0.0f) && (v.y == 0.0f || n.y == 0.0f) && (v.z == 0.0f || n.z == 0.0f);
  const bool unit_n = fabsf(n.x) <= 1.0000001f && fabsf(n.y) <= 1.0000001f && fabsf(n.z) <= 1.0000001f;
  // every rounding on the way to s carries a factor |n_k|: c_k = fl(p_k + fl(v_k t)) errs by
  // u(2|v_k|t + |c_k|), the product by u|c_k n_k| more, the two sums by u x partial sums, the final
  // subtraction by u|s|:  |s_computed - s_real| <= u (sum_k |n_k| (4|c_k| + 2|v_k| t) + |s|).  For an
  // axis-aligned wall only the coordinate along the normal counts.
  const float w1 = fabsf(n.x) * (4.0f * fabsf(c.x) + 4.0f * fabsf(v.x) * tmax) + fabsf(n.y) * (4.0f * fabsf(c.y) + 4.0f * fabsf(v.y) * tmax) +
                   fabsf(n.z) * (4.0f * fabsf(c.z) + 4.0f * fabsf(v.z) * tmax);
  const float s_mag = fabsf(c.x * n.x) + fabsf(c.y * n.y) + fabsf(c.z * n.z) + fabsf(d) + v1 * tmax;
  const float Es = constant_s ? 0.0f : 1.5f * kU * (w1 + s_mag) + 1.0e-30f;
  const float mm_hi = (nv * leaf.hi + 0.0f * leaf.hi) * 1.000001f;
  const float mm_lo = (nv * leaf.lo + 0.0f * leaf.lo) * 0.999999f;
  const bool finite_ok = unit_n && (nv < 1.0e30f) && (c1 < 1.0e30f) && (fabsf(d) < 1.0e30f) && (r < 1.0e30f) && (r >= 0.0f);
  // where the centre is closest to the plane (crossing time of the linear signed distance)
  const float s_begin = __fmaf_rn(c.x, n.x, __fmaf_rn(c.y, n.y, c.z * n.z)) - d;
  const float ndv = __fmaf_rn(v.x, n.x, __fmaf_rn(v.y, n.y, v.z * n.z));
  const float t_pref = ndv != 0.0f ? fminf(fmaxf(-s_begin / ndv, t_begin), t_end) : t_begin;
#ifdef PG_WALK_STATS
  unsigned long long dbg_sp = 0;
  struct DumpSp { const V3 &c, &v, &n; float d, r, t1, Es; unsigned long long &k;
    __device__ ~DumpSp() { if (k > 1000) printf("SP %llu nodes c %.9g %.9g %.9g v %.9g %.9g %.9g n %.9g %.9g %.9g d %.9g r %.9g dt %.9g Es %.9g\n",
      k, c.x, c.y, c.z, v.x, v.y, v.z, n.x, n.y, n.z, d, r, t1, Es); } } dump_sp{c, v, n, d, r, t_end, Es, dbg_sp};
#define PG_DBG_SP , 0x7fffffff, &dbg_sp
#else
#define PG_DBG_SP
#endif
  auto eval = [&](float t) { return sp_signed(muladds(v, t, c), n, d); };
  auto probe = [&](float t0, float t1, float sa, float sb) -> int {
      const float len = t1 - t0;
      const float mm = nv * len + 0.0f * len;      // the plane's |v| is 0 (world.c:51): 0*len, then the sum
      if (sp_dist_from_s(sa, r) > mm || sp_dist_from_s(sb, r) > mm) return kDead;
      if (len < kIntervalEps || !finite_ok) return kOpen;
      // computed s anywhere in the node lies in [lo, hi]
      const float lo = fminf(sa, sb) - 2.0f * Es, hi = fmaxf(sa, sb) + 2.0f * Es;
      const float amin = (lo <= 0.0f && hi >= 0.0f) ? 0.0f : fminf(fabsf(lo), fabsf(hi));
      const float amax = fmaxf(fabsf(lo), fabsf(hi));
      if ((amin - r) * 0.999999f > mm_hi) return kDead;
      if (fmaxf(amax - r, 0.0f) * 1.000001f <= mm_lo) return kSure;
      return kOpen;
    };
  if (!Hard) return interval_walk_guided<float, 1>(t_begin, t_end, t_pref, eval, probe, kWalkBudget);
  float t_pref_hard = t_pref;
  // Monotone signed distance (see bb_interval): the products fl(c_k(t) n_k) are monotone in t; if the
  // ones that move all move the same way, so does the computed s(t), exactly, and while it keeps its
  // sign so does the computed distance.  Its extremes over the interval are then the end values.
  if (finite_ok) {
    bool up = true, down = true;
    if (v.x != 0.0f && n.x != 0.0f) { const bool inc = (v.x > 0.0f) == (n.x > 0.0f); up = up && inc; down = down && !inc; }
    if (v.y != 0.0f && n.y != 0.0f) { const bool inc = (v.y > 0.0f) == (n.y > 0.0f); up = up && inc; down = down && !inc; }
    if (v.z != 0.0f && n.z != 0.0f) { const bool inc = (v.z > 0.0f) == (n.z > 0.0f); up = up && inc; down = down && !inc; }
    if (up || down) {
      const float sb = sp_signed(muladds(v, t_begin, c), n, d), se = sp_signed(muladds(v, t_end, c), n, d);
      if ((sb > 0.0f && se > 0.0f) || (sb < 0.0f && se < 0.0f)) {
        const float da = sp_dist_from_s(sb, r), db = sp_dist_from_s(se, r);
        if (fminf(da, db) > nv * leaf.hi + 0.0f * leaf.hi) return kWalkFalse;
        if (fmaxf(da, db) <= nv * leaf.lo + 0.0f * leaf.lo) return kWalkTrue;
        const bool near_end = db <= da;                    // see bb_interval
        float la, lb;
        extreme_leaf(t_begin, t_end, near_end, la, lb);
        const float far_d = sp_dist_from_s(eval(near_end ? la : lb), r);
        if (far_d > nv * leaf.hi + 0.0f * leaf.hi) return kWalkFalse;
        t_pref_hard = near_end ? t_end : t_begin;
      }
    }
  }
  return interval_walk_guided<float, 1>(t_begin, t_end, t_pref_hard, eval, probe PG_DBG_SP);
}
// A walk still going after kWalkBudget nodes is one of the rare band cases: out of line, first the
// exact monotone decision, then the plain walk to the end.
static __device__ __noinline__ bool sp_interval_hard(V3 c, V3 v, float nv, V3 n, float d, float r, float t_begin, float t_end, LeafLen leaf) {
  return sp_interval_impl<true>(c, v, nv, n, d, r, t_begin, t_end, leaf) == kWalkTrue;
}
__device__ inline bool sp_interval(V3 c, V3 v, float nv, V3 n, float d, float r, float t_begin, float t_end,
                                   LeafLen leaf = default_leaf_len()) {
  const int r0 = sp_interval_impl<false>(c, v, nv, n, d, r, t_begin, t_end, leaf);
  if (r0 != kWalkGaveUp) return r0 == kWalkTrue;
  return sp_interval_hard(c, v, nv, n, d, r, t_begin, t_end, leaf);
}

struct Contact { float t; bool hit; };
__device__ __forceinline__ Contact contact(float t, bool hit) { Contact c; c.t = t; c.hit = hit; return c; }

// narrow_phase_sphere_sphere, narrow_phase.c:459-520 (the early-outs fall through; see oracle).
__device__ __forceinline__ Contact ss_narrow(V3 cA, V3 vA, V3 cB, V3 vB, float rs) {
  V3 s = sub(cA, cB), v = sub(vA, vB);
  float c = dot(s, s) - (rs * rs);
  if (c < 0.0f) return contact(0.0f, true);
  float a = dot(v, v);
  float b = dot(v, s);
  float d = (b * b) - a * c;
  float t = (float)(((double)(-b) - sqrt((double)d)) / (double)a);   // [f64]
  return contact(t, true);
}
// narrow_phase.c:553-585 / 663-695 tail.
__device__ __forceinline__ Contact plane_tail(float s, float ndv, float r, float dt) {
  float disc = s * ndv;
  if (disc == 0) return fabs((double)s) <= (double)r ? contact(0.0f, true) : contact(-1.0f, false);
  float t = disc < 0 ? (r - s) / ndv : (-r - s) / ndv;
  if (t >= 0 && t <= dt) return contact(t, true);
  if (s <= r) return contact(0.0f, true);
  return contact(-1.0f, false);
}
__device__ __forceinline__ Contact sp_narrow(V3 c, V3 v, V3 n, float d, float r, float dt) {   // narrow_phase.c:526-588
  float s = dot(c, n) - d;
  float ndv = dot(v, n);
  return plane_tail(s, ndv, r, dt);
}

// "advance to hit time under gravity", resolution.c:474-479 and siblings.
__device__ __forceinline__ void advance_to_hit(V3 &pos, V3 vel, float t, V3 &vb) {
  vb = vel;
  vb.y -= kGravity * t;
  pos = muladds(vb, t, pos);
}
__device__ __forceinline__ V3 reflect(V3 vb, V3 n, float restitution) {   // resolution.c:606-609
  float vdn = dot(vb, n);
  V3 refl = scale(n, -2.0f * vdn * restitution);
  return add(vb, refl);
}
__device__ __forceinline__ void exchange_normal(V3 vbA, V3 vbB, V3 n, V3 &vA, V3 &vB) {   // resolution.c:515-528
  float a = dot(vbA, n), b = dot(vbB, n);
  V3 na = scale(n, a), nb = scale(n, b);
  vA = add(sub(vbA, na), nb);
  vB = add(sub(vbB, nb), na);
}
// resting rule, resolution.c:613-618: (double)v.n < 0.5 is the same set as v.n < 0.5f
__device__ __forceinline__ bool rest_rule(V3 n, V3 vel) {
  float vdn = dot(n, vel);
  return (vdn < 0.5f) && (dot(n, v3(0.0f, -1.0f, 0.0f)) < 0);
}

// resolve_collision_sphere_sphere, resolution.c:467-570.  oA/oB = local sphere centres.
__device__ inline void ss_resolve(V3 &pA, V3 &vA, V3 oA, V3 &pB, V3 &vB, V3 oB, float rs, float t) {
  V3 vbA, vbB;
  advance_to_hit(pA, vA, t, vbA);
  advance_to_hit(pB, vB, t, vbB);
  V3 diff = sub(add(oB, pB), add(oA, pA));
  float s = norm(diff);
  float pen = (s < rs) ? (float)((double)(rs - s) + 0.001) : 0.0f;   // [f64]
  V3 n = normalize(diff);
  if (pen <= 0.0f) return;
  V3 rv = sub(vbB, vbA);
  if (dot(rv, n) < 0.0f) {
    pB = muladds(n, pen / 2, pB);
    pA = mulsubs(n, pen / 2, pA);
    exchange_normal(vbA, vbB, n, vA, vB);
  } else { vA = vbA; vB = vbB; }
}
// resolve_collision_sphere_plane, resolution.c:576-622.  Returns true if the body came to rest.
__device__ inline bool sp_resolve(V3 &p, V3 &v, V3 o, float r, float restitution, V3 n, float d, float t) {
  V3 vb;
  advance_to_hit(p, v, t, vb);
  V3 c = add(o, p);
  float s = dot(c, n) - d;
  float pen = (s < r) ? (float)((double)(r - s) + 0.001) : 0.0f;   // [f64]
  if (pen > 0.0f) p = add(p, scale(n, pen));
  v = reflect(vb, n, restitution);
  if (rest_rule(n, v)) { v = v3(0.0f, 0.0f, 0.0f); return true; }
  return false;
}

// ---------------------------------------------------------------------------------------------
// Static AABB x dynamic sphere on scalars (large-world pass 3 against level geometry).  The box is
// the world-space AABB of a static body (host-computed once with AABB_update's arithmetic); the
// sphere is taken THROUGH ITS SCENE NODE, as the reference does (distance.c:113-125): its centre is
// the node translation p0 (the body position at the last scene_node_update, i.e. at step start),
// advanced by the CURRENT velocity, never by position changes made earlier in the step.
// ---------------------------------------------------------------------------------------------
struct BoxS { V3 c, e; };

// squared point-box distance, distance.c:146-154
__device__ __forceinline__ float bb_d2(V3 p, const BoxS &bx) {
  float d2 = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float v = get(p, i), lo = get(bx.c, i) - get(bx.e, i), hi = get(bx.c, i) + get(bx.e, i);
    if (v < lo) d2 += (lo - v) * (lo - v);
    if (v > hi) d2 += (v - hi) * (v - hi);
  }
  return d2;
}
// sphere centre as min_dist_at_time_AABB_sphere builds it: (p0 + v t), then local centre (0) + that
__device__ __forceinline__ V3 bb_centre(V3 p0, V3 v, float t) {
  const V3 q = muladds(v, t, p0);
  return v3(0.0f + q.x, 0.0f + q.y, 0.0f + q.z);
}
// interval_collision(box, sphere): mm = |v_box| len + |v_sphere| len with |v_box| = 0.  Same guided,
// enclosure-pruned walk as ss_interval: the distance of a moving point to a convex set is convex in
// t and |v|-Lipschitz.
template <bool Hard>
__device__ __forceinline__ int bb_interval_impl(V3 p0, V3 v, float nv, float r, const BoxS &bx, float t_begin, float t_end, LeafLen leaf) {
  const float tmax = fmaxf(fabsf(t_begin), fabsf(t_end));
  const float c1 = fabsf(p0.x) + fabsf(p0.y) + fabsf(p0.z);
  const float b1 = fabsf(bx.c.x) + fabsf(bx.c.y) + fabsf(bx.c.z) + fabsf(bx.e.x) + fabsf(bx.e.y) + fabsf(bx.e.z);
  const float v1 = fabsf(v.x) + fabsf(v.y) + fabsf(v.z);
  const float E0 = 1.25f * kU * (c1 + 2.0f * v1 * tmax + 3.0f * b1) + 1.0e-30f;
  const float L = nv * 1.0001f;
  const float mm_hi = (0.0f * leaf.hi + nv * leaf.hi) * 1.000001f;
  const float mm_lo = (0.0f * leaf.lo + nv * leaf.lo) * 0.999999f;
  const bool finite_ok = (nv < 1.0e30f) && (c1 < 1.0e30f) && (b1 < 1.0e30f) && (r < 1.0e30f);
  // closest approach to the box centre as the preferred time (a heuristic only)
  const V3 s0 = sub(p0, bx.c);
  const float vv = __fmaf_rn(v.x, v.x, __fmaf_rn(v.y, v.y, v.z * v.z));
  const float sv = __fmaf_rn(s0.x, v.x, __fmaf_rn(s0.y, v.y, s0.z * v.z));
  const float t_pref = vv > 0.0f ? fminf(fmaxf(-sv / vv, t_begin), t_end) : t_begin;
#ifdef PG_WALK_STATS
  unsigned long long dbg_bb = 0;
  struct DumpBb { const V3 &p, &v; const BoxS &b; float r, t1; unsigned long long &k;
    __device__ ~DumpBb() { if (k > 1000) printf("BB %llu nodes p %.9g %.9g %.9g v %.9g %.9g %.9g bc %.9g %.9g %.9g be %.9g %.9g %.9g r %.9g dt %.9g\n",
      k, p.x, p.y, p.z, v.x, v.y, v.z, b.c.x, b.c.y, b.c.z, b.e.x, b.e.y, b.e.z, r, t1); } } dump_bb{p0, v, bx, r, t_end, dbg_bb};
#define PG_DBG_BB , 0x7fffffff, &dbg_bb
#else
#define PG_DBG_BB
#endif
  auto eval = [&](float t) { return bb_d2(bb_centre(p0, v, t), bx); };
  auto probe = [&](float t0, float t1, float d2a, float d2b) -> int {
      const float len = t1 - t0;
      const float mm = 0.0f * len + nv * len;
      if (ss_dist_exceeds(d2a, r, mm) || ss_dist_exceeds(d2b, r, mm)) return kDead;   // same tail as sphere-sphere: d2 < r^2 ? 0 : sqrt(d2) - r
      if (len < kIntervalEps || !finite_ok) return kOpen;
      const float Da = sqrtf(d2a), Db = sqrtf(d2b);
      const float Dmax = fmaxf(Da, Db);
      const float E = E0 + 8.0f * kU * Dmax;
      const float lower = 0.5f * (Da + Db) * 0.9999999f - 0.5f * L * len - 2.0f * E - r;
      if (lower * 0.999999f > mm_hi) return kDead;
      const float upper = fmaxf((Dmax + 2.0f * E) * 1.0000001f - r, 0.0f) * 1.000001f;
      if (upper <= mm_lo) return kSure;
      return kOpen;
    };
  if (!Hard) return interval_walk_guided<float, 2>(t_begin, t_end, t_pref, eval, probe, kWalkBudget);
  float t_pref_hard = t_pref;
  // Monotone distance.  Each computed coordinate fl(p + fl(v t)) is monotone in t, and an axis'
  // contribution to the squared point-box distance -- (lo-p)^2 below the extent, 0 inside, (p-hi)^2
  // above -- is monotone in p on either side, roundings included.  Unless some axis crosses the
  // whole extent, every contribution therefore runs monotonically from its value at t_begin to its
  // value at t_end; if they all run the same way, so does the computed distance, EXACTLY, and its
  // extremes over every float t of the interval are the two end values.  Then
  //   smallest > mm(longest leaf)  -> no leaf can pass            -> false
  //   largest <= mm(shortest leaf) -> every node passes           -> true
  // with no rounding allowance at all.  This settles the sphere that rolls along a slab with a
  // vertical velocity of 1e-4: its height moves by less than an ulp per step, the distance sits a
  // few 1e-6 from the leaf reach, and the enclosures (ulp(72) = 7.6e-6) cannot tell.
  if (finite_ok) {
    const V3 pb = bb_centre(p0, v, t_begin), pe = bb_centre(p0, v, t_end);
    bool falling = true, rising = true;               // all contributions non-increasing / non-decreasing in t
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float lo = get(bx.c, i) - get(bx.e, i), hi = get(bx.c, i) + get(bx.e, i);
      const float a = get(pb, i), b = get(pe, i);
      const float fa = a < lo ? (lo - a) * (lo - a) : (a > hi ? (a - hi) * (a - hi) : 0.0f);
      const float fb = b < lo ? (lo - b) * (lo - b) : (b > hi ? (b - hi) * (b - hi) : 0.0f);
      const bool crosses = (a < lo && b > hi) || (a > hi && b < lo);
      falling = falling && !crosses && fa >= fb;
      rising = rising && !crosses && fa <= fb;
    }
    if (falling || rising) {
      const float da = ss_dist_from_d2(bb_d2(pb, bx), r), db = ss_dist_from_d2(bb_d2(pe, bx), r);
      if (fminf(da, db) > 0.0f * leaf.hi + nv * leaf.hi) return kWalkFalse;
      if (fmaxf(da, db) <= 0.0f * leaf.lo + nv * leaf.lo) return kWalkTrue;
      // Of all leaves the one at the near end has the smallest distances; a leaf passes iff the
      // larger of its two end values is within mm(its length) <= mm(longest leaf).
      const bool near_end = db <= da;
      float la, lb;
      extreme_leaf(t_begin, t_end, near_end, la, lb);
      const float far_d = ss_dist_from_d2(eval(near_end ? la : lb), r);
      if (far_d > 0.0f * leaf.hi + nv * leaf.hi) return kWalkFalse;
      t_pref_hard = near_end ? t_end : t_begin;            // and if a leaf passes, it is found there
    }
  }
  return interval_walk_guided<float, 2>(t_begin, t_end, t_pref_hard, eval, probe PG_DBG_BB);
}
static __device__ __noinline__ bool bb_interval_hard(V3 p0, V3 v, float nv, float r, const BoxS &bx, float t_begin, float t_end, LeafLen leaf) {
  return bb_interval_impl<true>(p0, v, nv, r, bx, t_begin, t_end, leaf) == kWalkTrue;
}
__device__ inline bool bb_interval(V3 p0, V3 v, float nv, float r, const BoxS &bx, float t_begin, float t_end,
                                   LeafLen leaf = default_leaf_len()) {
  const int r0 = bb_interval_impl<false>(p0, v, nv, r, bx, t_begin, t_end, leaf);
  if (r0 != kWalkGaveUp) return r0 == kWalkTrue;
  return bb_interval_hard(p0, v, nv, r, bx, t_begin, t_end, leaf);
}
// narrow_phase_AABB_sphere, narrow_phase.c:143-196 + ray_intersect_AABB :702-739 (static box: rel_v = v - 0)
__device__ inline Contact bb_narrow(V3 p0, V3 v, float r, const BoxS &bx, float dt) {
  const V3 c = v3(0.0f + p0.x, 0.0f + p0.y, 0.0f + p0.z);
  const V3 e = v3(bx.e.x + r, bx.e.y + r, bx.e.z + r);
  const V3 rv = sub(v, v3(0.0f, 0.0f, 0.0f));
  float hit = 0.0f, t_max = dt;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float lo = get(bx.c, i) - get(e, i), hi = get(bx.c, i) + get(e, i), pi = get(c, i), di = get(rv, i);
    if ((double)fabsf(di) < kNpEps) {
      if (pi < lo || pi > hi) return contact(-1.0f, false);
    } else {
      float t1 = (lo - pi) / di;
      float t2 = (hi - pi) / di;
      if (t1 > t2) { const float tmp = t1; t1 = t2; t2 = tmp; }
      if (t1 > hit) hit = t1;
      if (t2 < t_max) t_max = t2;
      if (hit > t_max) return contact(-1.0f, false);
    }
  }
  if (hit > dt) return contact(-1.0f, false);
  return contact(hit, true);
}
// resolve_collision_AABB_sphere with a static box and a dynamic sphere, resolution.c:152-241.
// pos/vel = the body's CURRENT position and velocity (updated in place); p0 = node translation.
__device__ inline void bb_resolve(V3 &pos, V3 &vel, V3 p0, float r, float restitution, const BoxS &bx, float t) {
  V3 vb;
  advance_to_hit(pos, vel, t, vb);
  const V3 q = muladds(vb, t, p0);                         // node position advanced by velocity_before (resolution.c:190)
  const V3 c = v3(0.0f + q.x, 0.0f + q.y, 0.0f + q.z);
  V3 closest;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float lo = get(bx.c, i) - get(bx.e, i), hi = get(bx.c, i) + get(bx.e, i), ci = get(c, i);
    set(closest, i, lo > ci ? lo : (hi > ci ? ci : hi));
  }
  const V3 diff = sub(c, closest);
  const float d = norm(diff);
  const float pen = (d < r) ? (float)((double)(r - d) + 0.001) : 0.0f;   // [f64]
  if (pen <= 0) return;
  const V3 n = normalize(diff);
  pos = muladds(n, pen * 0.5f, pos);
  vel = reflect(vb, n, restitution);
}

// =====================================================================================
// General path: all collider pairs, on PgBody records.
// =====================================================================================

__device__ __forceinline__ float max_move(const PgBody &b, float t0, float t1) {   // world.c:584-589
  float dt = t1 - t0;
  return norm(v3p(b.velocity)) * dt;
}

__device__ inline float dist_box_box(const PgBody &A, const PgBody &B, float t) {   // distance.c:20-92
  Box a = box_from_node(A, true, v3p(A.velocity), t), b = box_from_node(B, true, v3p(B.velocity), t);
  float d2 = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float minA = get(a.c, i) - get(a.e, i), maxA = get(a.c, i) + get(a.e, i);
    float minB = get(b.c, i) - get(b.e, i), maxB = get(b.c, i) + get(b.e, i);
    if (minB > maxA) d2 += (minB - maxA) * (minB - maxA);
    else if (maxB < minA) d2 += (minA - maxB) * (minA - maxB);
  }
  return sqrt_f64_as_f32(d2);
}
__device__ inline float dist_box_ball(const PgBody &A, const PgBody &B, float t) {   // distance.c:94-158
  Box bx = box_from_node(A, true, v3p(A.velocity), t);
  Ball s = ball_from_node(B, true, v3p(B.velocity), t);
  float d2 = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float v = get(s.c, i), lo = get(bx.c, i) - get(bx.e, i), hi = get(bx.c, i) + get(bx.e, i);
    if (v < lo) d2 += (lo - v) * (lo - v);
    if (v > hi) d2 += (v - hi) * (v - hi);
  }
  return d2 < (s.r * s.r) ? 0.0f : (float)(sqrt((double)d2) - (double)s.r);   // [f64]
}
__device__ inline float dist_box_pill(const PgBody &A, const PgBody &B, float t) {   // distance.c:160-244
  Box bx = box_from_node(A, true, v3p(A.velocity), t);
  Pill p = pill_from_node(B);
  V3 pq = pill_box_pq(p, bx);
  return glm_max(norm(pq) - p.r, 0);
}
__device__ inline float dist_box_flat(const PgBody &A, const PgBody &B, float t) {   // distance.c:246-276
  V3 pos = muladds(v3p(A.velocity), t, v3p(A.position));
  Box bx = box_from_body(A, pos);
  V3 n = v3p(B.shape);
  float r = proj_radius_f64(bx.e, n);
  float s = dot(n, bx.c) - B.shape[3];
  return glm_max(s - r, 0);
}
__device__ inline float dist_ball_ball(const PgBody &A, const PgBody &B, float t) {   // distance.c:278-299
  Ball a = ball_from_body(A, v3p(A.position)), b = ball_from_body(B, v3p(B.position));
  float d2 = ss_d2_at(a.c, v3p(A.velocity), b.c, v3p(B.velocity), t);
  return ss_dist_from_d2(d2, a.r + b.r);
}
__device__ inline float dist_ball_flat(const PgBody &A, const PgBody &B, float t) {   // distance.c:305-321
  Ball a = ball_from_body(A, v3p(A.position));
  return sp_dist(muladds(v3p(A.velocity), t, a.c), v3p(B.shape), B.shape[3], a.r);
}
__device__ inline float dist_pill_flat(const PgBody &A, const PgBody &B, float) {   // distance.c:327-410
  Flat f = flat_for_pill(B);
  Pill p = pill_from_body(A);
  V3 cp = pill_closest_to_flat(p, f);
  float s = dot(cp, f.n) - f.d;
  float d = (float)(fabs((double)s) - (double)p.r);   // [f64]
  return glm_max(d, 0);
}

// distance_functions table, distance.c:7-17.  Missing / undefined entries report 0 ("always close").
__device__ inline float min_dist(const PgBody &A, const PgBody &B, float t) {
  switch (A.collider_type * 4 + B.collider_type) {
    case PG_AABB * 4 + PG_AABB: return dist_box_box(A, B, t);
    case PG_AABB * 4 + PG_SPHERE: return dist_box_ball(A, B, t);
    case PG_AABB * 4 + PG_CAPSULE: return dist_box_pill(A, B, t);
    case PG_AABB * 4 + PG_PLANE: return dist_box_flat(A, B, t);
    case PG_SPHERE * 4 + PG_SPHERE: return dist_ball_ball(A, B, t);
    case PG_SPHERE * 4 + PG_PLANE: return dist_ball_flat(A, B, t);
    case PG_CAPSULE * 4 + PG_PLANE: return dist_pill_flat(A, B, t);
    default: return 0.0f;
  }
}

// ---- the time-independent part of a pair's distance function, taken out of the interval search ----
// interval_collision evaluates minimum_object_distance_at_time at two end points per node of its recursion, and every
// evaluation of a box rebuilds it from the scene node: three column norms, the rotation scaled by 1/scale.x, nine
// products and nine double-precision |rot| * extent * |scale| terms (distance.c:26-49 + aabb.c:59-78) -- none of which
// depends on the time; only the translation does (pos + velocity * t).  A game-sized visit is one thread walking some
// twenty nodes, so that rebuilding WAS the frame time of the drop-in (scenes/bouncehouse.json: 118 of 148 us).  PairPre
// holds exactly the values those expressions produce -- each rot * shape product, the finished extents, the capsule,
// the plane's projected radius -- and box_at() redoes only the operations that involve the time, in the reference's
// order (centre = ((tr + p0) + p1) + p2 per axis), so every distance is bit-identical to min_dist()'s.
struct BoxPre { V3 pos0, e; float p[9]; int ok; };
__device__ inline BoxPre box_pre(const float *shape, const float *rot, V3 pos0, V3 scl) {
  BoxPre r; r.ok = 1; r.pos0 = pos0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float e = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      r.p[j * 3 + i] = rot[j * 3 + i] * shape[j];
      e = (float)((double)e + fabs((double)rot[j * 3 + i]) * (double)shape[3 + j] * fabs((double)get(scl, i)));
    }
    set(r.e, i, e);
  }
  return r;
}
__device__ inline BoxPre box_pre_node(const PgBody &b) {          // box_from_node without the advance
  if (!b.has_node) { BoxPre r; r.ok = 0; r.pos0 = v3(0, 0, 0); r.e = v3(0, 0, 0); for (int k = 0; k < 9; k++) r.p[k] = 0.0f; return r; }
  NodeFrame f; node_decompose(b.node_xform, f);
  return box_pre(b.shape, f.rot, f.pos, f.scl);
}
__device__ __forceinline__ Box box_at(const BoxPre &r, V3 vel, float t) {     // == aabb_update(shape, rot, pos0 + vel t, scl)
  if (!r.ok) return zero_box();
  const V3 tr = muladds(vel, t, r.pos0);
  Box o;
  o.c.x = ((tr.x + r.p[0]) + r.p[3]) + r.p[6];
  o.c.y = ((tr.y + r.p[1]) + r.p[4]) + r.p[7];
  o.c.z = ((tr.z + r.p[2]) + r.p[5]) + r.p[8];
  o.e = r.e;
  return o;
}
struct PairPre {
  int code;               // A.collider_type * 4 + B.collider_type
  float nA, nB;           // |velocity|: maximum_object_movement_over_time = |v| (t1 - t0), world.c:584-589
  V3 vA, vB;
  BoxPre a, b;            // the boxes of the pair (b.pos0 / b.ok: the sphere's node for a box-sphere pair)
  Pill pill;              // box-capsule: the capsule does not move during the search (distance.c:185-189)
  float aux;              // box-sphere: world radius; box-plane: projected radius; capsule-plane: the distance itself
};
// na / nb: the caller already holds box_pre_node(A) / box_pre_node(B) (a kernel that builds them once per body and step)
__device__ inline PairPre pair_pre(const PgBody &A, const PgBody &B, const BoxPre *na = nullptr, const BoxPre *nb = nullptr) {
  PairPre q;
  q.code = A.collider_type * 4 + B.collider_type;
  q.vA = v3p(A.velocity); q.vB = v3p(B.velocity);
  q.nA = norm(q.vA); q.nB = norm(q.vB);
  q.aux = 0.0f;
  switch (q.code) {
    case PG_AABB * 4 + PG_AABB: q.a = na ? *na : box_pre_node(A); q.b = nb ? *nb : box_pre_node(B); break;
    case PG_AABB * 4 + PG_SPHERE:
      q.a = na ? *na : box_pre_node(A);
      q.b.ok = B.has_node; q.b.pos0 = v3(0, 0, 0);
      if (B.has_node) { NodeFrame f; node_decompose(B.node_xform, f); q.b.pos0 = f.pos; q.aux = B.shape[3] * f.scl.x; }
      break;
    case PG_AABB * 4 + PG_CAPSULE: q.a = na ? *na : box_pre_node(A); q.pill = pill_from_node(B); break;
    case PG_AABB * 4 + PG_PLANE:
      q.a = box_pre(A.shape, A.euler_raw, v3p(A.position), v3p(A.scale));
      q.aux = proj_radius_f64(q.a.e, v3p(B.shape));
      break;
    case PG_CAPSULE * 4 + PG_PLANE: q.aux = min_dist(A, B, 0.0f); break;      // (no term of it depends on the time)
    default: break;
  }
  return q;
}
// minimum_object_distance_at_time through the table (min_dist below), with the time-independent part precomputed
__device__ inline float min_dist_pre(const PgBody &A, const PgBody &B, const PairPre &q, float t) {
  switch (q.code) {
    case PG_AABB * 4 + PG_AABB: {
      const Box a = box_at(q.a, q.vA, t), b = box_at(q.b, q.vB, t);
      float d2 = 0.0f;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float minA = get(a.c, i) - get(a.e, i), maxA = get(a.c, i) + get(a.e, i);
        float minB = get(b.c, i) - get(b.e, i), maxB = get(b.c, i) + get(b.e, i);
        if (minB > maxA) d2 += (minB - maxA) * (minB - maxA);
        else if (maxB < minA) d2 += (minA - maxB) * (minA - maxB);
      }
      return sqrt_f64_as_f32(d2);
    }
    case PG_AABB * 4 + PG_SPHERE: {
      const Box bx = box_at(q.a, q.vA, t);
      Ball s; s.c = v3(0, 0, 0); s.r = 0.0f;
      if (q.b.ok) { s.c = add(v3p(B.shape), muladds(q.vB, t, q.b.pos0)); s.r = q.aux; }
      float d2 = 0.0f;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float v = get(s.c, i), lo = get(bx.c, i) - get(bx.e, i), hi = get(bx.c, i) + get(bx.e, i);
        if (v < lo) d2 += (lo - v) * (lo - v);
        if (v > hi) d2 += (v - hi) * (v - hi);
      }
      return d2 < (s.r * s.r) ? 0.0f : (float)(sqrt((double)d2) - (double)s.r);   // [f64]
    }
    case PG_AABB * 4 + PG_CAPSULE: {
      const Box bx = box_at(q.a, q.vA, t);
      const V3 pq = pill_box_pq(q.pill, bx);
      return glm_max(norm(pq) - q.pill.r, 0);
    }
    case PG_AABB * 4 + PG_PLANE: {
      const Box bx = box_at(q.a, q.vA, t);
      const float s = dot(v3p(B.shape), bx.c) - B.shape[3];
      return glm_max(s - q.aux, 0);
    }
    case PG_CAPSULE * 4 + PG_PLANE: return q.aux;
    default: return min_dist(A, B, t);
  }
}
// one node of interval_collision's recursion: do both end points pass?  (world.c:559-565)
__device__ __forceinline__ bool interval_node_open(const PgBody &A, const PgBody &B, const PairPre &q, float t0, float t1) {
  const float dt = t1 - t0;
  const float mm = q.nA * dt + q.nB * dt;
  return !(min_dist_pre(A, B, q, t0) > mm) && !(min_dist_pre(A, B, q, t1) > mm);
}

// interval_collision, world.c:557-582 (general collider pair).
// The walk of interval_walk() (children left first, true at the first surviving leaf, *hit = its t0) with the distance at
// t0 carried down left descents: the left child shares its parent's t0, and the same float time gives the same bits.
__device__ inline bool interval_hit(const PgBody &A, const PgBody &B, float t_begin, float t_end, float *hit) {
  const PairPre q = pair_pre(A, B);
  unsigned long long path = 0;
  int depth = 0;
  float t0 = t_begin, t1 = t_end, d0 = 0.0f;
  bool have_d0 = false;
  for (;;) {
    const float len = t1 - t0;
    const float mm = q.nA * len + q.nB * len;
    if (!have_d0) d0 = min_dist_pre(A, B, q, t0);
    if (!(d0 > mm) && !(min_dist_pre(A, B, q, t1) > mm)) {
      if (len < kIntervalEps) { if (hit) *hit = t0; return true; }
      if (depth >= 62) return false;                 // unreachable for finite end points
      path <<= 1; depth++;
      t1 = (t0 + t1) * 0.5f;
      have_d0 = true;
      continue;
    }
    have_d0 = false;
    while (depth > 0 && (path & 1ull)) { path >>= 1; depth--; }   // right children are finished
    if (depth == 0) return false;
    path |= 1ull;                                    // left child failed: take its right sibling
    t0 = t_begin; t1 = t_end;
    for (int k = depth - 1; k >= 0; k--) {
      const float mid = (t0 + t1) * 0.5f;
      if ((path >> k) & 1ull) t0 = mid; else t1 = mid;
    }
  }
}

__device__ inline bool aabb_hits_aabb(const Box &a, const Box &b) {   // aabb.c:100-105
  if (fabs((double)(a.c.x - b.c.x)) > (double)(a.e.x + b.e.x)) return false;
  if (fabs((double)(a.c.y - b.c.y)) > (double)(a.e.y + b.e.y)) return false;
  if (fabs((double)(a.c.z - b.c.z)) > (double)(a.e.z + b.e.z)) return false;
  return true;
}

__device__ inline Contact np_box_box(const PgBody &A, const PgBody &B, float dt) {   // narrow_phase.c:21-141
  Box a = box_from_node(A, false, v3(0, 0, 0), 0), b = box_from_node(B, false, v3(0, 0, 0), 0);
  if (aabb_hits_aabb(a, b)) return contact(0.0f, true);
  V3 rv = sub(v3p(A.velocity), v3p(B.velocity));
  float t_first = 0.0f, t_last = dt;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float minA = get(a.c, i) - get(a.e, i), maxA = get(a.c, i) + get(a.e, i);
    float minB = get(b.c, i) - get(b.e, i), maxB = get(b.c, i) + get(b.e, i);
    float r = get(rv, i);
    if (r < 0.0f) {
      if (maxA < minB) return contact(-1.0f, false);
      if (minA > maxB) t_first = fmaxf((maxB - minA) / r, t_first);
      if (maxA > minB) t_last = fminf((minB - maxA) / r, t_last);
    }
    if (r > 0.0f) {
      if (minA > maxB) return contact(-1.0f, false);
      if (maxA < minB) t_first = fmaxf((minB - maxA) / r, t_first);
      if (minA < maxB) t_last = fminf((maxB - minA) / r, t_last);
    }
    if (t_first > t_last) return contact(-1.0f, false);
  }
  return contact(t_first, true);
}
__device__ inline bool ray_box(V3 p, V3 d, V3 c, V3 e, float *hit, float t_end) {   // narrow_phase.c:702-739
  *hit = 0.0f;
  float t_max = t_end;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float lo = get(c, i) - get(e, i), hi = get(c, i) + get(e, i), pi = get(p, i), di = get(d, i);
    if ((double)fabsf(di) < kNpEps) {
      if (pi < lo || pi > hi) return false;
    } else {
      float t1 = (lo - pi) / di;
      float t2 = (hi - pi) / di;
      if (t1 > t2) { float tmp = t1; t1 = t2; t2 = tmp; }
      if (t1 > *hit) *hit = t1;
      if (t2 < t_max) t_max = t2;
      if (*hit > t_max) return false;
    }
  }
  return true;
}
__device__ inline Contact np_box_ball(const PgBody &A, const PgBody &B, float dt) {   // narrow_phase.c:143-196
  Box bx = box_from_node(A, false, v3(0, 0, 0), 0);
  Ball s = ball_from_node(B, false, v3(0, 0, 0), 0);
  V3 e = v3(bx.e.x + s.r, bx.e.y + s.r, bx.e.z + s.r);
  V3 rv = sub(v3p(B.velocity), v3p(A.velocity));
  float tmin;
  bool hit = ray_box(s.c, rv, bx.c, e, &tmin, dt);
  if (!hit || tmin > dt) return contact(-1.0f, false);
  return contact(tmin, true);
}
__device__ inline Contact np_box_pill(const PgBody &A, const PgBody &B, float dt) {   // narrow_phase.c:233-342
  Box bx = box_from_node(A, false, v3(0, 0, 0), 0);
  Pill p = pill_from_node(B);
  V3 pq = pill_box_pq(p, bx);
  float distance = norm(pq) - p.r;
  pq = normalize(pq);
  float pq_dot_v = dot(v3p(B.velocity), pq);
  float disc = distance * pq_dot_v;
  if (disc == 0) return distance < 0 ? contact(0.0f, true) : contact(-1.0f, false);
  float t = distance / pq_dot_v;
  if (t >= 0 && t <= dt) return contact(t, true);
  if (distance <= 0) return contact(0.0f, true);
  return contact(-1.0f, false);
}
__device__ inline Contact np_box_flat(const PgBody &A, const PgBody &B, float dt) {   // narrow_phase.c:344-457
  Box bx = box_from_body(A, v3p(A.position));
  V3 n = v3p(B.shape);
  float r = proj_radius_f64(bx.e, n);
  float s = dot(n, bx.c) - B.shape[3];
  float ndv = dot(n, v3p(A.velocity));
  bool inside = fabs((double)s) <= (double)r;
  if (ndv == 0) return inside ? contact(0.0f, true) : contact(-1.0f, false);
  Contact c;
  if (ndv < 0) { c.t = (r - s) / -ndv; c.hit = (c.t >= 0 && c.t <= dt); }
  else if (inside) { c.t = 0.0f; c.hit = true; }
  else { c.t = (r + s) / -ndv; c.hit = (c.t >= 0 && c.t <= dt); }
  if (c.hit) c.t = fmaxf(0.0f, fminf(c.t, dt));
  if (!c.hit) { if (inside) return contact(0.0f, true); c.t = -1.0f; }
  return c;
}
__device__ inline Contact np_ball_ball(const PgBody &A, const PgBody &B, float) {
  Ball a = ball_from_body(A, v3p(A.position)), b = ball_from_body(B, v3p(B.position));
  return ss_narrow(a.c, v3p(A.velocity), b.c, v3p(B.velocity), a.r + b.r);
}
__device__ inline Contact np_ball_flat(const PgBody &A, const PgBody &B, float dt) {
  Ball a = ball_from_body(A, v3p(A.position));
  return sp_narrow(a.c, v3p(A.velocity), v3p(B.shape), B.shape[3], a.r, dt);
}
__device__ inline Contact np_pill_flat(const PgBody &A, const PgBody &B, float dt) {   // narrow_phase.c:594-698
  Flat f = flat_for_pill(B);
  Pill p = pill_from_body(A);
  V3 cp = pill_closest_to_flat(p, f);
  float s = dot(cp, f.n) - f.d;
  float ndv = dot(v3p(A.velocity), f.n);
  return plane_tail(s, ndv, p.r, dt);
}
// narrow_phase_functions table, narrow_phase.c:8-19.  *has=false for NULL / empty-body entries.
__device__ inline Contact narrow(const PgBody &A, const PgBody &B, float dt, bool *has) {
  *has = true;
  switch (A.collider_type * 4 + B.collider_type) {
    case PG_AABB * 4 + PG_AABB: return np_box_box(A, B, dt);
    case PG_AABB * 4 + PG_SPHERE: return np_box_ball(A, B, dt);
    case PG_AABB * 4 + PG_CAPSULE: return np_box_pill(A, B, dt);
    case PG_AABB * 4 + PG_PLANE: return np_box_flat(A, B, dt);
    case PG_SPHERE * 4 + PG_SPHERE: return np_ball_ball(A, B, dt);
    case PG_SPHERE * 4 + PG_PLANE: return np_ball_flat(A, B, dt);
    case PG_CAPSULE * 4 + PG_PLANE: return np_pill_flat(A, B, dt);
    default: *has = false; return contact(0.0f, false);
  }
}

// ---- resolution (resolution.c) on records ----
__device__ __forceinline__ void body_advance(PgBody &b, float t, V3 &vb) {
  V3 p = v3p(b.position);
  advance_to_hit(p, v3p(b.velocity), t, vb);
  store(b.position, p);
}
__device__ __forceinline__ void body_rest_rule(PgBody &b, V3 n) {
  if (rest_rule(n, v3p(b.velocity))) { b.velocity[0] = b.velocity[1] = b.velocity[2] = 0.0f; b.at_rest = 1; }
}
__device__ inline void dyn_dyn_tail(PgBody &A, PgBody &B, V3 vbA, V3 vbB, V3 dir, V3 n, float half_pen) {
  V3 rv = sub(vbB, vbA);
  if (dot(rv, n) < 0.0f) {
    store(B.position, muladds(dir, half_pen, v3p(B.position)));
    store(A.position, mulsubs(dir, half_pen, v3p(A.position)));
    V3 vA, vB; exchange_normal(vbA, vbB, n, vA, vB);
    store(A.velocity, vA); store(B.velocity, vB);
  } else { store(A.velocity, vbA); store(B.velocity, vbB); }
}
__device__ inline void rs_box_box(PgBody &A, PgBody &B, float t) {   // resolution.c:21-150
  V3 vbA = v3(0, 0, 0), vbB = v3(0, 0, 0);      // static side: 0 (uninitialised upstream)
  if (A.dynamic) body_advance(A, t, vbA);
  if (B.dynamic) body_advance(B, t, vbB);
  Box a = box_from_node(A, true, vbA, t), b = box_from_node(B, true, vbB, t);
  V3 cd = sub(b.c, a.c);
  float min_pen = FLT_MAX; int axis = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float s = (float)fabs((double)get(cd, i));
    float r = get(a.e, i) + get(b.e, i);
    float pen = r - s;
    if (pen < min_pen) { min_pen = pen; axis = i; }
  }
  if (min_pen <= 0.0f) return;
  V3 sep = v3(0, 0, 0);
  set(sep, axis, get(cd, axis) < 0 ? -1.0f : 1.0f);
  V3 n = normalize(sep);
  if (A.dynamic && !B.dynamic) {
    store(A.position, mulsubs(sep, min_pen, v3p(A.position)));
    store(A.velocity, reflect(vbA, n, A.restitution));
  } else if (!A.dynamic && B.dynamic) {
    store(B.position, muladds(sep, min_pen, v3p(B.position)));
    float vdn = dot(vbB, n);
    V3 refl = scale(n, -2.0f * vdn * B.restitution);
    store(B.velocity, add(vbA, refl));          // sic: velocity_before_A, resolution.c:118
  } else {
    dyn_dyn_tail(A, B, vbA, vbB, sep, n, min_pen * 0.5f);
  }
}
__device__ inline void rs_box_ball(PgBody &A, PgBody &B, float t) {   // resolution.c:152-302
  V3 vbA = v3(0, 0, 0), vbB = v3(0, 0, 0);
  if (A.dynamic) body_advance(A, t, vbA);
  if (B.dynamic) body_advance(B, t, vbB);
  Box bx = box_from_node(A, true, vbA, t);
  Ball s = ball_from_node(B, true, vbB, t);
  V3 closest;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float lo = get(bx.c, i) - get(bx.e, i), hi = get(bx.c, i) + get(bx.e, i), ci = get(s.c, i);
    set(closest, i, lo > ci ? lo : (hi > ci ? ci : hi));
  }
  V3 diff = sub(s.c, closest);
  float d = norm(diff);
  float pen = (d < s.r) ? (float)((double)(s.r - d) + 0.001) : 0.0f;   // [f64]
  if (pen <= 0) return;
  V3 n = normalize(diff);
  if (A.dynamic && !B.dynamic) {
    store(A.position, mulsubs(n, pen * 0.5f, v3p(A.position)));
    store(A.velocity, reflect(vbA, n, A.restitution));
  } else if (!A.dynamic && B.dynamic) {
    store(B.position, muladds(n, pen * 0.5f, v3p(B.position)));
    store(B.velocity, reflect(vbB, n, B.restitution));
  } else {
    dyn_dyn_tail(A, B, vbA, vbB, n, n, pen * 0.5f);
  }
}
__device__ inline void rs_box_pill(PgBody &A, PgBody &B, float t) {   // resolution.c:304-414
  V3 vb; body_advance(B, t, vb);
  Box bx = box_from_node(A, false, v3(0, 0, 0), 0);
  Pill p = pill_from_node(B);
  if (B.has_node) { p.a = muladds(vb, t, p.a); p.b = muladds(vb, t, p.b); }
  V3 pq = pill_box_pq(p, bx);
  float distance = norm(pq);
  pq = normalize(pq);
  float pen = distance < p.r ? (p.r - distance) + 0.001f : 0.0f;
  if (pen <= 0.0f) return;
  store(B.position, sub(v3p(B.position), scale(pq, pen)));
  store(B.velocity, reflect(vb, pq, B.restitution));
  body_rest_rule(B, v3(-pq.x, -pq.y, -pq.z));
}
__device__ inline void rs_box_flat(PgBody &A, PgBody &B, float t) {   // resolution.c:416-466
  V3 vb; body_advance(A, t, vb);
  Box bx = box_from_body(A, v3p(A.position));
  V3 n = v3p(B.shape);
  float r = bx.e.x * fabsf(n.x) + bx.e.y * fabsf(n.y) + bx.e.z * fabsf(n.z);
  float s = dot(n, bx.c) - B.shape[3];
  float pen = (s < r) ? (float)((double)(r - s) + 0.001) : 0.0f;   // [f64]
  if (pen > 0.0f) store(A.position, add(v3p(A.position), scale(n, pen)));
  store(A.velocity, reflect(vb, n, A.restitution));
  body_rest_rule(A, n);
}
__device__ inline void rs_ball_ball(PgBody &A, PgBody &B, float t) {
  V3 pA = v3p(A.position), vA = v3p(A.velocity), pB = v3p(B.position), vB = v3p(B.velocity);
  float rs = A.shape[3] * A.scale[0] + B.shape[3] * B.scale[0];
  ss_resolve(pA, vA, v3p(A.shape), pB, vB, v3p(B.shape), rs, t);
  store(A.position, pA); store(A.velocity, vA); store(B.position, pB); store(B.velocity, vB);
}
__device__ inline void rs_ball_flat(PgBody &A, PgBody &B, float t) {
  V3 p = v3p(A.position), v = v3p(A.velocity);
  bool rest = sp_resolve(p, v, v3p(A.shape), A.shape[3] * A.scale[0], A.restitution, v3p(B.shape), B.shape[3], t);
  store(A.position, p); store(A.velocity, v);
  if (rest) A.at_rest = 1;
}
__device__ inline void rs_pill_flat(PgBody &A, PgBody &B, float t) {   // resolution.c:628-717
  V3 vb; body_advance(A, t, vb);
  Flat f = flat_for_pill(B);
  Pill p = pill_from_body(A);
  V3 cp = pill_closest_to_flat(p, f);
  float s = dot(cp, f.n) - f.d;
  float pen = (s < p.r) ? (float)((double)(p.r - s) + 0.001) : 0.0f;   // [f64]
  if (pen <= 0.0f) return;
  store(A.position, add(v3p(A.position), scale(f.n, pen)));
  store(A.velocity, reflect(vb, f.n, A.restitution));
  body_rest_rule(A, f.n);
}
// resolution_functions table, resolution.c:8-18.
__device__ inline bool resolve(PgBody &A, PgBody &B, float t) {
  switch (A.collider_type * 4 + B.collider_type) {
    case PG_AABB * 4 + PG_AABB: rs_box_box(A, B, t); return true;
    case PG_AABB * 4 + PG_SPHERE: rs_box_ball(A, B, t); return true;
    case PG_AABB * 4 + PG_CAPSULE: rs_box_pill(A, B, t); return true;
    case PG_AABB * 4 + PG_PLANE: rs_box_flat(A, B, t); return true;
    case PG_SPHERE * 4 + PG_SPHERE: rs_ball_ball(A, B, t); return true;
    case PG_SPHERE * 4 + PG_PLANE: rs_ball_flat(A, B, t); return true;
    case PG_CAPSULE * 4 + PG_PLANE: rs_pill_flat(A, B, t); return true;
    default: return false;
  }
}

__device__ __forceinline__ int collision_behavior(int ta, int tb) {   // collider.c:7-17; 0 = PHYSICS
  // rows/cols GROUPING, WORLD, ITEM, PLAYER: TRIGGER iff either side is GROUPING or the pair is {ITEM, PLAYER}
  if (ta == PG_ENT_GROUPING || tb == PG_ENT_GROUPING) return 1;
  if ((ta == PG_ENT_ITEM && tb == PG_ENT_PLAYER) || (ta == PG_ENT_PLAYER && tb == PG_ENT_ITEM)) return 1;
  return 0;
}
__device__ __forceinline__ int event_type(int ta, int tb) {   // event.c:14-20
  return ((ta == PG_ENT_ITEM && tb == PG_ENT_PLAYER) || (ta == PG_ENT_PLAYER && tb == PG_ENT_ITEM)) ? 1 : 0;
}

// One visit of world.c:481-547 on two records ALREADY ordered by collider type.
// Returns true when the reference would enqueue an event.
__device__ inline bool visit_ordered(PgBody &A, PgBody &B, float dt) {
  float hit;
  Contact c = contact(0.0f, false);
  if (interval_hit(A, B, 0.0f, dt, &hit)) {
    bool has;
    Contact r = narrow(A, B, dt, &has);
    if (has) c = r;
  }
  if (!(c.hit && c.t >= 0)) return false;
  if (collision_behavior(A.entity_type, B.entity_type) == 0) resolve(A, B, c.t);
  return true;
}

__device__ __forceinline__ void integrate(V3 &p, V3 &v, float dt) {   // world.c:549-553
  v.y -= kGravity * dt;
  p = muladds(v, dt, p);
}
__device__ __forceinline__ float clamp_dt(float dt) { if ((double)dt > kMaxDt) dt = (float)kMaxDt; return dt; }

// scene_node_update (scene.c:1051-1079) for a node whose parent is static:
// local = [R * diag(scale) | pos], world = parent * local (glm_mat4_mul, scalar order).
__device__ inline void node_refresh(PgBody &b) {
  float local[16];
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int r = 0; r < 3; r++) local[c * 4 + r] = b.euler_node[c * 3 + r] * b.scale[c];
    local[c * 4 + 3] = 0.0f * b.scale[c];
  }
  local[12] = b.position[0]; local[13] = b.position[1]; local[14] = b.position[2]; local[15] = 1.0f;
  if (!b.has_parent) {
#pragma unroll
    for (int k = 0; k < 16; k++) b.node_xform[k] = local[k];
    return;
  }
  const float *p = b.parent_xform;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int k = 0; k < 4; k++)
      b.node_xform[c * 4 + k] = p[0 * 4 + k] * local[c * 4 + 0] + p[1 * 4 + k] * local[c * 4 + 1] +
                                p[2 * 4 + k] * local[c * 4 + 2] + p[3 * 4 + k] * local[c * 4 + 3];
}

}  // namespace pgm
