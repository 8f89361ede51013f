/*
 * oracle/crux_oracle.c -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * Plain-C restatement of the Crux physics step: an independent re-derivation of the algorithm in
 * /root/reference/src/physics/{world,distance,narrow_phase,resolution,aabb,collider,sphere,utils}.c
 * used as the CPU checker for the CUDA path.  Every function cites the reference lines it follows.
 * Arithmetic is IEEE binary32 except where the reference's C silently promotes to double (a double
 * literal, or a <math.h> double function such as sqrt/fabs); those places are marked [f64].
 * Built with -ffp-contract=off (no FMA).  Vector helpers restate cglm's scalar definitions
 * (see oracle/shim/cglm/cglm.h for the third-party note).
 *
 * Parity status: pinned bit-for-bit against the compiled reference (oracle/_ref) by
 * tests/test_oracle_vs_ref.py and against committed fixtures in tests/golden/.
 */
#include "crux_oracle.h"

#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define GRAVITY 9.8f                 /* literal repeated in world.c:363,550; resolution.c:28,159,... */
#define MAX_DT 0.01666               /* world.c:18, a double */
#define INTERVAL_EPS 0.000001f       /* world.c:21 */
#define NP_EPS 0.0001                /* narrow_phase.c:6, a double */

typedef float v3[3];

/* ------------------------------------------------------------------ vector helpers (cglm scalar forms) */
static inline float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline float norm3(const float *a) { return sqrtf(dot3(a, a)); }
static inline void add3(const float *a, const float *b, float *d) { d[0] = a[0] + b[0]; d[1] = a[1] + b[1]; d[2] = a[2] + b[2]; }
static inline void sub3(const float *a, const float *b, float *d) { d[0] = a[0] - b[0]; d[1] = a[1] - b[1]; d[2] = a[2] - b[2]; }
static inline void scale3(const float *a, float s, float *d) { d[0] = a[0] * s; d[1] = a[1] * s; d[2] = a[2] * s; }
static inline void muladds3(const float *a, float s, float *d) { d[0] += a[0] * s; d[1] += a[1] * s; d[2] += a[2] * s; }
static inline void mulsubs3(const float *a, float s, float *d) { d[0] -= a[0] * s; d[1] -= a[1] * s; d[2] -= a[2] * s; }
static inline void copy3(const float *a, float *d) { d[0] = a[0]; d[1] = a[1]; d[2] = a[2]; }
static inline void normalize3(float *v) {
  float n = norm3(v);
  if (n < FLT_EPSILON) { v[0] = v[1] = v[2] = 0.0f; return; }
  scale3(v, 1.0f / n, v);
}
static inline float fmax_glm(float a, float b) { if (a > b) return a; return b; }
static inline float fmin_glm(float a, float b) { if (a < b) return a; return b; }
static inline float clamp_glm(float v, float lo, float hi) { return fmin_glm(fmax_glm(v, lo), hi); }

/* column-major 4x4 (m[c*4+r]) times (v,1), xyz of the result -- glm_mat4_mulv3(m, v, 1.0f, dest) */
static void xform_point(const float *m, const float *v, float *d) {
  float r[3];
  for (int k = 0; k < 3; k++) r[k] = m[0 * 4 + k] * v[0] + m[1 * 4 + k] * v[1] + m[2 * 4 + k] * v[2] + m[3 * 4 + k] * 1.0f;
  copy3(r, d);
}
/* glm_decompose_scalev + glm_mat4_pick3 + glm_mat3_scale(1/scale[0]) as every AABB path does
 * (e.g. distance.c:28-35).  rot is column-major 3x3: rot[c*3+r]. */
static void node_decompose(const float *m, float *pos, float *scale, float *rot) {
  const float zero[3] = {0.0f, 0.0f, 0.0f};
  xform_point(m, zero, pos);
  for (int c = 0; c < 3; c++) scale[c] = norm3(&m[c * 4]);
  for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) rot[c * 3 + r] = m[c * 4 + r];
  if (scale[0] != 0.0f) {
    float s = 1.0f / scale[0];
    for (int k = 0; k < 9; k++) rot[k] *= s;
  }
}
static void mat3_mulv(const float *m, const float *v, float *d) {
  float r[3];
  for (int k = 0; k < 3; k++) r[k] = m[0 * 3 + k] * v[0] + m[1 * 3 + k] * v[1] + m[2 * 3 + k] * v[2];
  copy3(r, d);
}

/* glm_euler_xyz upper 3x3 (cglm euler.h), angles in radians; the AABB-plane and capsule-plane
 * paths feed it PhysicsBody.rotation, which holds DEGREES (distance.c:255, :373) -- reproduced. */
void ora_euler_xyz(const float *a, float *rot) {
  float sx = sinf(a[0]), cx = cosf(a[0]);
  float sy = sinf(a[1]), cy = cosf(a[1]);
  float sz = sinf(a[2]), cz = cosf(a[2]);
  float czsx = cz * sx, cxcz = cx * cz, sysz = sy * sz;
  rot[0] = cy * cz;
  rot[1] = czsx * sy + cx * sz;
  rot[2] = -cxcz * sy + sx * sz;
  rot[3] = -cy * sz;
  rot[4] = cxcz - sx * sysz;
  rot[5] = czsx + cx * sysz;
  rot[6] = sy;
  rot[7] = -cy * sx;
  rot[8] = cx * cy;
}

/* ------------------------------------------------------------------ AABB helpers (aabb.c) */

/* AABB_update, aabb.c:59-78.  [f64]: fabs() products are double, accumulated into a float. */
static void aabb_update(const float *src_c, const float *src_e, const float *rot, const float *tr,
                        const float *scale, float *dc, float *de) {
  for (int i = 0; i < 3; i++) {
    dc[i] = tr[i];
    de[i] = 0.0f;
    for (int j = 0; j < 3; j++) {
      dc[i] += rot[j * 3 + i] * src_c[j];
      de[i] = (float)((double)de[i] + fabs((double)rot[j * 3 + i]) * (double)src_e[j] * fabs((double)scale[i]));
    }
  }
}
void ora_aabb_update(const float *s6, const float *rot9, const float *t3, const float *sc3, float *d6) {
  aabb_update(s6, s6 + 3, rot9, t3, sc3, d6, d6 + 3);
}
/* AABB_intersect_AABB, aabb.c:100-105 ('>' so touching counts as intersecting) */
static bool aabb_hits_aabb(const float *ac, const float *ae, const float *bc, const float *be) {
  for (int i = 0; i < 3; i++) if (fabs((double)(ac[i] - bc[i])) > (double)(ae[i] + be[i])) return false;
  return true;
}
int ora_aabb_intersect_aabb(const float *a, const float *b) { return aabb_hits_aabb(a, a + 3, b, b + 3); }
/* projection radius of box extents on a plane normal: e . |n| in double, stored to float
 * (aabb.c:111-114, distance.c:267-270, narrow_phase.c:367-370) */
static float proj_radius_f64(const float *e, const float *n) {
  return (float)((double)e[0] * fabs((double)n[0]) + (double)e[1] * fabs((double)n[1]) + (double)e[2] * fabs((double)n[2]));
}
/* AABB_intersect_plane, aabb.c:108-121 */
int ora_aabb_intersect_plane(const float *box, const float *pl) {
  float r = proj_radius_f64(box + 3, pl);
  float s = dot3(pl, box) - pl[3];
  return fabs((double)s) <= (double)r;
}
/* AABB_merge, aabb.c:8-51 */
void ora_aabb_merge(float *a, int *a_init, const float *b, int b_init) {
  if (!*a_init) { memcpy(a, b, 24); *a_init = 1; return; }
  if (!b_init) return;
  for (int i = 0; i < 3; i++) {
    float lo = fmin_glm(a[i] - a[3 + i], b[i] - b[3 + i]);
    float hi = fmax_glm(a[i] + a[3 + i], b[i] + b[3 + i]);
    a[i] = (lo + hi) * 0.5f;
    a[3 + i] = (hi - lo) * 0.5f;
  }
  *a_init = 1;
}
/* AABB_update_by_vertex, aabb.c:80-95 */
void ora_aabb_update_by_vertex(float *a, const float *v) {
  for (int i = 0; i < 3; i++) {
    if (v[i] > (a[i] + a[3 + i])) {
      float d = v[i] - (a[i] + a[3 + i]);
      a[i] += d * 0.5f; a[3 + i] += d * 0.5f;
    }
    if (v[i] < (a[i] - a[3 + i])) {
      float d = a[i] - a[3 + i] - v[i];
      a[i] -= d * 0.5f; a[3 + i] += d * 0.5f;
    }
  }
}
/* sphere_intersect_sphere, sphere.c:5-12 */
int ora_sphere_intersect_sphere(const float *a, const float *b) {
  v3 d; sub3(a, b, d);
  float rs = a[3] + b[3];
  return dot3(d, d) <= rs * rs;
}
/* distance_point_to_segment, physics/utils.c:33-60 */
float ora_distance_point_to_segment(const float *p, const float *a, const float *b) {
  v3 seg, ap; sub3(b, a, seg); sub3(p, a, ap);
  float d = dot3(seg, ap) / dot3(seg, seg);
  if (d <= 0) return norm3(ap);
  if (d >= 1) { v3 bp; sub3(p, b, bp); return norm3(bp); }
  v3 along, cp, pc; scale3(seg, d, along); add3(a, along, cp); sub3(cp, p, pc);
  return norm3(pc);
}
/* collider.c:7-17 */
int ora_collision_behavior(int ta, int tb) {
  static const int tab[4][4] = {{1, 1, 1, 1}, {1, 0, 0, 0}, {1, 0, 0, 1}, {1, 0, 1, 0}};
  return tab[ta][tb];   /* 0 = PHYSICS, 1 = TRIGGER */
}
/* event.c:14-20 */
int ora_event_type(int ta, int tb) {
  return ((ta == ORA_ENT_ITEM && tb == ORA_ENT_PLAYER) || (ta == ORA_ENT_PLAYER && tb == ORA_ENT_ITEM)) ? 1 : 0;
}

/* ------------------------------------------------------------------ world-space shapes */

typedef struct { v3 c, e; } Box;
typedef struct { v3 c; float r; } Ball;
typedef struct { v3 a, b; float r; } Pill;
typedef struct { v3 n; float d; } Flat;

/* Box from the scene-node transform, optionally advanced by `adv * time` (distance.c:26-37,
 * narrow_phase.c:28-38, resolution.c:41-52).  No node -> all-zero box. */
static void box_from_node(const OraBody *b, const float *adv, float time, Box *out) {
  memset(out, 0, sizeof(*out));
  if (!b->has_node) return;
  v3 pos, sc; float rot[9];
  node_decompose(b->node_xform, pos, sc, rot);
  if (adv) muladds3(adv, time, pos);
  aabb_update(b->shape, b->shape + 3, rot, pos, sc, out->c, out->e);
}
/* Box from PhysicsBody rotation (degrees read as radians) / position / scale:
 * distance.c:253-264, narrow_phase.c:350-360, resolution.c:427-435. */
static void box_from_body(const OraBody *b, const float *pos, Box *out) {
  float rot[9];
  ora_euler_xyz(b->rotation, rot);
  aabb_update(b->shape, b->shape + 3, rot, pos, b->scale, out->c, out->e);
}
/* Sphere from body position (sphere-sphere, sphere-plane paths): distance.c:282-285 */
static void ball_from_body(const OraBody *b, const float *pos, Ball *out) {
  add3(b->shape, pos, out->c);
  out->r = b->shape[3] * b->scale[0];
}
/* Sphere from the scene node (AABB-sphere paths): distance.c:113-125 */
static void ball_from_node(const OraBody *b, const float *adv, float time, Ball *out) {
  memset(out, 0, sizeof(*out));
  if (!b->has_node) return;
  v3 pos, sc; float rot[9];
  node_decompose(b->node_xform, pos, sc, rot);
  if (adv) muladds3(adv, time, pos);
  add3(b->shape, pos, out->c);
  out->r = b->shape[3] * sc[0];
}
/* Capsule end points through the node transform (AABB-capsule paths): distance.c:185-189,209 */
static void pill_from_node(const OraBody *b, Pill *out) {
  memset(out, 0, sizeof(*out));
  if (b->has_node) {
    xform_point(b->node_xform, b->shape, out->a);
    xform_point(b->node_xform, b->shape + 3, out->b);
  }
  out->r = b->shape[6] * b->scale[0];
}
/* Capsule from body scale/rotation/position (capsule-plane paths): distance.c:366-379 */
static void pill_from_body(const OraBody *b, Pill *out) {
  float rot[9];
  scale3(b->shape, b->scale[0], out->a);
  scale3(b->shape + 3, b->scale[0], out->b);
  ora_euler_xyz(b->rotation, rot);
  mat3_mulv(rot, out->a, out->a);
  mat3_mulv(rot, out->b, out->b);
  add3(out->a, b->position, out->a);
  add3(out->b, b->position, out->b);
  out->r = b->shape[6] * b->scale[0];
}
/* Plane seen by a capsule: rotated by the PARENT node's transform (distance.c:342-364).  The
 * reference dereferences parent_node unconditionally; records with has_node && !has_parent are
 * rejected when a world is built. */
static void flat_for_pill(const OraBody *b, Flat *out) {
  memset(out, 0, sizeof(*out));
  if (!b->has_node) return;
  v3 pos, sc; float rot[9];
  node_decompose(b->parent_xform, pos, sc, rot);
  mat3_mulv(rot, b->shape, out->n);
  normalize3(out->n);
  out->d = b->shape[3] + dot3(pos, out->n);
}
/* closest point of a capsule segment to a plane (distance.c:391-400) */
static void pill_closest_to_flat(const Pill *p, const Flat *f, float *cp) {
  float n_dot_a = dot3(f->n, p->a);
  v3 seg; sub3(p->b, p->a, seg);
  float n_dot_seg = dot3(f->n, seg);
  float t = (f->d - n_dot_a) / n_dot_seg;
  if (t <= 0) copy3(p->a, cp);
  else if (t >= 1) copy3(p->b, cp);
  else { v3 v; sub3(p->b, p->a, v); scale3(v, t, v); add3(p->a, v, cp); }   /* glm_vec3_lerp */
}
/* closest points capsule segment <-> box (distance.c:211-238); pq = q - closest */
static void pill_box_pq(const Pill *p, const Box *bx, float *pq) {
  v3 seg, a2c; sub3(p->b, p->a, seg); sub3(bx->c, p->a, a2c);
  float proj = dot3(seg, a2c);
  float t = proj / dot3(seg, seg);
  t = clamp_glm(t, 0.0f, 1.0f);
  v3 cp; copy3(p->a, cp); muladds3(seg, t, cp);
  v3 q;
  for (int i = 0; i < 3; i++) q[i] = clamp_glm(cp[i], bx->c[i] - bx->e[i], bx->c[i] + bx->e[i]);
  sub3(q, cp, pq);
}

/* ------------------------------------------------------------------ distance functions (distance.c) */

/* world.c:584-589 */
static float max_move(const OraBody *b, float t0, float t1) {
  float dt = t1 - t0;
  return norm3(b->velocity) * dt;
}

/* distance.c:20-92 */
static float dist_box_box(const OraBody *A, const OraBody *B, float t) {
  Box a, b; box_from_node(A, A->velocity, t, &a); box_from_node(B, B->velocity, t, &b);
  float d2 = 0.0f;
  for (int i = 0; i < 3; i++) {
    float minA = a.c[i] - a.e[i], maxA = a.c[i] + a.e[i];
    float minB = b.c[i] - b.e[i], maxB = b.c[i] + b.e[i];
    if (minB > maxA) d2 += (minB - maxA) * (minB - maxA);
    else if (maxB < minA) d2 += (minA - maxB) * (minA - maxB);
  }
  return (float)sqrt((double)d2);                                                  /* [f64] */
}
/* distance.c:94-158 */
static float dist_box_ball(const OraBody *A, const OraBody *B, float t) {
  Box bx; Ball s; box_from_node(A, A->velocity, t, &bx); ball_from_node(B, B->velocity, t, &s);
  float d2 = 0.0f;
  for (int i = 0; i < 3; i++) {
    float v = s.c[i], lo = bx.c[i] - bx.e[i], hi = bx.c[i] + bx.e[i];
    if (v < lo) d2 += (lo - v) * (lo - v);
    if (v > hi) d2 += (v - hi) * (v - hi);
  }
  return d2 < (s.r * s.r) ? 0.0f : (float)(sqrt((double)d2) - (double)s.r);          /* [f64] */
}
/* distance.c:160-244 (capsule velocity is NOT applied, :207-208 commented out) */
static float dist_box_pill(const OraBody *A, const OraBody *B, float t) {
  Box bx; Pill p; box_from_node(A, A->velocity, t, &bx); pill_from_node(B, &p);
  v3 pq; pill_box_pq(&p, &bx, pq);
  return fmax_glm(norm3(pq) - p.r, 0);
}
/* distance.c:246-276 (one-sided: s - r) */
static float dist_box_flat(const OraBody *A, const OraBody *B, float t) {
  v3 pos; copy3(A->position, pos); muladds3(A->velocity, t, pos);
  Box bx; box_from_body(A, pos, &bx);
  float r = proj_radius_f64(bx.e, B->shape);                                       /* [f64] */
  float s = dot3(B->shape, bx.c) - B->shape[3];
  return fmax_glm(s - r, 0);
}
/* distance.c:278-299 */
static float dist_ball_ball(const OraBody *A, const OraBody *B, float t) {
  Ball a, b;
  ball_from_body(A, A->position, &a); muladds3(A->velocity, t, a.c);
  ball_from_body(B, B->position, &b); muladds3(B->velocity, t, b.c);
  v3 d; sub3(a.c, b.c, d);
  float d2 = dot3(d, d);
  float rs = a.r + b.r;
  return d2 < (rs * rs) ? 0.0f : (float)(sqrt((double)d2) - (double)rs);            /* [f64] */
}
/* distance.c:305-321 (two-sided: |s| - r) */
static float dist_ball_flat(const OraBody *A, const OraBody *B, float t) {
  Ball a; ball_from_body(A, A->position, &a); muladds3(A->velocity, t, a.c);
  float s = dot3(a.c, B->shape) - B->shape[3];
  float d = (float)(fabs((double)s) - (double)a.r);                                 /* [f64] */
  return fmax_glm(d, 0);
}
/* distance.c:327-410 (time unused: the capsule is taken at body->position) */
static float dist_pill_flat(const OraBody *A, const OraBody *B, float t) {
  (void)t;
  Pill p; Flat f; flat_for_pill(B, &f); pill_from_body(A, &p);
  v3 cp; pill_closest_to_flat(&p, &f, cp);
  float s = dot3(cp, f.n) - f.d;
  float d = (float)(fabs((double)s) - (double)p.r);                                 /* [f64] */
  return fmax_glm(d, 0);
}

/* Table distance_functions, distance.c:7-17.  Returns 0 for a missing entry (world.c:594-598).
 * [SPHERE][CAPSULE] is an empty function in the reference (distance.c:301-303, undefined return
 * value) and [CAPSULE][CAPSULE] returns 0 (:323-325); both are reported as 0 = "always close". */
static float min_dist(const OraBody *A, const OraBody *B, float t) {
  switch (A->collider_type * 4 + B->collider_type) {
    case ORA_AABB * 4 + ORA_AABB: return dist_box_box(A, B, t);
    case ORA_AABB * 4 + ORA_SPHERE: return dist_box_ball(A, B, t);
    case ORA_AABB * 4 + ORA_CAPSULE: return dist_box_pill(A, B, t);
    case ORA_AABB * 4 + ORA_PLANE: return dist_box_flat(A, B, t);
    case ORA_SPHERE * 4 + ORA_SPHERE: return dist_ball_ball(A, B, t);
    case ORA_SPHERE * 4 + ORA_PLANE: return dist_ball_flat(A, B, t);
    case ORA_CAPSULE * 4 + ORA_PLANE: return dist_pill_flat(A, B, t);
    default: return 0.0f;
  }
}

/* interval_collision, world.c:557-582 */
static bool interval_hit(const OraBody *A, const OraBody *B, float t0, float t1, float *hit) {
  float mm = max_move(A, t0, t1) + max_move(B, t0, t1);
  if (min_dist(A, B, t0) > mm) return false;
  if (min_dist(A, B, t1) > mm) return false;
  if (t1 - t0 < INTERVAL_EPS) { *hit = t0; return true; }
  float mid = (t0 + t1) * 0.5f;
  if (interval_hit(A, B, t0, mid, hit)) return true;
  return interval_hit(A, B, mid, t1, hit);
}

/* ------------------------------------------------------------------ narrow phase (narrow_phase.c) */

typedef struct { float t; bool hit; } Contact;
static inline Contact contact(float t, bool hit) { Contact c; c.t = t; c.hit = hit; return c; }

/* narrow_phase.c:21-141 */
static Contact np_box_box(const OraBody *A, const OraBody *B, float dt) {
  Box a, b; box_from_node(A, NULL, 0, &a); box_from_node(B, NULL, 0, &b);
  if (aabb_hits_aabb(a.c, a.e, b.c, b.e)) return contact(0, true);
  v3 rv; sub3(A->velocity, B->velocity, rv);
  float t_first = 0.0f, t_last = dt;
  for (int i = 0; i < 3; i++) {
    float minA = a.c[i] - a.e[i], maxA = a.c[i] + a.e[i];
    float minB = b.c[i] - b.e[i], maxB = b.c[i] + b.e[i];
    if (rv[i] < 0.0f) {
      if (maxA < minB) return contact(-1, false);
      if (minA > maxB) t_first = fmaxf((maxB - minA) / rv[i], t_first);
      if (maxA > minB) t_last = fminf((minB - maxA) / rv[i], t_last);
    }
    if (rv[i] > 0.0f) {
      if (minA > maxB) return contact(-1, false);
      if (maxA < minB) t_first = fmaxf((minB - maxA) / rv[i], t_first);
      if (minA < maxB) t_last = fminf((maxB - minA) / rv[i], t_last);
    }
    if (t_first > t_last) return contact(-1, false);
  }
  return contact(t_first, true);
}
/* ray_intersect_AABB, narrow_phase.c:702-739 */
static bool ray_box(const float *p, const float *d, const float *c, const float *e, float *hit, float t_end) {
  *hit = 0.0f;
  float t_max = t_end;
  for (int i = 0; i < 3; i++) {
    float lo = c[i] - e[i], hi = c[i] + e[i];
    if ((double)fabsf(d[i]) < NP_EPS) {
      if (p[i] < lo || p[i] > hi) return false;
    } else {
      float t1 = (lo - p[i]) / d[i];
      float t2 = (hi - p[i]) / d[i];
      if (t1 > t2) { float tmp = t1; t1 = t2; t2 = tmp; }
      if (t1 > *hit) *hit = t1;
      if (t2 < t_max) t_max = t2;
      if (*hit > t_max) return false;
    }
  }
  return true;
}
/* narrow_phase.c:143-196 ("lazy version": ray vs box inflated by the radius) */
static Contact np_box_ball(const OraBody *A, const OraBody *B, float dt) {
  Box bx; Ball s; box_from_node(A, NULL, 0, &bx); ball_from_node(B, NULL, 0, &s);
  v3 e; for (int i = 0; i < 3; i++) e[i] = bx.e[i] + s.r;
  v3 rv; sub3(B->velocity, A->velocity, rv);
  float tmin;
  bool hit = ray_box(s.c, rv, bx.c, e, &tmin, dt);
  if (!hit || tmin > dt) return contact(-1, false);
  return contact(tmin, true);
}
/* narrow_phase.c:233-342 */
static Contact np_box_pill(const OraBody *A, const OraBody *B, float dt) {
  Box bx; Pill p; box_from_node(A, NULL, 0, &bx); pill_from_node(B, &p);
  v3 pq; pill_box_pq(&p, &bx, pq);
  float distance = norm3(pq) - p.r;
  normalize3(pq);
  float pq_dot_v = dot3(B->velocity, pq);
  float disc = distance * pq_dot_v;
  if (disc == 0) return distance < 0 ? contact(0, true) : contact(-1, false);
  float t = distance / pq_dot_v;
  bool hit = (t >= 0 && t <= dt);
  if (!hit) { if (distance <= 0) return contact(0, true); return contact(-1, false); }
  return contact(t, true);
}
/* narrow_phase.c:344-457 */
static Contact np_box_flat(const OraBody *A, const OraBody *B, float dt) {
  Box bx; box_from_body(A, A->position, &bx);
  const float *n = B->shape;
  float r = proj_radius_f64(bx.e, n);
  float s = dot3(n, bx.c) - B->shape[3];
  float ndv = dot3(n, A->velocity);
  bool inside = fabs((double)s) <= (double)r;
  if (ndv == 0) return inside ? contact(0, true) : contact(-1, false);
  Contact c;
  if (ndv < 0) { c.t = (r - s) / -ndv; c.hit = (c.t >= 0 && c.t <= dt); }
  else if (inside) { c.t = 0; c.hit = true; }
  else { c.t = (r + s) / -ndv; c.hit = (c.t >= 0 && c.t <= dt); }
  if (c.hit) c.t = fmaxf(0.0f, fminf(c.t, dt));
  if (!c.hit) { if (inside) return contact(0, true); c.t = -1; }
  return c;
}
/* narrow_phase.c:459-520.  The three "no collision" blocks (:493-512) fall through; the function
 * always ends with the quadratic root [f64] and colliding=true unless the spheres overlap. */
static Contact np_ball_ball(const OraBody *A, const OraBody *B, float dt) {
  (void)dt;
  Ball a, b; ball_from_body(A, A->position, &a); ball_from_body(B, B->position, &b);
  v3 s, v; sub3(a.c, b.c, s); sub3(A->velocity, B->velocity, v);
  float r = a.r + b.r;
  float c = dot3(s, s) - (r * r);
  if (c < 0.0f) return contact(0, true);
  float aa = dot3(v, v);
  float bb = dot3(v, s);
  float d = (bb * bb) - aa * c;
  float t = (float)(((double)(-bb) - sqrt((double)d)) / (double)aa);                 /* [f64] */
  return contact(t, true);
}
/* shared tail of narrow_phase.c:553-585 and :663-695.  The fallback test is `abs(s <= r)`, the
 * abs() of a comparison, i.e. one-sided (narrow_phase.c:577, :687). */
static Contact np_plane_tail(float s, float ndv, float r, float dt) {
  float disc = s * ndv;
  if (disc == 0) return fabs((double)s) <= (double)r ? contact(0, true) : contact(-1, false);
  float t = disc < 0 ? (r - s) / ndv : (-r - s) / ndv;
  if (t >= 0 && t <= dt) return contact(t, true);
  if (s <= r) return contact(0, true);
  return contact(-1, false);
}
/* narrow_phase.c:526-588 */
static Contact np_ball_flat(const OraBody *A, const OraBody *B, float dt) {
  Ball a; ball_from_body(A, A->position, &a);
  float s = dot3(a.c, B->shape) - B->shape[3];
  float ndv = dot3(A->velocity, B->shape);
  return np_plane_tail(s, ndv, a.r, dt);
}
/* narrow_phase.c:594-698 */
static Contact np_pill_flat(const OraBody *A, const OraBody *B, float dt) {
  Pill p; Flat f; flat_for_pill(B, &f); pill_from_body(A, &p);
  v3 cp; pill_closest_to_flat(&p, &f, cp);
  float s = dot3(cp, f.n) - f.d;
  float ndv = dot3(A->velocity, f.n);
  return np_plane_tail(s, ndv, p.r, dt);
}

/* Table narrow_phase_functions, narrow_phase.c:8-19.  Returns false in *has for the pairs whose
 * reference function is NULL or an empty body (sphere-capsule :522-524, capsule-capsule :590-592,
 * undefined behaviour upstream; rejected here). */
static Contact narrow(const OraBody *A, const OraBody *B, float dt, bool *has) {
  *has = true;
  switch (A->collider_type * 4 + B->collider_type) {
    case ORA_AABB * 4 + ORA_AABB: return np_box_box(A, B, dt);
    case ORA_AABB * 4 + ORA_SPHERE: return np_box_ball(A, B, dt);
    case ORA_AABB * 4 + ORA_CAPSULE: return np_box_pill(A, B, dt);
    case ORA_AABB * 4 + ORA_PLANE: return np_box_flat(A, B, dt);
    case ORA_SPHERE * 4 + ORA_SPHERE: return np_ball_ball(A, B, dt);
    case ORA_SPHERE * 4 + ORA_PLANE: return np_ball_flat(A, B, dt);
    case ORA_CAPSULE * 4 + ORA_PLANE: return np_pill_flat(A, B, dt);
    default: *has = false; return contact(0, false);
  }
}

/* ------------------------------------------------------------------ resolution (resolution.c) */

/* "advance to hit time under gravity": resolution.c:30-34 and its seven siblings */
static void advance_to_hit(OraBody *b, float t, float *v_before) {
  copy3(b->velocity, v_before);
  v_before[1] -= GRAVITY * t;
  muladds3(v_before, t, b->position);
}
/* reflect v_before about n with restitution, resolution.c:606-609 */
static void reflect(const float *v_before, const float *n, float restitution, float *v_out) {
  float vdn = dot3(v_before, n);
  v3 refl; scale3(n, -2.0f * vdn * restitution, refl);
  add3(v_before, refl, v_out);
}
/* swap normal velocity components (equal-mass elastic), resolution.c:128-142 */
static void exchange_normal(const float *vbA, const float *vbB, const float *n, float *vA, float *vB) {
  float a = dot3(vbA, n), b = dot3(vbB, n);
  v3 na, nb; scale3(n, a, na); scale3(n, b, nb);
  sub3(vbA, na, vA); add3(vA, nb, vA);
  sub3(vbB, nb, vB); add3(vB, na, vB);
}
/* resting rule, resolution.c:613-618: [f64] compare against 0.5 */
static void maybe_rest(OraBody *b, const float *n) {
  const float down[3] = {0.0f, -1.0f, 0.0f};
  float vdn = dot3(n, b->velocity);
  if ((double)vdn < 0.5 && dot3(n, down) < 0) {
    b->velocity[0] = b->velocity[1] = b->velocity[2] = 0.0f;
    b->at_rest = 1;
  }
}

/* resolution.c:21-150.  velocity_before of a static body is read uninitialised upstream (:62,:102);
 * it is 0 here, as in the reference build with -ftrivial-auto-var-init=zero. */
static void rs_box_box(OraBody *A, OraBody *B, float t) {
  v3 vbA = {0, 0, 0}, vbB = {0, 0, 0};
  if (A->dynamic) advance_to_hit(A, t, vbA);
  if (B->dynamic) advance_to_hit(B, t, vbB);
  Box a, b; box_from_node(A, vbA, t, &a); box_from_node(B, vbB, t, &b);
  v3 cd; sub3(b.c, a.c, cd);
  v3 sep = {0, 0, 0};
  float min_pen = FLT_MAX; int axis = 0;
  for (int i = 0; i < 3; i++) {
    float s = (float)fabs((double)cd[i]);
    float r = a.e[i] + b.e[i];
    float pen = r - s;
    if (pen < min_pen) { min_pen = pen; axis = i; }
  }
  if (min_pen <= 0.0f) return;
  sep[axis] = (cd[axis] < 0 ? -1.0f : 1.0f);
  v3 n; copy3(sep, n); normalize3(n);
  if (A->dynamic && !B->dynamic) {
    mulsubs3(sep, min_pen, A->position);
    reflect(vbA, n, A->restitution, A->velocity);
  } else if (!A->dynamic && B->dynamic) {
    muladds3(sep, min_pen, B->position);
    float vdn = dot3(vbB, n);
    v3 refl; scale3(n, -2.0f * vdn * B->restitution, refl);
    add3(vbA, refl, B->velocity);          /* sic: velocity_before_A, resolution.c:118 */
  } else {
    v3 rv; sub3(vbB, vbA, rv);
    if (dot3(rv, n) < 0.0f) {
      muladds3(sep, min_pen * 0.5f, B->position);
      mulsubs3(sep, min_pen * 0.5f, A->position);
      exchange_normal(vbA, vbB, n, A->velocity, B->velocity);
    } else { copy3(vbA, A->velocity); copy3(vbB, B->velocity); }
  }
}
/* resolution.c:152-302 (same uninitialised read, :182; 0 here) */
static void rs_box_ball(OraBody *A, OraBody *B, float t) {
  v3 vbA = {0, 0, 0}, vbB = {0, 0, 0};
  if (A->dynamic) advance_to_hit(A, t, vbA);
  if (B->dynamic) advance_to_hit(B, t, vbB);
  Box bx; Ball s; box_from_node(A, vbA, t, &bx); ball_from_node(B, vbB, t, &s);
  v3 closest;
  for (int i = 0; i < 3; i++) {
    float lo = bx.c[i] - bx.e[i], hi = bx.c[i] + bx.e[i];
    if (lo > s.c[i]) closest[i] = lo;
    else if (hi > s.c[i]) closest[i] = s.c[i];
    else closest[i] = hi;
  }
  v3 diff; sub3(s.c, closest, diff);
  float d = norm3(diff);
  float pen = (d < s.r) ? (float)((double)(s.r - d) + 0.001) : 0.0f;                 /* [f64] */
  if (pen <= 0) return;
  v3 n; copy3(diff, n); normalize3(n);
  if (A->dynamic && !B->dynamic) {
    mulsubs3(n, pen * 0.5f, A->position);
    reflect(vbA, n, A->restitution, A->velocity);
  } else if (!A->dynamic && B->dynamic) {
    muladds3(n, pen * 0.5f, B->position);
    reflect(vbB, n, B->restitution, B->velocity);
  } else {
    v3 rv; sub3(vbB, vbA, rv);
    if (dot3(rv, n) < 0.0f) {
      muladds3(n, pen * 0.5f, B->position);
      mulsubs3(n, pen * 0.5f, A->position);
      exchange_normal(vbA, vbB, n, A->velocity, B->velocity);
    } else { copy3(vbA, A->velocity); copy3(vbB, B->velocity); }
  }
}
/* resolution.c:304-414 (only the capsule moves; skin is 0.001f here, :386) */
static void rs_box_pill(OraBody *A, OraBody *B, float t) {
  v3 vb; advance_to_hit(B, t, vb);
  Box bx; Pill p; box_from_node(A, NULL, 0, &bx); pill_from_node(B, &p);
  if (B->has_node) { muladds3(vb, t, p.a); muladds3(vb, t, p.b); }
  v3 pq; pill_box_pq(&p, &bx, pq);
  float distance = norm3(pq);
  normalize3(pq);
  float pen = distance < p.r ? (p.r - distance) + 0.001f : 0.0f;
  if (pen <= 0.0f) return;
  v3 corr; scale3(pq, pen, corr);
  sub3(B->position, corr, B->position);
  reflect(vb, pq, B->restitution, B->velocity);
  pq[0] = -pq[0]; pq[1] = -pq[1]; pq[2] = -pq[2];
  maybe_rest(B, pq);
}
/* resolution.c:416-466 */
static void rs_box_flat(OraBody *A, OraBody *B, float t) {
  v3 vb; advance_to_hit(A, t, vb);
  Box bx; box_from_body(A, A->position, &bx);
  const float *n = B->shape;
  float r = bx.e[0] * fabsf(n[0]) + bx.e[1] * fabsf(n[1]) + bx.e[2] * fabsf(n[2]);   /* fabsf: float */
  float s = dot3(n, bx.c) - B->shape[3];
  float pen = (s < r) ? (float)((double)(r - s) + 0.001) : 0.0f;                     /* [f64] */
  if (pen > 0.0f) { v3 corr; scale3(n, pen, corr); add3(A->position, corr, A->position); }
  reflect(vb, n, A->restitution, A->velocity);
  maybe_rest(A, n);
}
/* resolution.c:467-570 (ignores restitution / dynamic / at_rest) */
static void rs_ball_ball(OraBody *A, OraBody *B, float t) {
  v3 vbA, vbB;
  copy3(A->velocity, vbA); copy3(B->velocity, vbB);
  vbA[1] -= GRAVITY * t; vbB[1] -= GRAVITY * t;
  muladds3(vbA, t, A->position); muladds3(vbB, t, B->position);
  Ball a, b; ball_from_body(A, A->position, &a); ball_from_body(B, B->position, &b);
  v3 diff; sub3(b.c, a.c, diff);
  float s = norm3(diff);
  float r = a.r + b.r;
  float pen = (s < r) ? (float)((double)(r - s) + 0.001) : 0.0f;                     /* [f64] */
  v3 n; copy3(diff, n); normalize3(n);
  if (pen <= 0.0f) return;
  v3 rv; sub3(vbB, vbA, rv);
  if (dot3(rv, n) < 0.0f) {
    muladds3(n, pen / 2, B->position);
    mulsubs3(n, pen / 2, A->position);
    exchange_normal(vbA, vbB, n, A->velocity, B->velocity);
  } else { copy3(vbA, A->velocity); copy3(vbB, B->velocity); }
}
/* resolution.c:576-622 */
static void rs_ball_flat(OraBody *A, OraBody *B, float t) {
  v3 vb; advance_to_hit(A, t, vb);
  Ball a; ball_from_body(A, A->position, &a);
  const float *n = B->shape;
  float s = dot3(a.c, n) - B->shape[3];
  float pen = (s < a.r) ? (float)((double)(a.r - s) + 0.001) : 0.0f;                 /* [f64] */
  if (pen > 0.0f) { v3 corr; scale3(n, pen, corr); add3(A->position, corr, A->position); }
  reflect(vb, n, A->restitution, A->velocity);
  maybe_rest(A, n);
}
/* resolution.c:628-717 */
static void rs_pill_flat(OraBody *A, OraBody *B, float t) {
  v3 vb; advance_to_hit(A, t, vb);
  Pill p; Flat f; flat_for_pill(B, &f); pill_from_body(A, &p);
  v3 cp; pill_closest_to_flat(&p, &f, cp);
  float s = dot3(cp, f.n) - f.d;
  float pen = (s < p.r) ? (float)((double)(p.r - s) + 0.001) : 0.0f;                 /* [f64] */
  if (pen <= 0.0f) return;
  v3 corr; scale3(f.n, pen, corr); add3(A->position, corr, A->position);
  reflect(vb, f.n, A->restitution, A->velocity);
  maybe_rest(A, f.n);
}
/* Table resolution_functions, resolution.c:8-18 */
static bool resolve(OraBody *A, OraBody *B, float t) {
  switch (A->collider_type * 4 + B->collider_type) {
    case ORA_AABB * 4 + ORA_AABB: rs_box_box(A, B, t); return true;
    case ORA_AABB * 4 + ORA_SPHERE: rs_box_ball(A, B, t); return true;
    case ORA_AABB * 4 + ORA_CAPSULE: rs_box_pill(A, B, t); return true;
    case ORA_AABB * 4 + ORA_PLANE: rs_box_flat(A, B, t); return true;
    case ORA_SPHERE * 4 + ORA_SPHERE: rs_ball_ball(A, B, t); return true;
    case ORA_SPHERE * 4 + ORA_PLANE: rs_ball_flat(A, B, t); return true;
    case ORA_CAPSULE * 4 + ORA_PLANE: rs_pill_flat(A, B, t); return true;
    default: return false;
  }
}

/* ------------------------------------------------------------------ world + step (world.c) */

typedef struct World {
  OraBody *cls[3];
  int n[3], cap[3];
  OraEvent *events;
  int n_events, cap_events, step_index, record;
  /* grid scratch (ora_world_step_grid / ora_world_pairs) */
  int *cell_of, *cell_start, *order;
  float *p0;
  int n_cells_alloc, n_alloc;
} World;

static int g_threads = 1;
void ora_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

void *ora_world_create(void) {
  World *w = (World *)calloc(1, sizeof(World));
  w->record = 1;
  return w;
}
void ora_world_destroy(void *h) {
  World *w = (World *)h;
  for (int c = 0; c < 3; c++) free(w->cls[c]);
  free(w->events); free(w->cell_of); free(w->cell_start); free(w->order); free(w->p0);
  free(w);
}
/* physics_add_body / physics_add_player, world.c:41-142: append; static velocity forced to 0
 * (:51); players get restitution 0 and dynamic=false (:137, calloc). */
int ora_world_add_many(void *h, const OraBody *b, int n, int as_player) {
  World *w = (World *)h;
  int last = -1;
  for (int i = 0; i < n; i++) {
    int c = as_player ? 2 : (b[i].dynamic ? 1 : 0);
    if (w->n[c] == w->cap[c]) {
      w->cap[c] = w->cap[c] ? w->cap[c] * 2 : 128;
      w->cls[c] = (OraBody *)realloc(w->cls[c], (size_t)w->cap[c] * sizeof(OraBody));
    }
    OraBody *d = &w->cls[c][w->n[c]];
    *d = b[i];
    if (c == 0) d->velocity[0] = d->velocity[1] = d->velocity[2] = 0.0f;
    if (c == 2) { d->restitution = 0.0f; d->dynamic = 0; }
    last = w->n[c]++;
  }
  return last;
}
/* physics_remove_body, world.c:147-172: swap with last, pop */
int ora_world_remove(void *h, int c, int index) {
  World *w = (World *)h;
  if (index < 0 || index >= w->n[c]) return -1;
  w->cls[c][index] = w->cls[c][w->n[c] - 1];
  w->n[c]--;
  return 0;
}
void ora_world_counts(void *h, int *out) { World *w = (World *)h; out[0] = w->n[0]; out[1] = w->n[1]; out[2] = w->n[2]; }
void ora_world_read(void *h, int c, OraBody *out) { World *w = (World *)h; memcpy(out, w->cls[c], (size_t)w->n[c] * sizeof(OraBody)); }
void ora_world_poke(void *h, int c, int i, const OraBody *o) {
  World *w = (World *)h; OraBody *b = &w->cls[c][i];
  copy3(o->position, b->position); copy3(o->rotation, b->rotation); copy3(o->velocity, b->velocity);
  b->at_rest = o->at_rest != 0;
}
void ora_world_record_events(void *h, int on) { ((World *)h)->record = on; }
int ora_events_count(void *h) { return ((World *)h)->n_events; }
void ora_events_read(void *h, OraEvent *out) { World *w = (World *)h; memcpy(out, w->events, (size_t)w->n_events * sizeof(OraEvent)); }
void ora_events_clear(void *h) { World *w = (World *)h; w->n_events = 0; w->step_index = 0; }

static void emit(World *w, int etype, int ac, int ai, int bc, int bi) {
  if (!w->record) return;
  if (w->n_events == w->cap_events) {
    w->cap_events = w->cap_events ? w->cap_events * 2 : 1024;
    w->events = (OraEvent *)realloc(w->events, (size_t)w->cap_events * sizeof(OraEvent));
  }
  OraEvent *e = &w->events[w->n_events++];
  e->step = w->step_index; e->event_type = etype;
  e->a_class = ac; e->a_index = ai; e->b_class = bc; e->b_index = bi;
}

/* scene_node_update, scene.c:1051-1079: local = T(pos) * R_xyz(rad(rot)) * S(scale); world =
 * parent * local (glm_mat4_mul) or local.  T*R is exact (identity factors); S scales columns. */
void ora_node_transform(const float *pos, const float *rot_deg, const float *scale,
                        const float *parent, float *world) {
  const float pi = 3.14159265358979323846264338327950288f;
  float rad[3] = {rot_deg[0] * pi / 180.0f, rot_deg[1] * pi / 180.0f, rot_deg[2] * pi / 180.0f};
  float r[9]; ora_euler_xyz(rad, r);
  float local[16];
  /* glm_mul(translate, rotation): upper 3x3 = 1*r + 0*.. + 0*.. ; column 3 = 0+0+0 + t */
  for (int c = 0; c < 3; c++) {
    for (int k = 0; k < 3; k++) {
      float tk[3] = {k == 0 ? 1.0f : 0.0f, k == 1 ? 1.0f : 0.0f, k == 2 ? 1.0f : 0.0f};
      local[c * 4 + k] = tk[0] * r[c * 3 + 0] + tk[1] * r[c * 3 + 1] + tk[2] * r[c * 3 + 2];
    }
    local[c * 4 + 3] = 0.0f;
  }
  for (int k = 0; k < 3; k++) {
    float tk[3] = {k == 0 ? 1.0f : 0.0f, k == 1 ? 1.0f : 0.0f, k == 2 ? 1.0f : 0.0f};
    local[12 + k] = tk[0] * 0.0f + tk[1] * 0.0f + tk[2] * 0.0f + pos[k];
  }
  local[15] = 1.0f;
  for (int c = 0; c < 3; c++) for (int k = 0; k < 4; k++) local[c * 4 + k] = local[c * 4 + k] * scale[c];
  if (!parent) { memcpy(world, local, 64); return; }
  float out[16];
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++)
      out[c * 4 + k] = parent[0 * 4 + k] * local[c * 4 + 0] + parent[1 * 4 + k] * local[c * 4 + 1] +
                       parent[2 * 4 + k] * local[c * 4 + 2] + parent[3 * 4 + k] * local[c * 4 + 3];
  memcpy(world, out, 64);
}
void ora_world_refresh_nodes(void *h) {
  World *w = (World *)h;
  for (int c = 0; c < 3; c++)
    for (int i = 0; i < w->n[c]; i++) {
      OraBody *b = &w->cls[c][i];
      if (!b->has_node) continue;
      ora_node_transform(b->position, b->rotation, b->scale, b->has_parent ? b->parent_xform : NULL, b->node_xform);
    }
}

/* One (A,B) visit, identical in all four passes (world.c:481-547): order by collider type,
 * broadphase gate, narrowphase, behaviour lookup, resolution, event. */
static bool visit(World *w, OraBody *A, int ac, int ai, OraBody *B, int bc, int bi, float dt) {
  if (A->collider_type > B->collider_type) {
    OraBody *t = A; A = B; B = t;
    int ti = ai; ai = bi; bi = ti; ti = ac; ac = bc; bc = ti;
  }
  float hit;
  Contact c = contact(0, false);
  if (interval_hit(A, B, 0, dt, &hit)) {
    bool has;
    Contact r = narrow(A, B, dt, &has);
    if (has) c = r;
  }
  if (!(c.hit && c.t >= 0)) return false;
  if (ora_collision_behavior(A->entity_type, B->entity_type) == 0) resolve(A, B, c.t);
  if (w) emit(w, ora_event_type(A->entity_type, B->entity_type), ac, ai, bc, bi);
  return true;
}

/* integration, world.c:362-366 and :549-553 */
static void integrate(OraBody *b, float dt) {
  if (b->at_rest) return;
  b->velocity[1] -= GRAVITY * dt;
  muladds3(b->velocity, dt, b->position);
}

static float clamp_dt(float dt) { if ((double)dt > MAX_DT) dt = (float)MAX_DT; return dt; }   /* world.c:175-177 */

/* physics_step, world.c:174-555 */
static void step_all_pairs(World *w, float dt) {
  dt = clamp_dt(dt);
  OraBody *S = w->cls[0], *D = w->cls[1], *P = w->cls[2];
  int ns = w->n[0], nd = w->n[1], np = w->n[2];
  for (int i = 0; i < np; i++)                                   /* pass 1, world.c:180-274 */
    for (int j = 0; j < ns; j++) visit(w, &P[i], 2, i, &S[j], 0, j, dt);
  for (int i = 0; i < np; i++) {                                 /* pass 2, world.c:277-367 */
    for (int j = 0; j < nd; j++) visit(w, &P[i], 2, i, &D[j], 1, j, dt);
    integrate(&P[i], dt);
  }
  for (int i = 0; i < nd; i++)                                   /* pass 3, world.c:369-471 */
    for (int j = 0; j < ns; j++) visit(w, &D[i], 1, i, &S[j], 0, j, dt);
  for (int i = 0; i < nd; i++) {                                 /* pass 4, world.c:473-554 */
    for (int j = 0; j < nd; j++) {
      if (i == j) continue;
      visit(w, &D[i], 1, i, &D[j], 1, j, dt);
    }
    integrate(&D[i], dt);
  }
}

void ora_world_step(void *h, float dt, int nsteps, int refresh_nodes) {
  World *w = (World *)h;
  for (int s = 0; s < nsteps; s++) {
    step_all_pairs(w, dt);
    if (refresh_nodes) ora_world_refresh_nodes(h);
    w->step_index++;
  }
}

/* ------------------------------------------------------------------ grid-accelerated step
 *
 * Same visits in the same order as step_all_pairs, minus visits that provably return "no hit".
 * interval_collision(A,B,0,dt) needs dist(A,B,0) <= (|vA|+|vB|) dt at the moment of the visit
 * (world.c:559-565).  If throughout the step every dynamic sphere k stays within D of its
 * start-of-step centre and keeps |v_k| <= S, a visit (i,j) can only pass when the start-of-step
 * centres are within R = rmax2 + 2 S dt + 2 D.  Bodies are binned by start-of-step centre into
 * cells of edge >= R; row i visits, in increasing j, the bodies of its 27 neighbouring cells.  S
 * and D are checked after every mutation; a violation restores the step's start state and re-runs
 * it with step_all_pairs, so the result is always the reference's.  Pass 3 (statics) is unchanged.
 */
static int cmp_int(const void *a, const void *b) { int x = *(const int *)a, y = *(const int *)b; return (x > y) - (x < y); }

typedef struct Grid { float org[3]; float inv_h; int dim[3]; long long n_cells; } Grid;

static bool grid_build(World *w, float R, Grid *g) {
  int nd = w->n[1];
  OraBody *D = w->cls[1];
  if (nd > w->n_alloc) {
    w->n_alloc = nd;
    w->cell_of = (int *)realloc(w->cell_of, sizeof(int) * (size_t)nd);
    w->order = (int *)realloc(w->order, sizeof(int) * (size_t)nd);
    w->p0 = (float *)realloc(w->p0, sizeof(float) * 6 * (size_t)nd);
  }
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < nd; i++) {
    float *c = &w->p0[6 * (size_t)i];
    add3(D[i].shape, D[i].position, c);
    copy3(D[i].velocity, c + 3);
    for (int k = 0; k < 3; k++) {
      if (!(c[k] == c[k]) || fabsf(c[k]) > 1e18f) return false;
      if (c[k] < lo[k]) lo[k] = c[k];
      if (c[k] > hi[k]) hi[k] = c[k];
    }
  }
  g->inv_h = 1.0f / R;
  g->n_cells = 1;
  for (int k = 0; k < 3; k++) {
    g->org[k] = lo[k];
    double span = ((double)hi[k] - (double)lo[k]) / (double)R;
    if (span > 2000.0) return false;
    g->dim[k] = (int)span + 1;
    g->n_cells *= g->dim[k];
  }
  if (g->n_cells > (1LL << 28)) return false;
  if (g->n_cells + 1 > w->n_cells_alloc) {
    w->n_cells_alloc = (int)g->n_cells + 1;
    w->cell_start = (int *)realloc(w->cell_start, sizeof(int) * (size_t)w->n_cells_alloc);
  }
  memset(w->cell_start, 0, sizeof(int) * (size_t)(g->n_cells + 1));
  for (int i = 0; i < nd; i++) {
    const float *c = &w->p0[6 * (size_t)i];
    int ix[3];
    for (int k = 0; k < 3; k++) {
      ix[k] = (int)((c[k] - g->org[k]) * g->inv_h);
      if (ix[k] < 0) ix[k] = 0;
      if (ix[k] >= g->dim[k]) ix[k] = g->dim[k] - 1;
    }
    int cell = (ix[2] * g->dim[1] + ix[1]) * g->dim[0] + ix[0];
    w->cell_of[i] = cell;
    w->cell_start[cell + 1]++;
  }
  for (long long c = 0; c < g->n_cells; c++) w->cell_start[c + 1] += w->cell_start[c];
  /* counting sort, stable in body index: bodies inside a cell end up in increasing index */
  int *fill = (int *)malloc(sizeof(int) * (size_t)g->n_cells);
  memcpy(fill, w->cell_start, sizeof(int) * (size_t)g->n_cells);
  for (int i = 0; i < nd; i++) w->order[fill[w->cell_of[i]]++] = i;
  free(fill);
  return true;
}

static int grid_candidates(World *w, const Grid *g, int i, int **buf, int *cap) {
  int cell = w->cell_of[i];
  int ix = cell % g->dim[0], iy = (cell / g->dim[0]) % g->dim[1], iz = cell / (g->dim[0] * g->dim[1]);
  int n = 0;
  for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
    int x = ix + dx, y = iy + dy, z = iz + dz;
    if (x < 0 || y < 0 || z < 0 || x >= g->dim[0] || y >= g->dim[1] || z >= g->dim[2]) continue;
    int c = (z * g->dim[1] + y) * g->dim[0] + x;
    for (int k = w->cell_start[c]; k < w->cell_start[c + 1]; k++) {
      if (n == *cap) { *cap = *cap ? *cap * 2 : 256; *buf = (int *)realloc(*buf, sizeof(int) * (size_t)*cap); }
      (*buf)[n++] = w->order[k];
    }
  }
  qsort(*buf, (size_t)n, sizeof(int), cmp_int);
  return n;
}

static bool all_dynamic_spheres(const World *w) {
  for (int i = 0; i < w->n[1]; i++) if (w->cls[1][i].collider_type != ORA_SPHERE) return false;
  return true;
}

static void grid_bounds(const World *w, float dt, float *S, float *Dm, float *R) {
  float vmax = 0.0f, rmax = 0.0f;
  for (int i = 0; i < w->n[1]; i++) {
    const OraBody *b = &w->cls[1][i];
    float v = norm3(b->velocity), r = fabsf(b->shape[3] * b->scale[0]);
    if (v > vmax) vmax = v;
    if (r > rmax) rmax = r;
  }
  *S = 2.0f * vmax + 2.0f;
  *Dm = 2.0f * (*S) * dt + 2.0f * rmax + 0.01f;
  *R = (2.0f * rmax + 2.0f * (*S) * dt + 2.0f * (*Dm)) * 1.001f + 1e-4f;
}

static bool within(const World *w, int k, float S, float Dm) {
  const OraBody *b = &w->cls[1][k];
  const float *c0 = &w->p0[6 * (size_t)k];
  if (!(norm3(b->velocity) <= S)) return false;
  v3 c; add3(b->shape, b->position, c);
  v3 d; sub3(c, c0, d);
  return norm3(d) <= Dm;
}

static bool step_grid_once(World *w, float dt_in) {
  float dt = clamp_dt(dt_in);
  OraBody *S_ = w->cls[0], *D = w->cls[1];
  int ns = w->n[0], nd = w->n[1];
  if (w->n[2] != 0 || !all_dynamic_spheres(w) || nd < 64) return false;
  float S, Dm, R; grid_bounds(w, dt, &S, &Dm, &R);
  Grid g;
  if (!grid_build(w, R, &g)) return false;
  OraBody *backup = (OraBody *)malloc(sizeof(OraBody) * (size_t)nd);
  memcpy(backup, D, sizeof(OraBody) * (size_t)nd);
  int ev0 = w->n_events;
  bool ok = true;
  for (int i = 0; i < nd && ok; i++) {
    for (int j = 0; j < ns; j++) visit(w, &D[i], 1, i, &S_[j], 0, j, dt);
    ok = within(w, i, S, Dm);
  }
  int *cand = NULL, cap = 0;
  for (int i = 0; i < nd && ok; i++) {
    int n = grid_candidates(w, &g, i, &cand, &cap);
    for (int c = 0; c < n && ok; c++) {
      int j = cand[c];
      if (j == i) continue;
      if (visit(w, &D[i], 1, i, &D[j], 1, j, dt)) ok = within(w, i, S, Dm) && within(w, j, S, Dm);
    }
    integrate(&D[i], dt);
    if (ok) ok = within(w, i, S, Dm);
  }
  free(cand);
  if (!ok) { memcpy(D, backup, sizeof(OraBody) * (size_t)nd); w->n_events = ev0; }
  free(backup);
  return ok;
}

void ora_world_step_grid(void *h, float dt, int nsteps, int refresh_nodes) {
  World *w = (World *)h;
  for (int s = 0; s < nsteps; s++) {
    if (!step_grid_once(w, dt)) step_all_pairs(w, dt);
    if (refresh_nodes) ora_world_refresh_nodes(h);
    w->step_index++;
  }
}

/* Broadphase pair set on the frozen current state (SURVEY.md "fact 1": the set of unordered
 * dynamic pairs for which interval_collision is true). */
long long ora_world_pairs(void *h, float dt_in, int use_grid, int32_t *out, long long cap) {
  World *w = (World *)h;
  float dt = clamp_dt(dt_in), hit;
  OraBody *D = w->cls[1];
  int nd = w->n[1];
  long long n = 0;
  Grid g; float S, Dm, R;
  bool grid = false;
  if (use_grid && all_dynamic_spheres(w)) {
    /* frozen state: the predicate's own reach, 2 rmax + 2 vmax dt, is all the margin needed */
    grid_bounds(w, dt, &S, &Dm, &R);
    float vmax = (S - 2.0f) * 0.5f, rmax = (Dm - 2.0f * S * dt - 0.01f) * 0.5f;
    R = (2.0f * rmax + 2.0f * vmax * dt) * 1.001f + 1e-4f;
    grid = grid_build(w, R, &g);
  }
  int *cand = NULL, ccap = 0;
  for (int i = 0; i < nd; i++) {
    if (grid) {
      int nc = grid_candidates(w, &g, i, &cand, &ccap);
      for (int c = 0; c < nc; c++) {
        int j = cand[c];
        if (j <= i) continue;
        OraBody *A = &D[i], *B = &D[j];
        if (A->collider_type > B->collider_type) { OraBody *t = A; A = B; B = t; }
        if (interval_hit(A, B, 0, dt, &hit)) { if (n < cap) { out[2 * n] = i; out[2 * n + 1] = j; } n++; }
      }
    } else {
      for (int j = i + 1; j < nd; j++) {
        OraBody *A = &D[i], *B = &D[j];
        if (A->collider_type > B->collider_type) { OraBody *t = A; A = B; B = t; }
        if (interval_hit(A, B, 0, dt, &hit)) { if (n < cap) { out[2 * n] = i; out[2 * n + 1] = j; } n++; }
      }
    }
  }
  free(cand);
  return n;
}

/* ------------------------------------------------------------------ per-pair KAT entry points */

float ora_max_move(const OraBody *a, float t0, float t1) { return max_move(a, t0, t1); }
float ora_min_dist(const OraBody *a, const OraBody *b, float t) { return min_dist(a, b, t); }
int ora_interval_collision(const OraBody *a, const OraBody *b, float t0, float t1, float *hit) {
  float h = -1.0f;
  int r = interval_hit(a, b, t0, t1, &h);
  if (hit) *hit = h;
  return r;
}
int ora_narrow_phase(const OraBody *a, const OraBody *b, float dt, float *hit_time) {
  bool has;
  Contact c = narrow(a, b, dt, &has);
  if (!has) { *hit_time = 0.0f; return -1; }
  *hit_time = c.t;
  return c.hit ? 1 : 0;
}
int ora_resolve(OraBody *a, OraBody *b, float hit_time, float dt) {
  (void)dt;
  return resolve(a, b, hit_time) ? 0 : -1;
}
int ora_visit(OraBody *a, OraBody *b, float dt) { return visit(NULL, a, 0, 0, b, 0, 0, dt) ? 1 : 0; }
