/*
 * crux_colliders.c -- the rest of the reference's public physics symbols for the drop-in layer:
 *   - the three [ColliderType][ColliderType] function tables and their 27 entry functions
 *     (distance.c:7-17, narrow_phase.c:8-19, resolution.c:8-18).  Each entry evaluates ONE pair on
 *     the GPU through pg_pair_* (the device dispatches on the collider types in the records, so
 *     all entries share three small wrappers).  They exist for API completeness and tests; the
 *     step never calls them one pair at a time.
 *   - the collider helpers that are plain host code upstream and are not on the step path
 *     (aabb.c, sphere.c, collider.c, physics/utils.c, ray_intersect_AABB): restated here in C.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "crux_physics.h"
#include "physics_gpu.h"

void crux_flatten_body(const struct PhysicsBody *b, PgBody *r);   /* crux_world.c */

/* ------------------------------------------------------------------ single-pair GPU wrappers */

static float pair_distance(struct PhysicsBody *A, struct PhysicsBody *B, float time) {
  PgBody a, b;
  crux_flatten_body(A, &a);
  crux_flatten_body(B, &b);
  float out = 0.0f;
  if (pg_pair_min_dist(&a, &b, 1, &time, &out) != PG_OK) fprintf(stderr, "Error: min_dist_at_time: %s\n", pg_last_error());
  return out;
}

static struct CollisionResult pair_narrow(struct PhysicsBody *A, struct PhysicsBody *B, float delta_time) {
  struct CollisionResult r;
  memset(&r, 0, sizeof(r));
  PgBody a, b;
  crux_flatten_body(A, &a);
  crux_flatten_body(B, &b);
  int32_t colliding = 0;
  float t = 0.0f;
  if (pg_pair_narrow_phase(&a, &b, 1, delta_time, &colliding, &t) != PG_OK) {
    fprintf(stderr, "Error: narrow_phase: %s\n", pg_last_error());
    return r;
  }
  r.hit_time = t;
  r.colliding = colliding > 0;
  return r;
}

static void pair_resolve(struct PhysicsBody *A, struct PhysicsBody *B, struct CollisionResult result, float delta_time) {
  PgBody a, b;
  crux_flatten_body(A, &a);
  crux_flatten_body(B, &b);
  if (pg_pair_resolve(&a, &b, 1, &result.hit_time, delta_time) != PG_OK) {
    fprintf(stderr, "Error: resolve_collision: %s\n", pg_last_error());
    return;
  }
  memcpy(A->position, a.position, 12); memcpy(A->velocity, a.velocity, 12); A->at_rest = a.at_rest != 0;
  memcpy(B->position, b.position, 12); memcpy(B->velocity, b.velocity, 12); B->at_rest = b.at_rest != 0;
}

#define CRUX_DEF_DIST(n) float min_dist_at_time_##n(struct PhysicsBody *A, struct PhysicsBody *B, float t) { return pair_distance(A, B, t); }
#define CRUX_DEF_NP(n) struct CollisionResult narrow_phase_##n(struct PhysicsBody *A, struct PhysicsBody *B, float dt) { return pair_narrow(A, B, dt); }
#define CRUX_DEF_RS(n) void resolve_collision_##n(struct PhysicsBody *A, struct PhysicsBody *B, struct CollisionResult r, float dt) { pair_resolve(A, B, r, dt); }
#define CRUX_LIVE_PAIRS(X) X(AABB_AABB) X(AABB_sphere) X(AABB_capsule) X(AABB_plane) X(sphere_sphere) X(sphere_plane) X(capsule_plane)
CRUX_LIVE_PAIRS(CRUX_DEF_DIST)
CRUX_LIVE_PAIRS(CRUX_DEF_NP)
CRUX_LIVE_PAIRS(CRUX_DEF_RS)

/* The two pairs the reference leaves as empty functions (distance.c:301-303,323-325;
 * narrow_phase.c:522-524,590-592; resolution.c:572-574,624-626).  Upstream their return values are
 * undefined; here they are defined as "distance 0, no contact, no-op", which is also what the
 * device does when the step meets such a pair. */
float min_dist_at_time_sphere_capsule(struct PhysicsBody *A, struct PhysicsBody *B, float t) { (void)A; (void)B; (void)t; return 0.0f; }
float min_dist_at_time_capsule_capsule(struct PhysicsBody *A, struct PhysicsBody *B, float t) { (void)A; (void)B; (void)t; return 0.0f; }
struct CollisionResult narrow_phase_sphere_capsule(struct PhysicsBody *A, struct PhysicsBody *B, float dt) {
  (void)A; (void)B; (void)dt; struct CollisionResult r; memset(&r, 0, sizeof(r)); return r;
}
struct CollisionResult narrow_phase_capsule_capsule(struct PhysicsBody *A, struct PhysicsBody *B, float dt) {
  (void)A; (void)B; (void)dt; struct CollisionResult r; memset(&r, 0, sizeof(r)); return r;
}
void resolve_collision_sphere_capsule(struct PhysicsBody *A, struct PhysicsBody *B, struct CollisionResult r, float dt) { (void)A; (void)B; (void)r; (void)dt; }
void resolve_collision_capsule_capsule(struct PhysicsBody *A, struct PhysicsBody *B, struct CollisionResult r, float dt) { (void)A; (void)B; (void)r; (void)dt; }

DistanceFunction distance_functions[NUM_COLLIDER_TYPES][NUM_COLLIDER_TYPES] = {
  [COLLIDER_AABB][COLLIDER_AABB] = min_dist_at_time_AABB_AABB,
  [COLLIDER_AABB][COLLIDER_SPHERE] = min_dist_at_time_AABB_sphere,
  [COLLIDER_AABB][COLLIDER_CAPSULE] = min_dist_at_time_AABB_capsule,
  [COLLIDER_AABB][COLLIDER_PLANE] = min_dist_at_time_AABB_plane,
  [COLLIDER_SPHERE][COLLIDER_SPHERE] = min_dist_at_time_sphere_sphere,
  [COLLIDER_SPHERE][COLLIDER_CAPSULE] = min_dist_at_time_sphere_capsule,
  [COLLIDER_SPHERE][COLLIDER_PLANE] = min_dist_at_time_sphere_plane,
  [COLLIDER_CAPSULE][COLLIDER_CAPSULE] = min_dist_at_time_capsule_capsule,
  [COLLIDER_CAPSULE][COLLIDER_PLANE] = min_dist_at_time_capsule_plane,
};
NarrowPhaseFunction narrow_phase_functions[NUM_COLLIDER_TYPES][NUM_COLLIDER_TYPES] = {
  [COLLIDER_AABB][COLLIDER_AABB] = narrow_phase_AABB_AABB,
  [COLLIDER_AABB][COLLIDER_SPHERE] = narrow_phase_AABB_sphere,
  [COLLIDER_AABB][COLLIDER_CAPSULE] = narrow_phase_AABB_capsule,
  [COLLIDER_AABB][COLLIDER_PLANE] = narrow_phase_AABB_plane,
  [COLLIDER_SPHERE][COLLIDER_SPHERE] = narrow_phase_sphere_sphere,
  [COLLIDER_SPHERE][COLLIDER_CAPSULE] = narrow_phase_sphere_capsule,
  [COLLIDER_SPHERE][COLLIDER_PLANE] = narrow_phase_sphere_plane,
  [COLLIDER_CAPSULE][COLLIDER_CAPSULE] = narrow_phase_capsule_capsule,
  [COLLIDER_CAPSULE][COLLIDER_PLANE] = narrow_phase_capsule_plane,
};
ResolutionFunction resolution_functions[NUM_COLLIDER_TYPES][NUM_COLLIDER_TYPES] = {
  [COLLIDER_AABB][COLLIDER_AABB] = resolve_collision_AABB_AABB,
  [COLLIDER_AABB][COLLIDER_SPHERE] = resolve_collision_AABB_sphere,
  [COLLIDER_AABB][COLLIDER_CAPSULE] = resolve_collision_AABB_capsule,
  [COLLIDER_AABB][COLLIDER_PLANE] = resolve_collision_AABB_plane,
  [COLLIDER_SPHERE][COLLIDER_SPHERE] = resolve_collision_sphere_sphere,
  [COLLIDER_SPHERE][COLLIDER_CAPSULE] = resolve_collision_sphere_capsule,
  [COLLIDER_SPHERE][COLLIDER_PLANE] = resolve_collision_sphere_plane,
  [COLLIDER_CAPSULE][COLLIDER_CAPSULE] = resolve_collision_capsule_capsule,
  [COLLIDER_CAPSULE][COLLIDER_PLANE] = resolve_collision_capsule_plane,
};

/* ------------------------------------------------------------------ host-only helpers */

static inline float minf(float a, float b) { return a < b ? a : b; }
static inline float maxf(float a, float b) { return a > b ? a : b; }

/* aabb.c:8-51: union of two boxes in centre/extents form; an uninitialised `a` adopts `b` */
void AABB_merge(struct AABB *a, struct AABB *b) {
  if (!a->initialized) {
    memcpy(a->center, b->center, 12);
    memcpy(a->extents, b->extents, 12);
    a->initialized = true;
    return;
  }
  if (!b->initialized) return;
  for (int i = 0; i < 3; i++) {
    float lo = minf(a->center[i] - a->extents[i], b->center[i] - b->extents[i]);
    float hi = maxf(a->center[i] + a->extents[i], b->center[i] + b->extents[i]);
    a->center[i] = (lo + hi) * 0.5f;
    a->extents[i] = (hi - lo) * 0.5f;
  }
  a->initialized = true;
}

/* aabb.c:59-78: OBB -> AABB; the centre is rotated and translated but NOT scaled; the extent
 * products run in double because fabs() is the double function */
void AABB_update(struct AABB *src, mat3 rotation, vec3 translation, vec3 scale, struct AABB *dest) {
  for (int i = 0; i < 3; i++) {
    dest->center[i] = translation[i];
    dest->extents[i] = 0.0f;
    for (int j = 0; j < 3; j++) {
      dest->center[i] += rotation[j][i] * src->center[j];
      dest->extents[i] = (float)((double)dest->extents[i] + fabs((double)rotation[j][i]) * (double)src->extents[j] * fabs((double)scale[i]));
    }
  }
}

/* aabb.c:80-95 */
void AABB_update_by_vertex(struct AABB *aabb, vec3 vertex) {
  for (int i = 0; i < 3; i++) {
    float hi = aabb->center[i] + aabb->extents[i];
    if (vertex[i] > hi) {
      float grow = (vertex[i] - hi) * 0.5f;
      aabb->center[i] += grow;
      aabb->extents[i] += grow;
    }
    float lo = aabb->center[i] - aabb->extents[i];
    if (vertex[i] < lo) {
      float grow = (aabb->center[i] - aabb->extents[i] - vertex[i]) * 0.5f;
      aabb->center[i] -= grow;
      aabb->extents[i] += grow;
    }
  }
}

/* aabb.c:100-105: separated on any axis -> false; touching counts as intersecting */
bool AABB_intersect_AABB(struct AABB *a, struct AABB *b) {
  for (int i = 0; i < 3; i++)
    if (fabs((double)(a->center[i] - b->center[i])) > (double)(a->extents[i] + b->extents[i])) return false;
  return true;
}

/* aabb.c:108-121 */
bool AABB_intersect_plane(struct AABB *box, struct Plane *plane) {
  float r = (float)((double)box->extents[0] * fabs((double)plane->normal[0]) + (double)box->extents[1] * fabs((double)plane->normal[1]) +
                    (double)box->extents[2] * fabs((double)plane->normal[2]));
  float s = plane->normal[0] * box->center[0] + plane->normal[1] * box->center[1] + plane->normal[2] * box->center[2] - plane->distance;
  return fabs((double)s) <= (double)r;
}

/* sphere.c:5-12 */
bool sphere_intersect_sphere(struct Sphere *sa, struct Sphere *sb) {
  float dx = sa->center[0] - sb->center[0], dy = sa->center[1] - sb->center[1], dz = sa->center[2] - sb->center[2];
  float rs = sa->radius + sb->radius;
  return dx * dx + dy * dy + dz * dz <= rs * rs;
}

/* collider.c:7-17 */
CollisionBehavior get_collision_behavior(EntityType type_A, EntityType type_B) {
  if (type_A == ENTITY_GROUPING || type_B == ENTITY_GROUPING) return COLLISION_BEHAVIOR_TRIGGER;
  if ((type_A == ENTITY_ITEM && type_B == ENTITY_PLAYER) || (type_A == ENTITY_PLAYER && type_B == ENTITY_ITEM)) return COLLISION_BEHAVIOR_TRIGGER;
  return COLLISION_BEHAVIOR_PHYSICS;
}

/* physics/utils.c:33-60 */
float distance_point_to_segment(vec3 point, vec3 segment_A, vec3 segment_B) {
  float seg[3], ap[3];
  for (int i = 0; i < 3; i++) { seg[i] = segment_B[i] - segment_A[i]; ap[i] = point[i] - segment_A[i]; }
  float proj = seg[0] * ap[0] + seg[1] * ap[1] + seg[2] * ap[2];
  float d = proj / (seg[0] * seg[0] + seg[1] * seg[1] + seg[2] * seg[2]);
  float v[3];
  if (d <= 0) { memcpy(v, ap, 12); }
  else if (d >= 1) { for (int i = 0; i < 3; i++) v[i] = point[i] - segment_B[i]; }
  else { for (int i = 0; i < 3; i++) v[i] = (segment_A[i] + seg[i] * d) - point[i]; }
  return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

/* narrow_phase.c:702-739: slab test of the ray p + t d against a box over [0, end_time] */
bool ray_intersect_AABB(vec3 p, vec3 d, struct AABB *aabb, float *hit_time, float end_time) {
  *hit_time = 0.0f;
  float t_max = end_time;
  for (int i = 0; i < 3; i++) {
    float lo = aabb->center[i] - aabb->extents[i], hi = aabb->center[i] + aabb->extents[i];
    if ((double)fabsf(d[i]) < 0.0001) {
      if (p[i] < lo || p[i] > hi) return false;
    } else {
      float t1 = (lo - p[i]) / d[i], t2 = (hi - p[i]) / d[i];
      if (t1 > t2) { float tmp = t1; t1 = t2; t2 = tmp; }
      if (t1 > *hit_time) *hit_time = t1;
      if (t2 < t_max) t_max = t2;
      if (*hit_time > t_max) return false;
    }
  }
  return true;
}

void print_aabb(struct AABB *aabb) {
  printf("AABB {\n  center: (%.7f, %.7f, %.7f)\n  extents: (%.7f, %.7f, %.7f)\n}\n", aabb->center[0], aabb->center[1], aabb->center[2],
         aabb->extents[0], aabb->extents[1], aabb->extents[2]);
}
void print_plane(struct Plane *plane) {
  printf("Plane:\n{\n  Normal: (%.3f, %.3f, %.3f)\n  Distance: %.2f\n}\n", plane->normal[0], plane->normal[1], plane->normal[2], plane->distance);
}
