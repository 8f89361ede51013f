/*
 * oracle/crux_oracle.h -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * API of the plain-C restatement of the Crux physics step (oracle/crux_oracle.c).  It is the
 * checker for the CUDA path; the product never links or loads it.  Same flat API as
 * oracle/ref_harness.c (prefix ora_ instead of ref_) so one test drives both.
 *
 * Pinned against: (1) the reference's own Unity vectors, test/physics/test_aabb.c:24-141
 * (tests/test_oracle_golden.py), (2) the unmodified reference compiled here (oracle/_ref), bit for
 * bit, on per-pair KATs, contact-event sequences and 1000-step trajectories
 * (tests/test_oracle_vs_ref.py; fixtures in tests/golden/ for boxes without /root/reference).
 */
#ifndef CRUX_ORACLE_H
#define CRUX_ORACLE_H

#include "oracle_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

void *ora_world_create(void);
void  ora_world_destroy(void *w);
int   ora_world_add_many(void *w, const OraBody *bodies, int n, int as_player);
int   ora_world_remove(void *w, int cls, int index);
void  ora_world_counts(void *w, int *counts3);
void  ora_world_read(void *w, int cls, OraBody *out);
void  ora_world_poke(void *w, int cls, int index, const OraBody *rec);
void  ora_world_refresh_nodes(void *w);
void  ora_world_step(void *w, float dt, int nsteps, int refresh_nodes);
/* same result as ora_world_step (falls back to it when its bounds are exceeded) for worlds whose
 * dynamic bodies are all spheres; O(N) per step through a uniform grid */
void  ora_world_step_grid(void *w, float dt, int nsteps, int refresh_nodes);
/* unordered dynamic-dynamic pairs (i<j) for which interval_collision(i,j,0,dt) holds on the
 * current (frozen) state; use_grid=0 is the all-pairs definition.  Returns the count, writes up
 * to `cap` (i,j) int32 pairs sorted lexicographically. */
long long ora_world_pairs(void *w, float dt, int use_grid, int32_t *out_ij, long long cap);
void  ora_world_record_events(void *w, int on);
int   ora_events_count(void *w);
void  ora_events_read(void *w, OraEvent *out);
void  ora_events_clear(void *w);
void  ora_set_threads(int n);

float ora_max_move(const OraBody *a, float t0, float t1);
float ora_min_dist(const OraBody *a, const OraBody *b, float t);
int   ora_interval_collision(const OraBody *a, const OraBody *b, float t0, float t1, float *hit);
int   ora_narrow_phase(const OraBody *a, const OraBody *b, float dt, float *hit_time);
int   ora_resolve(OraBody *a, OraBody *b, float hit_time, float dt);
int   ora_visit(OraBody *a, OraBody *b, float dt);

/* collider helpers outside physics_step */
void  ora_aabb_update(const float *src_center_extents6, const float *rot9, const float *translation3,
                      const float *scale3, float *dst_center_extents6);
int   ora_aabb_intersect_aabb(const float *a6, const float *b6);
int   ora_aabb_intersect_plane(const float *box6, const float *plane4);
void  ora_aabb_merge(float *a6, int *a_init, const float *b6, int b_init);
void  ora_aabb_update_by_vertex(float *a6, const float *v3);
int   ora_sphere_intersect_sphere(const float *a4, const float *b4);
float ora_distance_point_to_segment(const float *p3, const float *a3, const float *b3);
int   ora_collision_behavior(int entity_type_a, int entity_type_b);
int   ora_event_type(int entity_type_a, int entity_type_b);
void  ora_euler_xyz(const float *angles3, float *rot9);
void  ora_node_transform(const float *pos3, const float *rot_deg3, const float *scale3,
                         const float *parent16_or_null, float *world16);

#ifdef __cplusplus
}
#endif
#endif /* CRUX_ORACLE_H */
