/*
 * physics_gpu.h -- thin C-ABI between Crux's C host code and the B200 physics step.
 *
 * This is the boundary BASELINE.json's north_star names ("Host code stays in C and calls CUDA
 * through a thin C-ABI layer (physics_gpu.h/.cu)").  Plain C types only: pointers, sizes, POD
 * records.  No CUDA or torch type crosses it; device buffers an outside caller owns are passed as
 * `void *` device addresses.  The library is libphysics_gpu.so (crux_b200/csrc/physics_gpu.cu and
 * the kernel translation units next to it), sm_100a only, no CPU fallback: every entry point that
 * needs the GPU returns PG_ERR_CUDA when there is none.
 *
 * What each group replaces in the reference (paths relative to /root/reference):
 *   pg_worlds_*        struct PhysicsWorld + physics_world_create / physics_add_body /
 *                      physics_add_player / physics_remove_body     include/physics/world.h:28-43,
 *                                                                   src/physics/world.c:23-172
 *   pg_worlds_step     physics_step                                 world.h:44, world.c:174-555
 *                      (+ the per-frame scene_node_update refresh   src/scene.c:1051-1079)
 *   pg_worlds_events   the game_event_queue_enqueue calls made from the step
 *                                                                   world.c:232-259, 329-352, 441-457, 524-540
 *   pg_pair_*          interval_collision, minimum_object_distance_at_time,
 *                      maximum_object_movement_over_time and the three function tables
 *                                                                   world.h:47-50, distance.h:13-28,
 *                                                                   narrow_phase.h:16-33, resolution.h:8-20
 *   pg_large_*         the same physics_step for ONE world too large for a thread block
 *                      (uniform-grid broadphase; new acceleration structure)
 *
 * The reference-named C API (physics_world_create, physics_step, ...) lives one layer up, in
 * crux_b200/csrc/host/crux_world.c, and is written against this header.
 */
#ifndef PHYSICS_GPU_H
#define PHYSICS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums mirror include/physics/collider.h:18-24 and include/types.h:3-9 ---- */
enum { PG_AABB = 0, PG_SPHERE = 1, PG_CAPSULE = 2, PG_PLANE = 3 };
enum { PG_ENT_GROUPING = 0, PG_ENT_WORLD = 1, PG_ENT_ITEM = 2, PG_ENT_PLAYER = 3 };
enum { PG_CLASS_STATIC = 0, PG_CLASS_DYNAMIC = 1, PG_CLASS_PLAYER = 2 };

enum {
  PG_OK = 0,
  PG_ERR_CUDA = -1,       /* a CUDA runtime call failed (pg_last_error() has the text) */
  PG_ERR_ARG = -2,        /* bad handle / index / count */
  PG_ERR_CAPACITY = -3,   /* more bodies than the handle was created for */
  PG_ERR_UNSUPPORTED = -4,/* collider pair the reference leaves undefined (sphere-capsule,
                             capsule-capsule: distance.c:301-303,323-325; narrow_phase.c:522-524,590-592),
                             or a body mix an entry point does not serve */
  PG_ERR_UNPROVEN = -5    /* pg_large_step: the step ran, but its result is NOT proven identical to the
                             reference's sequential sweep (see pg_large_exactness); the state was advanced */
};

/*
 * Flat, pointer-free body record: `struct PhysicsBody` (world.h:8-26) field for field, plus the
 * matrices the physics path reads through body->scene_node (distance.c:26-49, 345-347) and two
 * rotation matrices the host computes with libm so that the device never evaluates sinf/cosf:
 *   euler_raw  = upper 3x3 of glm_euler_xyz(rotation)            -- rotation (DEGREES) passed as
 *                radians, exactly as distance.c:255,373 / narrow_phase.c:352,629 / resolution.c:431,667 do
 *   euler_node = upper 3x3 of glm_euler_xyz(glm_rad(rotation))   -- scene.c:1062-1068
 * Both are column-major (cglm mat3): m[col*3 + row].  pg_body_prepare() fills them.
 */
typedef struct PgBody {
  float position[3];
  float rotation[3];      /* degrees */
  float scale[3];
  float velocity[3];
  float restitution;
  float shape[7];         /* AABB: center,extents | sphere: center,radius | capsule: A,B,radius | plane: normal,distance */
  int32_t collider_type;
  int32_t entity_type;
  int32_t dynamic;
  int32_t at_rest;
  int32_t has_node;       /* scene_node != NULL */
  int32_t has_parent;     /* scene_node->parent_node != NULL */
  float node_xform[16];   /* scene_node->world_transform, column-major */
  float parent_xform[16]; /* scene_node->parent_node->world_transform */
  float euler_raw[9];
  float euler_node[9];
} PgBody;                 /* 304 bytes */

/* One contact = one game_event_queue_enqueue call of the reference, in solve order. */
typedef struct PgEvent {
  int32_t world;
  int32_t step;           /* 0-based within the pg_worlds_step call that produced it */
  int32_t event_type;     /* 0 = EVENT_COLLISION, 1 = EVENT_PLAYER_ITEM_PICKUP (event.h:8-11) */
  int32_t a_class, a_index;   /* body_A after the type-order swap (world.c:482-488) */
  int32_t b_class, b_index;
  int32_t pad_;           /* pass of physics_step that produced it (1..4, world.c:180,277,369,473) */
} PgEvent;                /* 32 bytes */

/* Per-step aggregate statistics over all dynamic bodies a handle owns (the vector a multi-GPU
 * run all-reduces; SURVEY.md section 5 "Metrics").  Plain doubles so that SUM is the reduction. */
typedef struct PgStats {
  double bodies;          /* number of dynamic bodies */
  double kinetic;         /* sum 0.5 |v|^2 (unit mass) */
  double momentum[3];     /* sum v */
  double contacts;        /* events emitted by the last pg_*_step call */
  double at_rest;         /* bodies with at_rest set */
  double nonfinite;       /* bodies with a NaN/Inf position or velocity component */
} PgStats;                /* 8 doubles */

const char *pg_last_error(void);
int  pg_device_count(void);
/* Fill euler_raw / euler_node of n records from their rotation field (host libm). */
void pg_body_prepare(PgBody *bodies, int n);
/* scene_node_update for one node (scene.c:1051-1079): world = parent * T(pos) R(rad(rot)) S(scale). */
void pg_node_transform(const float *pos3, const float *rot_deg3, const float *scale3,
                       const float *parent16_or_null, float *world16);

/* Page-locked host memory for the state-exchange calls (plain malloc'd buffers work too, slower). */
void *pg_host_alloc(size_t bytes);
void pg_host_free(void *p);

/* ------------------------------------------------------------------------------------------
 * Batched worlds, reference solve order preserved (one thread block per world).
 * All worlds of a handle share the same capacities; each keeps its own counts.
 * ------------------------------------------------------------------------------------------ */
typedef struct PgWorlds PgWorlds;

PgWorlds *pg_worlds_create(int n_worlds, int max_static, int max_dynamic, int max_player,
                           int max_events_per_world, int device);
void pg_worlds_destroy(PgWorlds *w);
int  pg_worlds_count(const PgWorlds *w);

/* Replace the bodies of one class of one world (n may be 0), records taken VERBATIM (the caller's
 * mirror is the truth, as it is for the reference's physics_step).  Order = reference array order. */
int  pg_worlds_set_bodies(PgWorlds *w, int world, int cls, const PgBody *bodies, int n);
/* Dirty-range upload: overwrite records [first, first+n) of one class of one world, verbatim; the
 * class keeps its length.  This is the per-frame player write path (player.c:11-65,160-173 pokes
 * position / rotation / velocity / at_rest of a handful of bodies between two steps). */
int  pg_worlds_update_bodies(PgWorlds *w, int world, int cls, int first, const PgBody *bodies, int n);
/* physics_add_body / physics_add_player: append one record applying the reference's add-time rules
 * (static velocity forced to 0, world.c:51; player restitution 0 and dynamic=false, world.c:137);
 * returns its index or <0. */
int  pg_worlds_add_body(PgWorlds *w, int world, const PgBody *body, int as_player);
/* physics_remove_body: swap-with-last and pop (world.c:147-172). */
int  pg_worlds_remove_body(PgWorlds *w, int world, int cls, int index);
int  pg_worlds_body_count(const PgWorlds *w, int world, int cls);
/* Copy one class of one world back (returns the count, <0 on error). */
int  pg_worlds_get_bodies(PgWorlds *w, int world, int cls, PgBody *out, int cap);

/* Homogeneous sphere worlds (the RL-style batch): every world [first, first+count) gets the same
 * static bodies and n_spheres dynamic spheres (local centre 0, scale 1) from SoA host arrays of
 * shape [count][n_spheres][3] (pos, vel) -- one call, one H2D copy per array. */
int  pg_worlds_set_spheres(PgWorlds *w, int first, int count, const PgBody *statics, int n_static,
                           int n_spheres, const float *pos, const float *vel, float radius,
                           float restitution);
/* SoA state exchange for the same worlds: pos/vel [count][n_dynamic][3], at_rest [count][n_dynamic]. */
int  pg_worlds_upload_state(PgWorlds *w, int first, int count, const float *pos, const float *vel,
                            const uint8_t *at_rest_or_null);
int  pg_worlds_download_state(PgWorlds *w, int first, int count, float *pos, float *vel,
                              uint8_t *at_rest_or_null);

/* n_steps x { physics_step(dt) ; if refresh_nodes: scene_node_update of every body's node }.
 * Asynchronous on the handle's stream; pg_worlds_sync() waits. */
int  pg_worlds_step(PgWorlds *w, float dt, int n_steps, int refresh_nodes);
int  pg_worlds_sync(PgWorlds *w);
/* Host-to-host step for sphere batches: upload pos/vel/at_rest ([n_worlds][n][3] / [n_worlds][n]),
 * run n_steps steps, download into the same arrays -- one call, internally pipelined over chunks of
 * worlds so that both PCIe directions and the kernels overlap.  Synchronous. */
int  pg_worlds_step_host(PgWorlds *w, float dt, int n_steps, float *pos, float *vel, uint8_t *at_rest_or_null);
/* The same with ONE host buffer in a packed layout, one record per world:
 *     float pos[n][3]; float vel[n][3]; uint8_t at_rest[n]; padding up to pg_worlds_packed_world_bytes()
 * A chunk of worlds is then one contiguous range: one DMA per chunk and direction instead of three. */
size_t pg_worlds_packed_world_bytes(const PgWorlds *w);
int  pg_worlds_step_host_packed(PgWorlds *w, float dt, int n_steps, void *state);
/* One game-loop frame of ONE world in a single round trip (the latency path of the drop-in's physics_step: a
 * game-sized scene is bound by the number of blocking CUDA calls, not by bytes or flops).  `counts` are the class
 * lengths of this frame; `ranges` name the records the host changed since the device last saw them (a class whose
 * length changed is passed whole) and `records` holds those records back to back, in range order, verbatim.  Then
 * n_steps steps as pg_worlds_step does them, and everything comes back: the records of all three classes (NULL to
 * skip a class) and the contact events in solve order.  Dirty records go up as asynchronous copies from a pinned
 * staging block, results come down into one, and the call synchronises ONCE. */
typedef struct PgRange { int32_t cls, first, count; } PgRange;
int  pg_worlds_frame(PgWorlds *w, int world, const int32_t counts[3], const PgRange *ranges, int n_ranges,
                     const PgBody *records, float dt, int n_steps, int refresh_nodes,
                     PgBody *out_static, PgBody *out_dynamic, PgBody *out_player,
                     PgEvent *events, long long events_cap, long long *n_events, int *overflowed);
/* Events of the last pg_worlds_step call, world-major then solve order.  Returns the total
 * number produced (may exceed cap; a world that overflowed max_events_per_world is reported
 * through *overflowed). */
long long pg_worlds_events(PgWorlds *w, PgEvent *out, long long cap, int *overflowed);
/* Aggregate statistics of THIS handle's bodies.  Sphere batches: the step launch itself leaves them behind
 * (its epilogue reduces them while the final state is still in shared memory), so this call only copies 8
 * doubles; for a state that was uploaded and not stepped it runs the same kernel with zero steps.  `contacts`
 * counts the events of the last step call (0 if the state was replaced since).  If dev_out8 != NULL the 8
 * doubles are ALSO copied to that device address on the handle's stream; host_out may be NULL to stay
 * asynchronous. */
int  pg_worlds_stats(PgWorlds *w, PgStats *host_out, void *dev_out8);

/* ---- multi-GPU: worlds shard across GPUs as independent units (no body data ever crosses GPUs); the only
 * exchange of the path is the SUM all-reduce of the PgStats vector, done HERE over NCCL (NVLink / NVSwitch).
 * One process (or thread) per GPU:  rank 0 calls pg_comm_unique_id and hands the 128 bytes to the other ranks by
 * any means (file, pipe, MPI, torch.distributed broadcast); every rank then calls pg_comm_init.  NCCL is bound
 * at run time (libnccl.so.2); without it these calls return PG_ERR_UNSUPPORTED / NULL and pg_last_error says so. */
typedef struct PgComm PgComm;
int  pg_comm_unique_id(void *out128);
PgComm *pg_comm_init(const void *unique_id128, int rank, int n_ranks, int device);
void pg_comm_destroy(PgComm *c);
int  pg_comm_rank(const PgComm *c);
int  pg_comm_size(const PgComm *c);
int  pg_comm_nccl_version(void);      /* e.g. 22809; 0 if NCCL is not available */
/* Statistics summed over ALL ranks of `comm` (NULL: this GPU alone).  Enqueued on the handle's stream behind the
 * step launch that produced the partials -- no host synchronisation unless host_out != NULL -- and the end mark
 * of the step's CUDA-event pair moves behind the collective: pg_worlds_last_step_ms then covers launch +
 * all-reduce.  Every rank of the communicator must make the call (it is a collective). */
int  pg_worlds_stats_allreduce(PgWorlds *w, PgComm *comm, PgStats *host_out, void *dev_out8);
/* Number of kernel launches this handle has issued (bench.py's gpu_launches) and the device
 * time, in ms, of the last pg_worlds_step call measured with CUDA events on the handle's stream. */
long long pg_worlds_launches(const PgWorlds *w);
float pg_worlds_last_step_ms(PgWorlds *w);
void *pg_worlds_stream(PgWorlds *w);   /* cudaStream_t as void* */

/* ------------------------------------------------------------------------------------------
 * Debug-geometry interop: what physics_debug_render (src/physics/debug_renderer.c:11-213) reads from the
 * bodies every frame, produced ON THE DEVICE from the resident state, so that a GL host shares a buffer
 * with CUDA (cudaGraphicsGLRegisterBuffer -> cudaGraphicsResourceGetMappedPointer -> `dst`, dst_is_device = 1)
 * instead of reading every body back.  One record per body in the reference's draw order: players, statics,
 * dynamics.  GL itself (VAO/VBO/EBO creation, shaders, draw calls) stays with the host; no GL symbol is linked here.
 *   PG_DEBUG_AABB_LINES     data[0..23] = the 8 corner vertices physics_debug_AABB_render uploads with glBufferSubData
 *                           (debug_renderer.c:662-688): world AABB through the scene node (translation, column norms,
 *                           upper 3x3 / scale[0], AABB_update), corners in the order the 24 line indices (:321-337) expect
 *   PG_DEBUG_SPHERE_MODEL   data[0..15] = column-major `model` given to shader_set_mat4 (:700): translate(position
 *                           [+ sphere centre for a dynamic body, :184-193]) * scale(body scale)
 *   PG_DEBUG_CAPSULE_MODEL  data[0..15] = the scene node's world transform (:49-51, :206)
 *   PG_DEBUG_CAPSULE_HOST   capsule without a scene node (:53-59, :109-116): model = T R_y R_x R_z S needs sinf/cosf, which
 *                           never run on the device (parity contract); pg_debug_capsule_model() builds it on the host
 *   PG_DEBUG_NONE           planes (never drawn, :61, :119, :208) and AABBs without a scene node (a NULL dereference upstream)
 * ------------------------------------------------------------------------------------------ */
enum { PG_DEBUG_NONE = 0, PG_DEBUG_AABB_LINES = 1, PG_DEBUG_SPHERE_MODEL = 2, PG_DEBUG_CAPSULE_MODEL = 3, PG_DEBUG_CAPSULE_HOST = 4 };
typedef struct PgDebugDraw {
  int32_t kind;
  int32_t cls, index;     /* which body */
  int32_t n_indices;      /* element count of the draw call: 24 (GL_LINES), 2280 / 4920 (GL_TRIANGLES), 0 */
  float data[24];
} PgDebugDraw;            /* 112 bytes */
/* Fill one record per body of `world` (players, statics, dynamics).  dst_is_device: `dst` is a device address (a mapped
 * GL buffer) written by the kernel on the handle's stream -- asynchronous, pg_worlds_sync() or the caller's own
 * stream order makes it visible; otherwise `dst` is host memory and the call synchronises.  Returns the number of
 * records (<0: error; PG_ERR_CAPACITY if max_records is too small). */
int  pg_worlds_debug_geometry(PgWorlds *w, int world, void *dst, int dst_is_device, int max_records);
/* Host-side halves of the debug renderer (init-time meshes and the one model matrix the device does not build), libm like
 * the reference.  Sizes: sphere 441 vertices x 3 floats, 2280 indices (:371-432); capsule 882 vertices x 3 floats, 4920
 * indices (:469-616); box 8 vertices x 3 floats, 24 indices (:311-337). */
void pg_debug_sphere_mesh(float radius, float *vertices1323, uint32_t *indices2280);
void pg_debug_capsule_mesh(const float *seg_a3, const float *seg_b3, float radius, float *vertices2646, uint32_t *indices4920);
void pg_debug_box_mesh(const float *center3, const float *extents3, float *vertices24, uint32_t *indices24);
void pg_debug_capsule_model(const PgBody *body, float *model16);

/* ------------------------------------------------------------------------------------------
 * One large world of dynamic spheres + static bodies (uniform-grid broadphase).
 * ------------------------------------------------------------------------------------------ */
typedef struct PgLarge PgLarge;

/* Static bodies may be planes (at most 16) and AABBs with scene nodes (level geometry, any number
 * up to max_static); their array order is the reference's visit order.  When AABBs are present the
 * spheres are visited THROUGH SCENE NODES like the reference does (distance.c:113-125): node
 * translation = body position at the start of the step (refreshed by the per-frame
 * scene_node_update), identity rotation, unit scale.  Player bodies (capsules) run passes 1-2
 * against the level; sphere-capsule visits never fire (undefined upstream). */
PgLarge *pg_large_create(int max_spheres, int max_static, int device);
void pg_large_destroy(PgLarge *w);
int  pg_large_set(PgLarge *w, const PgBody *statics, int n_static, int n_spheres,
                  const float *pos, const float *vel, const float *radius_or_null, float radius,
                  float restitution);
int  pg_large_set_players(PgLarge *w, const PgBody *players, int n);
int  pg_large_get_players(PgLarge *w, PgBody *out, int cap);
int  pg_large_upload_state(PgLarge *w, const float *pos, const float *vel, const uint8_t *at_rest_or_null);
int  pg_large_download_state(PgLarge *w, float *pos, float *vel, uint8_t *at_rest_or_null);
/* PG_OK only if every step is PROVEN identical to the reference's sequential sweep.  PG_ERR_UNPROVEN: all n_steps ran, but
 * at least one fixed point did not converge (the state HAS been advanced).  Any other error: the step that failed left
 * the state as it found it (earlier steps of the same call stay). */
int  pg_large_step(PgLarge *w, float dt, int n_steps);
int  pg_large_sync(PgLarge *w);
/* Broadphase pair set on the current (frozen) state: unordered pairs i<j for which
 * interval_collision(i,j,0,dt) holds, sorted lexicographically.  Returns the count. */
long long pg_large_pairs(PgLarge *w, float dt, int32_t *out_ij, long long cap);
/* Contacts (events) of the last step call in solve order; same record as the batched path. */
long long pg_large_events(PgLarge *w, PgEvent *out, long long cap, int *overflowed);
int  pg_large_stats(PgLarge *w, PgStats *host_out, void *dev_out8);
long long pg_large_launches(const PgLarge *w);
float pg_large_last_step_ms(PgLarge *w);
/* Per-kernel device time of the last step (ms), in pipeline order; names via pg_large_kernel_name. */
int  pg_large_kernel_count(void);
const char *pg_large_kernel_name(int k);
float pg_large_kernel_ms(PgLarge *w, int k);
/* Per-kernel timing costs a CUDA event pair around every launch and is OFF by default (a small world's
 * pipeline is launch-bound); switch it on before the steps whose pg_large_kernel_ms you want. */
int  pg_large_set_timing(PgLarge *w, int on);
/* Fixed-point statistics of the last step: iterations used, 1 if the step is proven identical to
 * the sequential reference order, candidate-visit and hit counts. */
int  pg_large_exactness(PgLarge *w, int *iterations, int *proven_exact, long long *candidates, long long *hits);

/* ------------------------------------------------------------------------------------------
 * Per-pair entry points: the reference's public predicate/table functions evaluated on the GPU
 * for n independent pairs (records in host memory; results in host memory).  A[k] must already be
 * the lower collider type, as the tables require (world.c:482-488).  The table-level queries (min_dist,
 * interval_collision, narrow_phase, resolve) return PG_ERR_UNSUPPORTED for a sphere-capsule or
 * capsule-capsule pair -- empty functions with undefined results upstream; pg_pair_visit,
 * which is what the step does with such a pair, reports "no event" and leaves both bodies alone.
 * ------------------------------------------------------------------------------------------ */
int  pg_pair_max_move(const PgBody *a, int n, float t0, float t1, float *out);
int  pg_pair_min_dist(const PgBody *a, const PgBody *b, int n, const float *t, float *out);
int  pg_pair_interval_collision(const PgBody *a, const PgBody *b, int n, float t0, float t1,
                                int32_t *out_hit, float *out_hit_time);
int  pg_pair_narrow_phase(const PgBody *a, const PgBody *b, int n, float dt,
                          int32_t *out_colliding, float *out_hit_time);
int  pg_pair_resolve(PgBody *a, PgBody *b, int n, const float *hit_time, float dt);
/* one full visit (order swap, gate, narrowphase, behaviour, resolve); out_event[k] = 1 if an event fires */
int  pg_pair_visit(PgBody *a, PgBody *b, int n, float dt, int32_t *out_event);

#ifdef __cplusplus
}
#endif
#endif /* PHYSICS_GPU_H */
