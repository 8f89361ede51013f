/*
 * crux_physics.h -- the reference's physics-world C API, as the drop-in host layer exports it.
 *
 * Inside a Crux checkout this file is NOT used: crux_world.c is compiled with
 * -DCRUX_USE_ENGINE_HEADERS and includes the engine's own headers (include/physics/world.h,
 * entity.h, scene.h, event.h), so layouts come from the engine (and from the cglm it links).
 * Standalone (this repo's build and tests) the declarations below stand in for those headers.
 * They are layout-compatible with gcc/x86-64 and cglm's default 16-byte mat4 alignment:
 *   sizeof(PhysicsBody)=120, Collider=32, CollisionResult=24, Entity=104, GameEvent=64
 * (SURVEY.md section 8 header), checked by the static assertions at the bottom.
 *
 * Interface citations: /root/reference/include/physics/world.h:8-50, collider.h:12-39, aabb.h:14-31,
 * sphere.h:5-12, capsule.h:3-7, plane.h:5-8, distance.h:13-28, narrow_phase.h:9-36,
 * resolution.h:8-20, entity.h:11-25, types.h:3-9, event.h:8-52, scene.h:29-41, item.h:3-6.
 */
#ifndef CRUX_PHYSICS_COMPAT_H
#define CRUX_PHYSICS_COMPAT_H

#ifdef CRUX_USE_ENGINE_HEADERS

#include "entity.h"
#include "item.h"
#include "scene.h"
#include "event.h"
#include "physics/world.h"
#include "narrow_phase.h"
#include "distance.h"
#include "resolution.h"
#include "physics/utils.h"

#else /* standalone declarations */

#include <stdbool.h>
#include <stdint.h>
#include <time.h>

typedef float vec3[3];
typedef __attribute__((aligned(16))) float vec4[4];
typedef vec3 mat3[3];
typedef __attribute__((aligned(16))) vec4 mat4[4];
typedef unsigned char uuid_t[16];
typedef unsigned int GLuint;

typedef enum { ENTITY_GROUPING = 0, ENTITY_WORLD, ENTITY_ITEM, ENTITY_PLAYER, ENTITY_TYPE_COUNT } EntityType;
typedef enum { COLLISION_BEHAVIOR_PHYSICS = 0, COLLISION_BEHAVIOR_TRIGGER } CollisionBehavior;
typedef enum { COLLIDER_AABB = 0, COLLIDER_SPHERE, COLLIDER_CAPSULE, COLLIDER_PLANE, COLLIDER_COUNT } ColliderType;
typedef enum { EVENT_COLLISION = 0, EVENT_PLAYER_ITEM_PICKUP } EventType;

struct AABB { vec3 center; vec3 extents; bool initialized; };
struct Sphere { vec3 center; float radius; };
struct Capsule { vec3 segment_A; vec3 segment_B; float radius; };
struct Plane { vec3 normal; float distance; };
union ColliderData { struct AABB aabb; struct Sphere sphere; struct Capsule capsule; struct Plane plane; };
struct Collider { ColliderType type; union ColliderData data; };

struct Entity;
struct SceneNode;
struct ItemComponent { int id; int count; };

struct PhysicsBody {
  struct Collider collider;
  vec3 position, rotation, scale, velocity;
  float restitution;
  bool dynamic, at_rest;
  struct Entity *entity;
  struct SceneNode *scene_node;
  GLuint VAO, VBO, EBO;
};

struct PhysicsWorld {
  struct PhysicsBody *static_bodies, *dynamic_bodies, *player_bodies;
  unsigned int num_static_bodies, num_dynamic_bodies, num_player_bodies;
};

struct Entity {
  uuid_t id;
  EntityType type;
  vec3 position, rotation, scale, velocity;
  struct PhysicsBody *physics_body;
  struct Model *model;
  void *shader;
  struct ItemComponent *item;
};

struct SceneNode {
  uuid_t entity_id;
  unsigned int ID;
  mat4 local_transform, world_transform;
  vec3 position, rotation, scale;
  struct Entity *entity;
  struct SceneNode *parent_node;
  struct SceneNode **children;
  unsigned int num_children;
};

struct GameEvent {
  EventType type;
  struct timespec timestamp;
  union {
    struct { uuid_t entity_A_id, entity_B_id; } collision;
    struct { uuid_t player_entity_id; int item_id; int item_count; uuid_t item_entity_id; } item_pickup;
  } data;
};

struct CollisionResult { float hit_time, penetration; vec3 point_of_contact; bool colliding; };

#define NUM_COLLIDER_TYPES 4
typedef float (*DistanceFunction)(struct PhysicsBody *, struct PhysicsBody *, float);
typedef struct CollisionResult (*NarrowPhaseFunction)(struct PhysicsBody *, struct PhysicsBody *, float);
typedef void (*ResolutionFunction)(struct PhysicsBody *, struct PhysicsBody *, struct CollisionResult, float);
extern DistanceFunction distance_functions[NUM_COLLIDER_TYPES][NUM_COLLIDER_TYPES];
extern NarrowPhaseFunction narrow_phase_functions[NUM_COLLIDER_TYPES][NUM_COLLIDER_TYPES];
extern ResolutionFunction resolution_functions[NUM_COLLIDER_TYPES][NUM_COLLIDER_TYPES];

/* world.h:39-50 */
struct PhysicsWorld *physics_world_create(void);
struct PhysicsBody *physics_add_body(struct PhysicsWorld *, struct SceneNode *, struct Entity *, struct Collider, float restitution, bool dynamic);
struct PhysicsBody *physics_add_player(struct PhysicsWorld *, struct SceneNode *, struct Entity *, struct Collider);
void physics_remove_body(struct PhysicsWorld *, struct PhysicsBody *);
void physics_step(struct PhysicsWorld *, float delta_time);
void physics_sync_entities(struct PhysicsWorld *);
bool interval_collision(struct PhysicsBody *, struct PhysicsBody *, float start_time, float end_time, float *hit_time);
float maximum_object_movement_over_time(struct PhysicsBody *, float start_time, float end_time);
float minimum_object_distance_at_time(struct PhysicsBody *, struct PhysicsBody *, float time);

/* distance.h:17-28, narrow_phase.h:20-36, resolution.h:10-20 */
#define CRUX_PAIR_LIST(X) X(AABB_AABB) X(AABB_sphere) X(AABB_capsule) X(AABB_plane) X(sphere_sphere) \
  X(sphere_capsule) X(sphere_plane) X(capsule_capsule) X(capsule_plane)
#define CRUX_DECL_DIST(n) float min_dist_at_time_##n(struct PhysicsBody *, struct PhysicsBody *, float);
#define CRUX_DECL_NP(n) struct CollisionResult narrow_phase_##n(struct PhysicsBody *, struct PhysicsBody *, float);
#define CRUX_DECL_RS(n) void resolve_collision_##n(struct PhysicsBody *, struct PhysicsBody *, struct CollisionResult, float);
CRUX_PAIR_LIST(CRUX_DECL_DIST)
CRUX_PAIR_LIST(CRUX_DECL_NP)
CRUX_PAIR_LIST(CRUX_DECL_RS)
bool ray_intersect_AABB(vec3 p, vec3 d, struct AABB *e, float *hit_time, float end_time);

/* aabb.h:25-31, sphere.h:11, collider.h:39, physics/utils.h:8-12 */
void AABB_merge(struct AABB *a, struct AABB *b);
void AABB_update(struct AABB *src, mat3 rotation, vec3 translation, vec3 scale, struct AABB *dest);
void AABB_update_by_vertex(struct AABB *aabb, vec3 vertex);
bool AABB_intersect_AABB(struct AABB *a, struct AABB *b);
bool AABB_intersect_plane(struct AABB *box, struct Plane *plane);
bool sphere_intersect_sphere(struct Sphere *a, struct Sphere *b);
CollisionBehavior get_collision_behavior(EntityType a, EntityType b);
float distance_point_to_segment(vec3 point, vec3 segment_A, vec3 segment_B);
void print_aabb(struct AABB *aabb);
void print_plane(struct Plane *plane);

/* provided by the engine (src/event.c); the standalone tests link recording stubs */
void game_event_queue_enqueue(struct GameEvent game_event);
EventType get_event_type(EntityType type_A, EntityType type_B);

_Static_assert(sizeof(struct PhysicsBody) == 120, "PhysicsBody layout");
_Static_assert(sizeof(struct Collider) == 32, "Collider layout");
_Static_assert(sizeof(struct CollisionResult) == 24, "CollisionResult layout");
_Static_assert(sizeof(struct Entity) == 104, "Entity layout");
_Static_assert(sizeof(struct GameEvent) == 64, "GameEvent layout");
_Static_assert(sizeof(struct AABB) == 28, "AABB layout");

#endif /* CRUX_USE_ENGINE_HEADERS */

/* ---- additions of the B200 drop-in (not in the reference) ---- */
/* Release the device side of a world (the reference has no destroy; scene.c:569-571 frees the
 * three arrays with free(), which stays valid: they are plain calloc memory). */
void physics_world_release_gpu(struct PhysicsWorld *world);
/* Step `n_steps` frames on the device without the per-frame host round trip; scene-node matrices
 * are refreshed on the device after every step like scene_update does (scene.c:401-408). */
void physics_step_n(struct PhysicsWorld *world, float delta_time, int n_steps);
/* How many body records physics_step has sent to the device so far.  Only records that differ from the
 * device's copy travel (dirty ranges): a frame in which only the player was poked sends one. */
unsigned long long physics_world_uploaded_records(struct PhysicsWorld *world);

/* Debug-renderer interop (src/physics/debug_renderer.c): one PgDebugDraw record (include/physics_gpu.h) per body in draw
 * order -- AABB corner vertices, sphere / capsule model matrices -- written by a kernel into host memory or straight
 * into a device buffer the GL host shares with CUDA.  Returns the number of records, < 0 on error. */
int physics_debug_geometry(struct PhysicsWorld *world, void *dst, int dst_is_device, int max_records);

#endif /* CRUX_PHYSICS_COMPAT_H */
