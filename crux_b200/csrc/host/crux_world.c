/*
 * crux_world.c -- drop-in replacement for /root/reference/src/physics/world.c on top of the C-ABI
 * of include/physics_gpu.h.  Same exported symbols, same argument meaning, same error style
 * (NULL + a line on stderr; physics_step is void and only logs), same ownership rules:
 *   - the world and its three body arrays are plain calloc memory (the scene frees them with
 *     free(), scene.c:569-571), 128 slots per class like world.c:15,32-34;
 *   - returned PhysicsBody pointers point INTO those arrays and stay valid; physics_remove_body
 *     swaps with the last element and re-points that entity (world.c:147-172);
 *   - callers (player.c:27-45,155-172; scene.c:1054-1058) read and write body fields directly
 *     between steps, so the host arrays are the truth: every physics_step flattens them into
 *     PgBody records, steps on the GPU, and writes position / velocity / at_rest back.
 * The device handle hangs off a side table keyed by the world pointer.
 * Contacts come back as an ordered list and are replayed into game_event_queue_enqueue exactly
 * as world.c:232-259, 329-352, 441-457, 524-540 build them.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "crux_physics.h"
#include "physics_gpu.h"

#define MAX_PHYSICS_BODIES 128          /* world.c:15 */
#define CRUX_MAX_EVENTS 4096
#define CRUX_MAX_RANGES 4               /* dirty ranges per class and frame before they are merged into one */

typedef struct Shadow {
  struct PhysicsWorld *world;
  PgWorlds *gpu;
  int cap[3];
  PgBody *rec[3];             /* scratch: this frame's flattened mirror / the downloaded records */
  PgBody *dev[3];             /* what the device holds (as of the last upload or download) */
  int dev_n[3];               /* ... and how many records of each class; -1 = unknown, upload everything */
  int rec_cap[3];
  unsigned long long uploaded_records;   /* statistics: body records sent to the device so far */
  PgBody *up;                 /* this frame's dirty records, back to back in range order */
  size_t up_cap;
  PgEvent *events;
  struct Shadow *next;
} Shadow;

static Shadow *g_shadows = NULL;

static Shadow *shadow_find(struct PhysicsWorld *w, int create) {
  for (Shadow *s = g_shadows; s; s = s->next) if (s->world == w) return s;
  if (!create) return NULL;
  Shadow *s = (Shadow *)calloc(1, sizeof(Shadow));
  if (!s) return NULL;
  s->world = w;
  s->next = g_shadows;
  g_shadows = s;
  return s;
}

/* ------------------------------------------------------------------ flatten / unflatten */

void crux_flatten_body(const struct PhysicsBody *b, PgBody *r) {
  memset(r, 0, sizeof(*r));
  memcpy(r->position, b->position, 12);
  memcpy(r->rotation, b->rotation, 12);
  memcpy(r->scale, b->scale, 12);
  memcpy(r->velocity, b->velocity, 12);
  r->restitution = b->restitution;
  r->collider_type = (int32_t)b->collider.type;
  switch (b->collider.type) {
    case COLLIDER_AABB:
      memcpy(r->shape, b->collider.data.aabb.center, 12);
      memcpy(r->shape + 3, b->collider.data.aabb.extents, 12);
      break;
    case COLLIDER_SPHERE:
      memcpy(r->shape, b->collider.data.sphere.center, 12);
      r->shape[3] = b->collider.data.sphere.radius;
      break;
    case COLLIDER_CAPSULE:
      memcpy(r->shape, b->collider.data.capsule.segment_A, 12);
      memcpy(r->shape + 3, b->collider.data.capsule.segment_B, 12);
      r->shape[6] = b->collider.data.capsule.radius;
      break;
    default:
      memcpy(r->shape, b->collider.data.plane.normal, 12);
      r->shape[3] = b->collider.data.plane.distance;
      break;
  }
  r->entity_type = b->entity ? (int32_t)b->entity->type : 0;
  r->dynamic = b->dynamic ? 1 : 0;
  r->at_rest = b->at_rest ? 1 : 0;
  if (b->scene_node) {
    r->has_node = 1;
    memcpy(r->node_xform, b->scene_node->world_transform, 64);
    if (b->scene_node->parent_node) {
      r->has_parent = 1;
      memcpy(r->parent_xform, b->scene_node->parent_node->world_transform, 64);
    }
  }
  pg_body_prepare(r, 1);
}

static void unflatten_state(const PgBody *r, struct PhysicsBody *b) {
  memcpy(b->position, r->position, 12);
  memcpy(b->velocity, r->velocity, 12);
  b->at_rest = r->at_rest != 0;
}

static struct PhysicsBody *class_array(struct PhysicsWorld *w, int cls, unsigned int *n) {
  switch (cls) {
    case PG_CLASS_STATIC: *n = w->num_static_bodies; return w->static_bodies;
    case PG_CLASS_DYNAMIC: *n = w->num_dynamic_bodies; return w->dynamic_bodies;
    default: *n = w->num_player_bodies; return w->player_bodies;
  }
}

/* ------------------------------------------------------------------ world management */

struct PhysicsWorld *physics_world_create(void) {
  struct PhysicsWorld *world = (struct PhysicsWorld *)calloc(1, sizeof(struct PhysicsWorld));
  if (!world) {
    printf("Error: failed to allocate PhysicsWorld in physics_world_create\n");
    return NULL;
  }
  world->static_bodies = (struct PhysicsBody *)calloc(MAX_PHYSICS_BODIES, sizeof(struct PhysicsBody));
  world->dynamic_bodies = (struct PhysicsBody *)calloc(MAX_PHYSICS_BODIES, sizeof(struct PhysicsBody));
  world->player_bodies = (struct PhysicsBody *)calloc(MAX_PHYSICS_BODIES, sizeof(struct PhysicsBody));
  if (!world->static_bodies || !world->dynamic_bodies || !world->player_bodies) {
    printf("Error: failed to allocate PhysicsWorld in physics_world_create\n");
    free(world->static_bodies); free(world->dynamic_bodies); free(world->player_bodies); free(world);
    return NULL;
  }
  return world;
}

void physics_world_release_gpu(struct PhysicsWorld *world) {
  Shadow **pp = &g_shadows;
  while (*pp) {
    Shadow *s = *pp;
    if (s->world == world) {
      *pp = s->next;
      if (s->gpu) pg_worlds_destroy(s->gpu);
      for (int c = 0; c < 3; c++) { free(s->rec[c]); free(s->dev[c]); }
      free(s->events); free(s->up);
      free(s);
      return;
    }
    pp = &s->next;
  }
}

static void fill_common(struct PhysicsBody *body, struct SceneNode *scene_node, struct Entity *entity, struct Collider collider) {
  memcpy(body->position, entity->position, 12);
  memcpy(body->rotation, entity->rotation, 12);
  memcpy(body->scale, entity->scale, 12);
  body->collider = collider;
  body->entity = entity;
  body->scene_node = scene_node;
}

/* world.c:41-121.  Like the reference the slot is taken BEFORE the type check (a rejected body
 * leaves a zeroed hole that the step then visits); unlike it, a full array is refused instead of
 * written past (world.c:46,50 have no bounds check). */
struct PhysicsBody *physics_add_body(struct PhysicsWorld *physics_world, struct SceneNode *scene_node, struct Entity *entity,
                                     struct Collider collider, float restitution, bool dynamic) {
  struct PhysicsBody *body;
  if (dynamic) {
    if (physics_world->num_dynamic_bodies >= MAX_PHYSICS_BODIES) { fprintf(stderr, "Error: physics_add_body: dynamic body array is full (%d)\n", MAX_PHYSICS_BODIES); return NULL; }
    body = &physics_world->dynamic_bodies[physics_world->num_dynamic_bodies++];
    memcpy(body->velocity, entity->velocity, 12);
  } else {
    if (physics_world->num_static_bodies >= MAX_PHYSICS_BODIES) { fprintf(stderr, "Error: physics_add_body: static body array is full (%d)\n", MAX_PHYSICS_BODIES); return NULL; }
    body = &physics_world->static_bodies[physics_world->num_static_bodies++];
    body->velocity[0] = body->velocity[1] = body->velocity[2] = 0.0f;
  }
  if ((int)collider.type < 0 || collider.type > COLLIDER_COUNT) {
    fprintf(stderr, "Error: collider type provided to physics_add_body is invalid\n");
    return NULL;
  }
  fill_common(body, scene_node, entity, collider);
  body->restitution = restitution;
  body->dynamic = dynamic;
  return body;
}

/* world.c:123-142 */
struct PhysicsBody *physics_add_player(struct PhysicsWorld *physics_world, struct SceneNode *scene_node, struct Entity *entity,
                                       struct Collider collider) {
  if ((int)collider.type < 0 || collider.type > COLLIDER_COUNT) {
    fprintf(stderr, "Error: collider type provided to physics_add_body is invalid\n");
    return NULL;
  }
  if (physics_world->num_player_bodies >= MAX_PHYSICS_BODIES) { fprintf(stderr, "Error: physics_add_player: player body array is full (%d)\n", MAX_PHYSICS_BODIES); return NULL; }
  struct PhysicsBody *body = &physics_world->player_bodies[physics_world->num_player_bodies++];
  fill_common(body, scene_node, entity, collider);
  memcpy(body->velocity, entity->velocity, 12);
  body->restitution = 0.0f;
  return body;
}

/* world.c:147-172: players, then statics, then dynamics; first entity-id match wins */
void physics_remove_body(struct PhysicsWorld *physics_world, struct PhysicsBody *physics_body) {
  for (int cls = 2; ; cls = (cls == 2 ? 0 : cls + 1)) {
    unsigned int n;
    struct PhysicsBody *arr = class_array(physics_world, cls, &n);
    for (unsigned int i = 0; i < n; i++) {
      if (memcmp(arr[i].entity->id, physics_body->entity->id, 16) == 0) {
        arr[i] = arr[n - 1];
        arr[i].entity->physics_body = &arr[i];
        if (cls == PG_CLASS_STATIC) physics_world->num_static_bodies--;
        else if (cls == PG_CLASS_DYNAMIC) physics_world->num_dynamic_bodies--;
        else physics_world->num_player_bodies--;
        return;
      }
    }
    if (cls == 1) return;
  }
}

/* declared in world.h:45, defined nowhere upstream: copy body pose into the owning entities */
void physics_sync_entities(struct PhysicsWorld *physics_world) {
  for (int cls = 0; cls < 3; cls++) {
    unsigned int n;
    struct PhysicsBody *arr = class_array(physics_world, cls, &n);
    for (unsigned int i = 0; i < n; i++) {
      if (!arr[i].entity) continue;
      memcpy(arr[i].entity->position, arr[i].position, 12);
      memcpy(arr[i].entity->rotation, arr[i].rotation, 12);
    }
  }
}

/* ------------------------------------------------------------------ step */

static int ensure_gpu(Shadow *s) {
  struct PhysicsWorld *w = s->world;
  int need[3] = {(int)w->num_static_bodies, (int)w->num_dynamic_bodies, (int)w->num_player_bodies};
  int grow = !s->gpu;
  for (int c = 0; c < 3; c++) if (need[c] > s->cap[c]) grow = 1;
  if (grow) {
    if (s->gpu) pg_worlds_destroy(s->gpu);
    for (int c = 0; c < 3; c++) { s->cap[c] = need[c] > MAX_PHYSICS_BODIES ? need[c] * 2 : MAX_PHYSICS_BODIES; }
    s->gpu = pg_worlds_create(1, s->cap[0], s->cap[1], s->cap[2], CRUX_MAX_EVENTS, 0);
    for (int c = 0; c < 3; c++) s->dev_n[c] = -1;
    if (!s->gpu) {
      fprintf(stderr, "Error: physics_step: no usable CUDA device (%s); the step is skipped\n", pg_last_error());
      return -1;
    }
    if (!s->events) s->events = (PgEvent *)malloc(sizeof(PgEvent) * CRUX_MAX_EVENTS);
  }
  for (int c = 0; c < 3; c++) {
    if (need[c] > s->rec_cap[c]) {
      s->rec_cap[c] = need[c] > MAX_PHYSICS_BODIES ? need[c] * 2 : MAX_PHYSICS_BODIES;
      s->rec[c] = (PgBody *)realloc(s->rec[c], sizeof(PgBody) * (size_t)s->rec_cap[c]);
      s->dev[c] = (PgBody *)realloc(s->dev[c], sizeof(PgBody) * (size_t)s->rec_cap[c]);
      s->dev_n[c] = -1;
    }
  }
  return 0;
}

static void replay_events(Shadow *s, long long n_events) {
  struct PhysicsWorld *w = s->world;
  for (long long k = 0; k < n_events; k++) {
    const PgEvent *e = &s->events[k];
    unsigned int na, nb;
    struct PhysicsBody *A = class_array(w, e->a_class, &na) + e->a_index;
    struct PhysicsBody *B = class_array(w, e->b_class, &nb) + e->b_index;
    struct GameEvent event;
    memset(&event, 0, sizeof(event));
    struct timespec timestamp;
    if (clock_gettime(CLOCK_REALTIME, &timestamp) == -1) {
      perror("clock_gettime");
      timestamp.tv_nsec = 0;
    }
    event.timestamp = timestamp;
    event.type = get_event_type(A->entity->type, B->entity->type);
    switch (event.type) {
      case EVENT_COLLISION:
        memcpy(event.data.collision.entity_A_id, A->entity->id, 16);
        memcpy(event.data.collision.entity_B_id, B->entity->id, 16);
        break;
      case EVENT_PLAYER_ITEM_PICKUP:      /* A is the item (lower collider type), B the player: world.c:250-256 */
        memcpy(event.data.item_pickup.player_entity_id, B->entity->id, 16);
        event.data.item_pickup.item_id = A->entity->item ? A->entity->item->id : 0;
        event.data.item_pickup.item_count = A->entity->item ? A->entity->item->count : 0;
        memcpy(event.data.item_pickup.item_entity_id, A->entity->id, 16);
        break;
    }
    game_event_queue_enqueue(event);
  }
}

static void step_impl(struct PhysicsWorld *physics_world, float delta_time, int n_steps, int refresh_nodes) {
  Shadow *s = shadow_find(physics_world, 1);
  if (!s || ensure_gpu(s)) return;
  /* The host arrays are the truth (player.c and the scene poke them between frames), the device keeps a copy:
   * only the records that differ from what it holds travel, as contiguous dirty ranges (a class whose length
   * changed -- physics_add_body / physics_remove_body -- goes whole).  The whole frame is ONE call into the
   * C-ABI and one synchronisation (pg_worlds_frame): a game-sized scene is bound by round trips, not by work. */
  PgRange ranges[3 * CRUX_MAX_RANGES];
  int n_ranges = 0;
  int32_t counts[3];
  size_t n_up = 0;
  for (int c = 0; c < 3; c++) {
    unsigned int n;
    struct PhysicsBody *arr = class_array(physics_world, c, &n);
    counts[c] = (int32_t)n;
    for (unsigned int i = 0; i < n; i++) crux_flatten_body(&arr[i], &s->rec[c][i]);
    if (s->dev_n[c] != (int)n) {
      if (n) { ranges[n_ranges].cls = c; ranges[n_ranges].first = 0; ranges[n_ranges].count = (int32_t)n; n_ranges++; }
      continue;
    }
    const int first_range = n_ranges;
    unsigned int i = 0;
    while (i < n) {
      if (memcmp(&s->rec[c][i], &s->dev[c][i], sizeof(PgBody)) == 0) { i++; continue; }
      unsigned int j = i + 1;
      while (j < n && memcmp(&s->rec[c][j], &s->dev[c][j], sizeof(PgBody)) != 0) j++;
      if (n_ranges - first_range == CRUX_MAX_RANGES) {
        /* too fragmented: one range from the first dirty record of the class to this run's end */
        n_ranges = first_range + 1;
        ranges[first_range].count = (int32_t)j - ranges[first_range].first;
      } else {
        ranges[n_ranges].cls = c; ranges[n_ranges].first = (int32_t)i; ranges[n_ranges].count = (int32_t)(j - i); n_ranges++;
      }
      i = j;
    }
  }
  for (int r = 0; r < n_ranges; r++) n_up += (size_t)ranges[r].count;
  if (n_up > s->up_cap) {
    s->up_cap = n_up * 2 + 16;
    s->up = (PgBody *)realloc(s->up, sizeof(PgBody) * s->up_cap);
  }
  size_t at = 0;
  for (int r = 0; r < n_ranges; r++) {
    memcpy(s->up + at, &s->rec[ranges[r].cls][ranges[r].first], sizeof(PgBody) * (size_t)ranges[r].count);
    at += (size_t)ranges[r].count;
  }
  long long n_events = 0;
  int overflow = 0;
  if (pg_worlds_frame(s->gpu, 0, counts, ranges, n_ranges, s->up, delta_time, n_steps, refresh_nodes,
                      s->rec[0], s->rec[1], s->rec[2], s->events, CRUX_MAX_EVENTS, &n_events, &overflow) != PG_OK) {
    fprintf(stderr, "Error: physics_step: %s\n", pg_last_error());
    for (int c = 0; c < 3; c++) s->dev_n[c] = -1;
    return;
  }
  s->uploaded_records += n_up;
  for (int c = 0; c < 3; c++) {
    unsigned int n;
    struct PhysicsBody *arr = class_array(physics_world, c, &n);
    s->dev_n[c] = (int)n;
    if (n == 0) continue;
    memcpy(s->dev[c], s->rec[c], sizeof(PgBody) * (size_t)n);
    for (unsigned int i = 0; i < n; i++) {
      unflatten_state(&s->rec[c][i], &arr[i]);
      if (refresh_nodes && arr[i].scene_node) {
        /* what scene_node_update leaves behind (scene.c:1054-1079) */
        struct SceneNode *node = arr[i].scene_node;
        memcpy(node->world_transform, s->rec[c][i].node_xform, 64);
        memcpy(node->position, arr[i].position, 12);
        memcpy(node->rotation, arr[i].rotation, 12);
        if (arr[i].entity) { memcpy(arr[i].entity->position, arr[i].position, 12); memcpy(arr[i].entity->rotation, arr[i].rotation, 12); }
      }
    }
  }
  if (n_events > CRUX_MAX_EVENTS) n_events = CRUX_MAX_EVENTS;
  if (overflow) fprintf(stderr, "Error: physics_step: more than %d contacts in one call, the rest are dropped\n", CRUX_MAX_EVENTS);
  if (n_events > 0) replay_events(s, n_events);
}

/* What physics_debug_render (src/physics/debug_renderer.c:11-213) reads from the bodies each frame, produced on the
 * device: the records the host changed since the last step go up first (a zero-step frame, same dirty-range path as
 * physics_step), then one kernel writes a PgDebugDraw per body (players, statics, dynamics) into `dst` -- host memory,
 * or with dst_is_device a device address such as a GL buffer mapped through cudaGraphicsResourceGetMappedPointer. */
int physics_debug_geometry(struct PhysicsWorld *physics_world, void *dst, int dst_is_device, int max_records) {
  Shadow *s = shadow_find(physics_world, 1);
  if (!s || ensure_gpu(s)) return PG_ERR_CUDA;
  step_impl(physics_world, 0.0f, 0, 0);
  int n = pg_worlds_debug_geometry(s->gpu, 0, dst, dst_is_device, max_records);
  if (n < 0) fprintf(stderr, "Error: physics_debug_geometry: %s\n", pg_last_error());
  return n;
}

unsigned long long physics_world_uploaded_records(struct PhysicsWorld *world) {
  Shadow *s = shadow_find(world, 0);
  return s ? s->uploaded_records : 0ull;
}

/* world.c:174-555 */
void physics_step(struct PhysicsWorld *physics_world, float delta_time) { step_impl(physics_world, delta_time, 1, 0); }

void physics_step_n(struct PhysicsWorld *physics_world, float delta_time, int n_steps) {
  /* in chunks, so that the contact list of one device call (CRUX_MAX_EVENTS) is replayed before it can fill up */
  while (n_steps > 0) {
    const int k = n_steps < 16 ? n_steps : 16;
    step_impl(physics_world, delta_time, k, 1);
    n_steps -= k;
  }
}

/* ------------------------------------------------------------------ public predicates (world.h:47-50) */

bool interval_collision(struct PhysicsBody *body_A, struct PhysicsBody *body_B, float start_time, float end_time, float *hit_time) {
  PgBody a, b;
  crux_flatten_body(body_A, &a);
  crux_flatten_body(body_B, &b);
  int32_t hit = 0;
  float t = 0.0f;
  if (pg_pair_interval_collision(&a, &b, 1, start_time, end_time, &hit, &t) != PG_OK) {
    fprintf(stderr, "Error: interval_collision: %s\n", pg_last_error());
    return false;
  }
  if (hit && hit_time) *hit_time = t;
  return hit != 0;
}

float maximum_object_movement_over_time(struct PhysicsBody *body, float start_time, float end_time) {
  PgBody a;
  crux_flatten_body(body, &a);
  float out = 0.0f;
  if (pg_pair_max_move(&a, 1, start_time, end_time, &out) != PG_OK) fprintf(stderr, "Error: maximum_object_movement_over_time: %s\n", pg_last_error());
  return out;
}

float minimum_object_distance_at_time(struct PhysicsBody *body_A, struct PhysicsBody *body_B, float time) {
  if (!distance_functions[body_A->collider.type][body_B->collider.type]) {       /* world.c:593-598 */
    fprintf(stderr, "Error: no minimum distance function found for types %d and %d\n", body_A->collider.type, body_B->collider.type);
    return 0;
  }
  PgBody a, b;
  crux_flatten_body(body_A, &a);
  crux_flatten_body(body_B, &b);
  float out = 0.0f;
  if (pg_pair_min_dist(&a, &b, 1, &time, &out) != PG_OK) fprintf(stderr, "Error: minimum_object_distance_at_time: %s\n", pg_last_error());
  return out;
}
