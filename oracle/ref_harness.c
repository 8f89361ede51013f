/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * Headless driver around the UNMODIFIED reference physics sources.  oracle/Makefile compiles
 * /root/reference/src/physics/{world,distance,narrow_phase,resolution,aabb,collider,sphere,utils}.c
 * where they lie and links them with this file into oracle/_ref/libcrux_ref.so.  Nothing here
 * re-implements physics: it only
 *   (1) provides the two game-layer symbols the physics step calls out to
 *       (get_event_type: /root/reference/src/event.c:14-20,139-141; game_event_queue_enqueue:
 *       event.c:62-72) -- the latter records the contact instead of queueing audio/pickup work;
 *   (2) builds PhysicsBody / Entity / SceneNode objects from flat OraBody records, growing the
 *       world arrays past the 128-body allocation of physics_world_create (world.c:15,32-34;
 *       `struct PhysicsWorld` is public, world.h:28-35);
 *   (3) after every physics_step refreshes scene-node transforms exactly as scene_update does
 *       (scene.c:401-408 -> scene_node_update, scene.c:1051-1079), without which no AABB or
 *       capsule path ever sees a body move (SURVEY.md section 8c "headless stepping recipe").
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>

#include "entity.h"
#include "item.h"
#include "scene.h"
#include "event.h"
#include "physics/world.h"
#include "narrow_phase.h"
#include "distance.h"
#include "resolution.h"

#include "oracle_abi.h"

/* ------------------------------------------------------------------ world wrapper */

typedef struct RefWorld {
  struct PhysicsWorld *pw;
  unsigned int cap[3];           /* capacity of static / dynamic / player arrays */
  OraEvent *events;
  int n_events, cap_events;
  int step_index;
  int record_events;
} RefWorld;

static RefWorld *g_current = NULL;   /* world whose physics_step is running (single-threaded) */

static struct ItemComponent g_dummy_item = {0, 1};

/* ---- symbols the reference physics step links against (event.c) ---- */

EventType get_event_type(EntityType type_A, EntityType type_B) {
  /* table restated from /root/reference/src/event.c:14-20 */
  if ((type_A == ENTITY_ITEM && type_B == ENTITY_PLAYER) ||
      (type_A == ENTITY_PLAYER && type_B == ENTITY_ITEM))
    return EVENT_PLAYER_ITEM_PICKUP;
  return EVENT_COLLISION;
}

static void locate(RefWorld *w, const unsigned char *id, int32_t *cls, int32_t *idx) {
  struct Entity *e;
  memcpy(&e, id, sizeof(e));
  struct PhysicsBody *b = e->physics_body;
  struct PhysicsWorld *pw = w->pw;
  if (b >= pw->static_bodies && b < pw->static_bodies + w->cap[0]) {
    *cls = ORA_CLASS_STATIC; *idx = (int32_t)(b - pw->static_bodies);
  } else if (b >= pw->dynamic_bodies && b < pw->dynamic_bodies + w->cap[1]) {
    *cls = ORA_CLASS_DYNAMIC; *idx = (int32_t)(b - pw->dynamic_bodies);
  } else {
    *cls = ORA_CLASS_PLAYER; *idx = (int32_t)(b - pw->player_bodies);
  }
}

void game_event_queue_enqueue(struct GameEvent ev) {
  RefWorld *w = g_current;
  if (!w || !w->record_events) return;
  if (w->n_events == w->cap_events) {
    w->cap_events = w->cap_events ? w->cap_events * 2 : 1024;
    w->events = (OraEvent *)realloc(w->events, (size_t)w->cap_events * sizeof(OraEvent));
  }
  OraEvent *o = &w->events[w->n_events++];
  o->step = w->step_index;
  o->event_type = (int32_t)ev.type;
  if (ev.type == EVENT_COLLISION) {
    locate(w, ev.data.collision.entity_A_id, &o->a_class, &o->a_index);
    locate(w, ev.data.collision.entity_B_id, &o->b_class, &o->b_index);
  } else {   /* item pickup: A = item, B = player (world.c:250-256) */
    locate(w, ev.data.item_pickup.item_entity_id, &o->a_class, &o->a_index);
    locate(w, ev.data.item_pickup.player_entity_id, &o->b_class, &o->b_index);
  }
}

/* ---- body construction ---- */

static void fill_collider(const OraBody *o, struct Collider *c) {
  memset(c, 0, sizeof(*c));
  c->type = (ColliderType)o->collider_type;
  switch (o->collider_type) {
    case COLLIDER_AABB:
      memcpy(c->data.aabb.center, o->shape, 12);
      memcpy(c->data.aabb.extents, o->shape + 3, 12);
      c->data.aabb.initialized = true;
      break;
    case COLLIDER_SPHERE:
      memcpy(c->data.sphere.center, o->shape, 12);
      c->data.sphere.radius = o->shape[3];
      break;
    case COLLIDER_CAPSULE:
      memcpy(c->data.capsule.segment_A, o->shape, 12);
      memcpy(c->data.capsule.segment_B, o->shape + 3, 12);
      c->data.capsule.radius = o->shape[6];
      break;
    default:
      memcpy(c->data.plane.normal, o->shape, 12);
      c->data.plane.distance = o->shape[3];
      break;
  }
}

static void read_collider(const struct Collider *c, OraBody *o) {
  memset(o->shape, 0, sizeof(o->shape));
  o->collider_type = (int32_t)c->type;
  switch (c->type) {
    case COLLIDER_AABB:
      memcpy(o->shape, c->data.aabb.center, 12);
      memcpy(o->shape + 3, c->data.aabb.extents, 12);
      break;
    case COLLIDER_SPHERE:
      memcpy(o->shape, c->data.sphere.center, 12);
      o->shape[3] = c->data.sphere.radius;
      break;
    case COLLIDER_CAPSULE:
      memcpy(o->shape, c->data.capsule.segment_A, 12);
      memcpy(o->shape + 3, c->data.capsule.segment_B, 12);
      o->shape[6] = c->data.capsule.radius;
      break;
    default:
      memcpy(o->shape, c->data.plane.normal, 12);
      o->shape[3] = c->data.plane.distance;
      break;
  }
}

/* Entity (+ optional SceneNode and parent SceneNode) backing one body. */
static struct Entity *make_entity(const OraBody *o) {
  struct Entity *e = (struct Entity *)calloc(1, sizeof(struct Entity));
  memcpy(e->id, &e, sizeof(e));            /* id = own address: unique, and lets locate() work */
  e->type = (EntityType)o->entity_type;
  memcpy(e->position, o->position, 12);
  memcpy(e->rotation, o->rotation, 12);
  memcpy(e->scale, o->scale, 12);
  memcpy(e->velocity, o->velocity, 12);
  e->item = &g_dummy_item;
  return e;
}

static struct SceneNode *make_node(const OraBody *o, struct Entity *e) {
  if (!o->has_node) return NULL;
  struct SceneNode *n = (struct SceneNode *)calloc(1, sizeof(struct SceneNode));
  memcpy(n->entity_id, e->id, 16);
  n->entity = e;
  memcpy(n->position, o->position, 12);
  memcpy(n->rotation, o->rotation, 12);
  memcpy(n->scale, o->scale, 12);
  memcpy(n->world_transform, o->node_xform, 64);
  memcpy(n->local_transform, o->node_xform, 64);
  if (o->has_parent) {
    struct SceneNode *p = (struct SceneNode *)calloc(1, sizeof(struct SceneNode));
    memcpy(p->world_transform, o->parent_xform, 64);
    memcpy(p->local_transform, o->parent_xform, 64);
    n->parent_node = p;
  }
  return n;
}

static void free_body_aux(struct PhysicsBody *b) {
  if (b->scene_node) {
    free(b->scene_node->parent_node);
    free(b->scene_node);
  }
  free(b->entity);
}

static void body_to_ora(const struct PhysicsBody *b, OraBody *o) {
  memset(o, 0, sizeof(*o));
  memcpy(o->position, b->position, 12);
  memcpy(o->rotation, b->rotation, 12);
  memcpy(o->scale, b->scale, 12);
  memcpy(o->velocity, b->velocity, 12);
  o->restitution = b->restitution;
  read_collider(&b->collider, o);
  o->entity_type = b->entity ? (int32_t)b->entity->type : 0;
  o->dynamic = b->dynamic;
  o->at_rest = b->at_rest;
  if (b->scene_node) {
    o->has_node = 1;
    memcpy(o->node_xform, b->scene_node->world_transform, 64);
    if (b->scene_node->parent_node) {
      o->has_parent = 1;
      memcpy(o->parent_xform, b->scene_node->parent_node->world_transform, 64);
    }
  }
}

/* ------------------------------------------------------------------ exported: world API */

void *ref_world_create(void) {
  RefWorld *w = (RefWorld *)calloc(1, sizeof(RefWorld));
  w->pw = physics_world_create();
  w->cap[0] = w->cap[1] = w->cap[2] = 128;   /* MAX_PHYSICS_BODIES, world.c:15 */
  w->record_events = 1;
  return w;
}

/* the wrapped struct PhysicsWorld, for oracle/ref_debug_harness.c (which drives the reference's debug renderer) */
void *ref_world_pw(void *h) { return ((RefWorld *)h)->pw; }

void ref_world_destroy(void *h) {
  RefWorld *w = (RefWorld *)h;
  struct PhysicsWorld *pw = w->pw;
  for (unsigned i = 0; i < pw->num_static_bodies; i++) free_body_aux(&pw->static_bodies[i]);
  for (unsigned i = 0; i < pw->num_dynamic_bodies; i++) free_body_aux(&pw->dynamic_bodies[i]);
  for (unsigned i = 0; i < pw->num_player_bodies; i++) free_body_aux(&pw->player_bodies[i]);
  free(pw->static_bodies); free(pw->dynamic_bodies); free(pw->player_bodies);
  free(pw); free(w->events); free(w);
}

static void repoint(struct PhysicsBody *arr, unsigned n) {
  for (unsigned i = 0; i < n; i++) arr[i].entity->physics_body = &arr[i];
}

static void ensure_cap(RefWorld *w, int cls, unsigned need) {
  if (need <= w->cap[cls]) return;
  unsigned nc = w->cap[cls];
  while (nc < need) nc *= 2;
  struct PhysicsWorld *pw = w->pw;
  struct PhysicsBody **arr = cls == 0 ? &pw->static_bodies : cls == 1 ? &pw->dynamic_bodies : &pw->player_bodies;
  unsigned n = cls == 0 ? pw->num_static_bodies : cls == 1 ? pw->num_dynamic_bodies : pw->num_player_bodies;
  struct PhysicsBody *na = (struct PhysicsBody *)calloc(nc, sizeof(struct PhysicsBody));
  memcpy(na, *arr, (size_t)n * sizeof(struct PhysicsBody));
  free(*arr);
  *arr = na;
  w->cap[cls] = nc;
  repoint(na, n);
}

/* cls: ORA_CLASS_PLAYER -> physics_add_player, else physics_add_body with o->dynamic. */
int ref_world_add(void *h, const OraBody *o, int as_player) {
  RefWorld *w = (RefWorld *)h;
  struct PhysicsWorld *pw = w->pw;
  struct Collider c;
  fill_collider(o, &c);
  struct Entity *e = make_entity(o);
  struct SceneNode *n = make_node(o, e);
  struct PhysicsBody *b;
  if (as_player) {
    ensure_cap(w, 2, pw->num_player_bodies + 1);
    b = physics_add_player(pw, n, e, c);
  } else {
    int cls = o->dynamic ? 1 : 0;
    ensure_cap(w, cls, (cls ? pw->num_dynamic_bodies : pw->num_static_bodies) + 1);
    b = physics_add_body(pw, n, e, c, o->restitution, o->dynamic != 0);
  }
  if (!b) return -1;
  b->at_rest = o->at_rest != 0;
  e->physics_body = b;
  if (as_player) return (int)(b - pw->player_bodies);
  return (int)(o->dynamic ? b - pw->dynamic_bodies : b - pw->static_bodies);
}

int ref_world_add_many(void *h, const OraBody *o, int n, int as_player) {
  int last = -1;
  for (int i = 0; i < n; i++) last = ref_world_add(h, &o[i], as_player);
  return last;
}

int ref_world_remove(void *h, int cls, int index) {
  RefWorld *w = (RefWorld *)h;
  struct PhysicsWorld *pw = w->pw;
  struct PhysicsBody *arr = cls == 0 ? pw->static_bodies : cls == 1 ? pw->dynamic_bodies : pw->player_bodies;
  struct PhysicsBody victim = arr[index];
  physics_remove_body(pw, &arr[index]);
  free_body_aux(&victim);
  return 0;
}

void ref_world_counts(void *h, int *counts) {
  RefWorld *w = (RefWorld *)h;
  counts[0] = (int)w->pw->num_static_bodies;
  counts[1] = (int)w->pw->num_dynamic_bodies;
  counts[2] = (int)w->pw->num_player_bodies;
}

void ref_world_read(void *h, int cls, OraBody *out) {
  RefWorld *w = (RefWorld *)h;
  struct PhysicsWorld *pw = w->pw;
  struct PhysicsBody *arr = cls == 0 ? pw->static_bodies : cls == 1 ? pw->dynamic_bodies : pw->player_bodies;
  unsigned n = cls == 0 ? pw->num_static_bodies : cls == 1 ? pw->num_dynamic_bodies : pw->num_player_bodies;
  for (unsigned i = 0; i < n; i++) body_to_ora(&arr[i], &out[i]);
}

/* Overwrite the mutable state of one body (what player.c:27-45,155-172 does through
 * entity->physics_body): position, rotation, velocity, at_rest. */
void ref_world_poke(void *h, int cls, int index, const OraBody *o) {
  RefWorld *w = (RefWorld *)h;
  struct PhysicsWorld *pw = w->pw;
  struct PhysicsBody *b = &(cls == 0 ? pw->static_bodies : cls == 1 ? pw->dynamic_bodies : pw->player_bodies)[index];
  memcpy(b->position, o->position, 12);
  memcpy(b->rotation, o->rotation, 12);
  memcpy(b->velocity, o->velocity, 12);
  b->at_rest = o->at_rest != 0;
}

/* scene_node_update (scene.c:1051-1079) for one node; parents are static in the headless harness. */
static void refresh_node(struct PhysicsBody *b) {
  struct SceneNode *n = b->scene_node;
  if (!n) return;
  glm_vec3_copy(b->position, n->position);
  glm_vec3_copy(b->position, b->entity->position);
  glm_vec3_copy(b->rotation, n->rotation);
  glm_vec3_copy(b->rotation, b->entity->rotation);
  glm_mat4_identity(n->local_transform);
  glm_translate(n->local_transform, n->position);
  vec3 rotation_radians = { glm_rad(n->rotation[0]), glm_rad(n->rotation[1]), glm_rad(n->rotation[2]) };
  mat4 rotation;
  glm_euler_xyz(rotation_radians, rotation);
  glm_mul(n->local_transform, rotation, n->local_transform);
  glm_scale(n->local_transform, n->scale);
  if (n->parent_node) glm_mat4_mul(n->parent_node->world_transform, n->local_transform, n->world_transform);
  else glm_mat4_copy(n->local_transform, n->world_transform);
}

void ref_world_refresh_nodes(void *h) {
  RefWorld *w = (RefWorld *)h;
  struct PhysicsWorld *pw = w->pw;
  for (unsigned i = 0; i < pw->num_static_bodies; i++) refresh_node(&pw->static_bodies[i]);
  for (unsigned i = 0; i < pw->num_dynamic_bodies; i++) refresh_node(&pw->dynamic_bodies[i]);
  for (unsigned i = 0; i < pw->num_player_bodies; i++) refresh_node(&pw->player_bodies[i]);
}

/* nsteps x { physics_step ; scene_node_update }.  refresh_nodes=0 reproduces a caller that never
 * updates the scene graph (sphere/plane worlds are insensitive to it). */
void ref_world_step(void *h, float dt, int nsteps, int refresh_nodes) {
  RefWorld *w = (RefWorld *)h;
  g_current = w;
  for (int s = 0; s < nsteps; s++) {
    physics_step(w->pw, dt);
    if (refresh_nodes) ref_world_refresh_nodes(h);
    w->step_index++;
  }
  g_current = NULL;
}

void ref_world_record_events(void *h, int on) { ((RefWorld *)h)->record_events = on; }
int ref_events_count(void *h) { return ((RefWorld *)h)->n_events; }
void ref_events_read(void *h, OraEvent *out) {
  RefWorld *w = (RefWorld *)h;
  memcpy(out, w->events, (size_t)w->n_events * sizeof(OraEvent));
}
void ref_events_clear(void *h) { RefWorld *w = (RefWorld *)h; w->n_events = 0; w->step_index = 0; }

/* ------------------------------------------------------------------ exported: per-pair KATs */

typedef struct TmpBody {
  struct PhysicsBody b;
  struct Entity e;
  struct SceneNode n, p;
} TmpBody;

static void tmp_from_ora(const OraBody *o, TmpBody *t) {
  memset(t, 0, sizeof(*t));
  fill_collider(o, &t->b.collider);
  memcpy(t->b.position, o->position, 12);
  memcpy(t->b.rotation, o->rotation, 12);
  memcpy(t->b.scale, o->scale, 12);
  memcpy(t->b.velocity, o->velocity, 12);
  t->b.restitution = o->restitution;
  t->b.dynamic = o->dynamic != 0;
  t->b.at_rest = o->at_rest != 0;
  t->e.type = (EntityType)o->entity_type;
  t->e.item = &g_dummy_item;
  t->e.physics_body = &t->b;
  t->b.entity = &t->e;
  if (o->has_node) {
    memcpy(t->n.world_transform, o->node_xform, 64);
    memcpy(t->n.scale, o->scale, 12);
    t->n.entity = &t->e;
    if (o->has_parent) {
      memcpy(t->p.world_transform, o->parent_xform, 64);
      t->n.parent_node = &t->p;
    }
    t->b.scene_node = &t->n;
  }
}

float ref_max_move(const OraBody *a, float t0, float t1) {
  TmpBody A; tmp_from_ora(a, &A);
  return maximum_object_movement_over_time(&A.b, t0, t1);
}

float ref_min_dist(const OraBody *a, const OraBody *b, float t) {
  TmpBody A, B; tmp_from_ora(a, &A); tmp_from_ora(b, &B);
  return minimum_object_distance_at_time(&A.b, &B.b, t);
}

int ref_interval_collision(const OraBody *a, const OraBody *b, float t0, float t1, float *hit) {
  TmpBody A, B; tmp_from_ora(a, &A); tmp_from_ora(b, &B);
  float ht = -1.0f;
  int r = interval_collision(&A.b, &B.b, t0, t1, &ht);
  if (hit) *hit = ht;
  return r;
}

/* returns result.colliding; *hit_time = result.hit_time */
int ref_narrow_phase(const OraBody *a, const OraBody *b, float dt, float *hit_time) {
  TmpBody A, B; tmp_from_ora(a, &A); tmp_from_ora(b, &B);
  NarrowPhaseFunction f = narrow_phase_functions[A.b.collider.type][B.b.collider.type];
  if (!f) { *hit_time = 0.0f; return -1; }
  struct CollisionResult r = f(&A.b, &B.b, dt);
  *hit_time = r.hit_time;
  return r.colliding ? 1 : 0;
}

/* in/out: position, velocity, at_rest of both records are overwritten with the resolver's output */
int ref_resolve(OraBody *a, OraBody *b, float hit_time, float dt) {
  TmpBody A, B; tmp_from_ora(a, &A); tmp_from_ora(b, &B);
  ResolutionFunction f = resolution_functions[A.b.collider.type][B.b.collider.type];
  if (!f) return -1;
  struct CollisionResult r;
  memset(&r, 0, sizeof(r));
  r.hit_time = hit_time;
  r.colliding = true;
  f(&A.b, &B.b, r, dt);
  memcpy(a->position, A.b.position, 12); memcpy(a->velocity, A.b.velocity, 12); a->at_rest = A.b.at_rest;
  memcpy(b->position, B.b.position, 12); memcpy(b->velocity, B.b.velocity, 12); b->at_rest = B.b.at_rest;
  return 0;
}

/* One full visit of world.c:481-547 on a detached pair: type-order swap, broadphase gate,
 * narrowphase, behaviour lookup, resolution.  Returns 1 if an event would be emitted. */
int ref_visit(OraBody *a, OraBody *b, float dt) {
  OraBody *x = a, *y = b;
  if (x->collider_type > y->collider_type) { OraBody *t = x; x = y; y = t; }
  TmpBody A, B; tmp_from_ora(x, &A); tmp_from_ora(y, &B);
  float ht;
  struct CollisionResult r;
  memset(&r, 0, sizeof(r));
  if (interval_collision(&A.b, &B.b, 0, dt, &ht)) {
    NarrowPhaseFunction f = narrow_phase_functions[A.b.collider.type][B.b.collider.type];
    if (f) r = f(&A.b, &B.b, dt);
  }
  if (!(r.colliding && r.hit_time >= 0)) return 0;
  if (get_collision_behavior(A.e.type, B.e.type) == COLLISION_BEHAVIOR_PHYSICS) {
    ResolutionFunction g = resolution_functions[A.b.collider.type][B.b.collider.type];
    if (g) g(&A.b, &B.b, r, dt);
  }
  memcpy(x->position, A.b.position, 12); memcpy(x->velocity, A.b.velocity, 12); x->at_rest = A.b.at_rest;
  memcpy(y->position, B.b.position, 12); memcpy(y->velocity, B.b.velocity, 12); y->at_rest = B.b.at_rest;
  return 1;
}

/* struct sizes, so the tests can record what this toolchain laid out (SURVEY.md section 8 header) */
void ref_sizes(int *out) {
  out[0] = (int)sizeof(struct PhysicsBody);
  out[1] = (int)sizeof(struct Collider);
  out[2] = (int)sizeof(struct CollisionResult);
  out[3] = (int)sizeof(struct Entity);
  out[4] = (int)sizeof(struct SceneNode);
  out[5] = (int)sizeof(struct GameEvent);
  out[6] = (int)sizeof(struct AABB);
}
