/*
 * crux_scene.c -- headless scene loader for the physics path (see crux_scene.h).
 *
 * Follows /root/reference/src/scene.c:59-393 (scene_load), :606-908 (scene_process_node_json),
 * :965-1050 (scene_player_create), :1051-1079 (scene_node_update), :1106-1153 (scene_node_create),
 * keeping only what reaches the physics world.  Differences, all on purpose (SURVEY.md section 5
 * "Scene schema"):
 *   - a node without a "components" array is accepted (Gen-2 files such as scenes/bouncehouse.json;
 *     the current upstream loader stops recursing there, scene.c:797-801, and yields zero bodies);
 *   - a file with a flat "meshes": {"static": [...], "dynamic": [...]} layout (Gen-1) is accepted,
 *     with that generation's collider enum (2 = plane; today 2 = capsule, 3 = plane);
 *   - JSON is read by the ~150-line reader below instead of cJSON (third-party code is not copied).
 * Matrix arithmetic restates cglm's scalar glm_mat4_mul / glm_euler_xyz / glm_rad order, as
 * oracle/shim/cglm does, so world transforms are bit-identical to a cglm build's.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crux_scene.h"

/* ------------------------------------------------------------------ a small JSON reader */

typedef enum { J_NULL, J_BOOL, J_NUM, J_STR, J_ARR, J_OBJ } JType;
typedef struct JVal {
  JType type;
  double num;              /* J_NUM, J_BOOL (0/1) */
  char *str;               /* J_STR */
  char *key;               /* member name when inside an object */
  struct JVal *child, *next;
} JVal;

typedef struct { const char *p; int ok; } JParser;

static void j_free(JVal *v) {
  while (v) {
    JVal *n = v->next;
    j_free(v->child);
    free(v->str); free(v->key); free(v);
    v = n;
  }
}
static void j_ws(JParser *P) { while (*P->p && isspace((unsigned char)*P->p)) P->p++; }
static JVal *j_value(JParser *P);

static char *j_string(JParser *P) {
  if (*P->p != '"') { P->ok = 0; return NULL; }
  P->p++;
  size_t cap = 32, n = 0;
  char *out = (char *)malloc(cap);
  while (*P->p && *P->p != '"') {
    char c = *P->p++;
    if (c == '\\' && *P->p) {
      char e = *P->p++;
      switch (e) { case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'r': c = '\r'; break;
                   case 'b': c = '\b'; break; case 'f': c = '\f'; break;
                   case 'u': c = '?'; for (int k = 0; k < 4 && *P->p; k++) P->p++; break;
                   default: c = e; }
    }
    if (n + 2 > cap) { cap *= 2; out = (char *)realloc(out, cap); }
    out[n++] = c;
  }
  if (*P->p != '"') { P->ok = 0; free(out); return NULL; }
  P->p++;
  out[n] = 0;
  return out;
}

static JVal *j_new(JType t) { JVal *v = (JVal *)calloc(1, sizeof(JVal)); v->type = t; return v; }

static JVal *j_value(JParser *P) {
  j_ws(P);
  char c = *P->p;
  if (c == '{' || c == '[') {
    const int is_obj = c == '{';
    JVal *v = j_new(is_obj ? J_OBJ : J_ARR), **tail = &v->child;
    P->p++;
    j_ws(P);
    if (*P->p == (is_obj ? '}' : ']')) { P->p++; return v; }
    for (;;) {
      char *key = NULL;
      j_ws(P);
      if (is_obj) {
        key = j_string(P);
        j_ws(P);
        if (!P->ok || *P->p != ':') { P->ok = 0; free(key); break; }
        P->p++;
      }
      JVal *item = j_value(P);
      if (!item) { free(key); P->ok = 0; break; }
      item->key = key;
      *tail = item; tail = &item->next;
      j_ws(P);
      if (*P->p == ',') { P->p++; continue; }
      if (*P->p == (is_obj ? '}' : ']')) { P->p++; return v; }
      P->ok = 0;
      break;
    }
    j_free(v);
    return NULL;
  }
  if (c == '"') { char *s = j_string(P); if (!s) return NULL; JVal *v = j_new(J_STR); v->str = s; return v; }
  if (!strncmp(P->p, "true", 4)) { P->p += 4; JVal *v = j_new(J_BOOL); v->num = 1; return v; }
  if (!strncmp(P->p, "false", 5)) { P->p += 5; return j_new(J_BOOL); }
  if (!strncmp(P->p, "null", 4)) { P->p += 4; return j_new(J_NULL); }
  char *end;
  double d = strtod(P->p, &end);
  if (end == P->p) { P->ok = 0; return NULL; }
  P->p = end;
  JVal *v = j_new(J_NUM);
  v->num = d;
  return v;
}

static JVal *j_get(const JVal *obj, const char *key) {
  if (!obj || obj->type != J_OBJ) return NULL;
  for (JVal *c = obj->child; c; c = c->next) if (c->key && !strcmp(c->key, key)) return c;
  return NULL;
}
static JVal *j_at(const JVal *arr, int i) {
  if (!arr || arr->type != J_ARR) return NULL;
  JVal *c = arr->child;
  while (c && i-- > 0) c = c->next;
  return c;
}
/* scene_process_vec3_json (scene.c:595-604): three numbers, each narrowed double -> float */
static int j_vec3(const JVal *v, float *dest) {
  if (!v || v->type != J_ARR) return 0;
  for (int i = 0; i < 3; i++) {
    JVal *e = j_at(v, i);
    dest[i] = (e && e->type == J_NUM) ? (float)e->num : 0.0f;
  }
  return 1;
}

/* ------------------------------------------------------------------ transforms (cglm scalar order) */

static void euler_xyz_deg(const float *deg, float *r) {          /* glm_euler_xyz(glm_rad(deg)) upper 3x3 */
  const float pi = 3.14159265358979323846264338327950288f;
  float a[3] = {deg[0] * pi / 180.0f, deg[1] * pi / 180.0f, deg[2] * pi / 180.0f};
  float sx = sinf(a[0]), cx = cosf(a[0]), sy = sinf(a[1]), cy = cosf(a[1]), sz = sinf(a[2]), cz = cosf(a[2]);
  float czsx = cz * sx, cxcz = cx * cz, sysz = sy * sz;
  r[0] = cy * cz;             r[1] = czsx * sy + cx * sz;  r[2] = -cxcz * sy + sx * sz;
  r[3] = -cy * sz;            r[4] = cxcz - sx * sysz;     r[5] = czsx + cx * sysz;
  r[6] = sy;                  r[7] = -cy * sx;             r[8] = cx * cy;
}

/* local = T(position) R(rotation) S(scale); world = parent.world * local (glm_mat4_mul) or local */
static void node_build_transforms(struct SceneNode *n) {
  float r[9];
  euler_xyz_deg(n->rotation, r);
  float (*L)[4] = n->local_transform;
  for (int c = 0; c < 3; c++) {
    for (int k = 0; k < 3; k++) L[c][k] = r[c * 3 + k] * n->scale[c];
    L[c][3] = 0.0f * n->scale[c];
  }
  L[3][0] = n->position[0]; L[3][1] = n->position[1]; L[3][2] = n->position[2]; L[3][3] = 1.0f;
  if (!n->parent_node) { memcpy(n->world_transform, n->local_transform, 64); return; }
  float (*P)[4] = n->parent_node->world_transform, out[4][4];
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++)
      out[c][k] = P[0][k] * L[c][0] + P[1][k] * L[c][1] + P[2][k] * L[c][2] + P[3][k] * L[c][3];
  memcpy(n->world_transform, out, 64);
}

/* scene_node_update, scene.c:1051-1079 (+ the recursion over children that follows it) */
void crux_scene_node_update(struct SceneNode *node) {
  if (node->entity && node->entity->physics_body) {
    memcpy(node->position, node->entity->physics_body->position, 12);
    memcpy(node->entity->position, node->entity->physics_body->position, 12);
    memcpy(node->rotation, node->entity->physics_body->rotation, 12);
    memcpy(node->entity->rotation, node->entity->physics_body->rotation, 12);
  }
  node_build_transforms(node);
  for (unsigned int i = 0; i < node->num_children; i++) crux_scene_node_update(node->children[i]);
}

/* ------------------------------------------------------------------ scene construction */

static unsigned long long g_next_id = 1;

static struct SceneNode *new_node(CruxScene *s, struct SceneNode *parent, EntityType type, const float *pos, const float *rot,
                                  const float *scale) {
  struct Entity *e = (struct Entity *)calloc(1, sizeof(struct Entity));
  struct SceneNode *n = (struct SceneNode *)calloc(1, sizeof(struct SceneNode));
  if (!e || !n) { free(e); free(n); return NULL; }
  unsigned long long id = g_next_id++;             /* stands in for uuid_generate (scene.c:618) */
  memcpy(e->id, &id, sizeof(id));
  memcpy(e->id + 8, "cruxb200", 8);
  e->type = type;
  memcpy(e->position, pos, 12); memcpy(e->rotation, rot, 12); memcpy(e->scale, scale, 12);
  memcpy(n->entity_id, e->id, 16);
  n->ID = (unsigned int)id;
  memcpy(n->position, pos, 12); memcpy(n->rotation, rot, 12); memcpy(n->scale, scale, 12);
  n->entity = e;
  n->parent_node = parent;
  if (parent) {
    parent->children = (struct SceneNode **)realloc(parent->children, sizeof(struct SceneNode *) * (parent->num_children + 1));
    parent->children[parent->num_children++] = n;
  }
  node_build_transforms(n);
  if (s->n_nodes == s->cap_nodes) {
    s->cap_nodes = s->cap_nodes ? s->cap_nodes * 2 : 64;
    s->nodes = (struct SceneNode **)realloc(s->nodes, sizeof(struct SceneNode *) * s->cap_nodes);
    s->entities = (struct Entity **)realloc(s->entities, sizeof(struct Entity *) * s->cap_nodes);
  }
  s->nodes[s->n_nodes] = n;
  s->entities[s->n_nodes++] = e;
  return n;
}

/* collider object -> struct Collider (scene.c:692-766).  gen1: the pre-capsule enum (2 = plane). */
static int parse_collider(const JVal *cj, int gen1, struct Collider *out) {
  JVal *type = j_get(cj, "type"), *data = j_get(cj, "data");
  if (!type || type->type != J_NUM) { fprintf(stderr, "Error: failed to get type in collider object, either invalid or does not exist\n"); return 0; }
  if (!data) { fprintf(stderr, "Error: failed to get data in collider object, either invalid or does not exist\n"); return 0; }
  int t = (int)type->num;
  if (gen1 && t == 2) t = COLLIDER_PLANE;
  memset(out, 0, sizeof(*out));
  out->type = (ColliderType)t;
  JVal *radius;
  switch (t) {
    case COLLIDER_AABB:
      j_vec3(j_get(data, "center"), out->data.aabb.center);
      j_vec3(j_get(data, "extents"), out->data.aabb.extents);
      out->data.aabb.initialized = true;
      return 1;
    case COLLIDER_SPHERE:
      j_vec3(j_get(data, "center"), out->data.sphere.center);
      radius = j_get(data, "radius");
      if (!radius || radius->type != J_NUM) { fprintf(stderr, "Error: failed to get radius in collider object\n"); return 0; }
      out->data.sphere.radius = (float)radius->num;
      return 1;
    case COLLIDER_CAPSULE:
      j_vec3(j_get(data, "segment_A"), out->data.capsule.segment_A);
      j_vec3(j_get(data, "segment_B"), out->data.capsule.segment_B);
      radius = j_get(data, "radius");
      if (!radius || radius->type != J_NUM) { fprintf(stderr, "Error: failed to get radius in collider object\n"); return 0; }
      out->data.capsule.radius = (float)radius->num;
      return 1;
    case COLLIDER_PLANE: {
      j_vec3(j_get(data, "normal"), out->data.plane.normal);
      JVal *distance = j_get(data, "distance");
      if (!distance || distance->type != J_NUM) { fprintf(stderr, "Error: failed to get distance in collider object\n"); return 0; }
      out->data.plane.distance = (float)distance->num;
      return 1;
    }
    default:
      return 1;      /* upstream falls through its switch with an unset collider (scene.c:764-766) */
  }
}

static void attach_body(CruxScene *s, struct SceneNode *n, const JVal *node_json, const JVal *cj, int gen1, int gen1_dynamic) {
  struct Collider collider;
  if (!parse_collider(cj, gen1, &collider)) return;
  float restitution = 0.0f;
  bool dynamic = false;
  if (gen1) {
    dynamic = gen1_dynamic != 0;
    JVal *rj = j_get(cj, "restitution");
    restitution = dynamic ? (rj && rj->type == J_NUM ? (float)rj->num : 1.0f) : 0.0f;
  } else {
    JVal *dj = j_get(cj, "dynamic");
    if (!dj || dj->type != J_BOOL) { fprintf(stderr, "Error: failed to get dynamic bool in collider object, either invalid or does not exist\n"); return; }
    if (dj->num != 0) {
      JVal *rj = j_get(cj, "restitution");
      if (!rj || rj->type != J_NUM) { fprintf(stderr, "Error: failed to get restitution in collider object\n"); return; }
      restitution = (float)rj->num;
      dynamic = true;
    }
  }
  if (dynamic) j_vec3(j_get(node_json, "velocity"), n->entity->velocity);
  n->entity->physics_body = physics_add_body(s->world, n, n->entity, collider, restitution, dynamic);
}

/* scene_process_node_json, scene.c:606-908 (physics-relevant part), recursive over "children" */
static void load_node(CruxScene *s, const JVal *node_json, struct SceneNode *parent) {
  float pos[3] = {0, 0, 0}, rot[3] = {0, 0, 0}, scale[3] = {0, 0, 0};
  j_vec3(j_get(node_json, "position"), pos);
  j_vec3(j_get(node_json, "rotation"), rot);
  j_vec3(j_get(node_json, "scale"), scale);
  JVal *et = j_get(node_json, "entity_type");
  if (!et || et->type != J_NUM) {
    fprintf(stderr, "Error: failed to get entity type in scene node, either invalid or does not exist\n");
    return;
  }
  struct SceneNode *n = new_node(s, parent, (EntityType)(int)et->num, pos, rot, scale);
  if (!n) return;
  if (!parent) s->root = n;
  JVal *cj = j_get(node_json, "collider");
  if (!cj) { fprintf(stderr, "Error: failed to get collider object in scene node, either invalid or does not exist\n"); return; }
  if (cj->type != J_NULL) attach_body(s, n, node_json, cj, 0, 0);
  if (j_get(node_json, "components")) s->schema_generation = 3;     /* Gen-2 nodes simply have none */
  JVal *children = j_get(node_json, "children");
  if (children && children->type == J_ARR)
    for (JVal *c = children->child; c; c = c->next) load_node(s, c, n);
}

/* Gen-1: {"meshes": {"static": [...], "dynamic": [...]}} -- flat lists under an identity root */
static void load_gen1(CruxScene *s, const JVal *meshes) {
  const float zero[3] = {0, 0, 0}, one[3] = {1, 1, 1};
  s->root = new_node(s, NULL, ENTITY_GROUPING, zero, zero, one);
  const char *lists[2] = {"static", "dynamic"};
  for (int d = 0; d < 2; d++) {
    JVal *arr = j_get(meshes, lists[d]);
    if (!arr || arr->type != J_ARR) continue;
    for (JVal *m = arr->child; m; m = m->next) {
      float pos[3] = {0, 0, 0}, rot[3] = {0, 0, 0}, scale[3] = {1, 1, 1};
      j_vec3(j_get(m, "position"), pos); j_vec3(j_get(m, "rotation"), rot); j_vec3(j_get(m, "scale"), scale);
      struct SceneNode *n = new_node(s, s->root, ENTITY_WORLD, pos, rot, scale);
      JVal *cj = j_get(m, "collider");
      if (n && cj && cj->type == J_OBJ) attach_body(s, n, m, cj, 1, d);
    }
  }
}

/* scene_player_create + the player body of scene_load (scene.c:303-332, 965-1008, 1106-1153) */
static void add_player(CruxScene *s) {
  const float pos[3] = {0.0f, 0.0f, 2.0f}, rot[3] = {0.0f, 180.0f, 0.0f}, scale[3] = {1.0f, 1.0f, 1.0f};
  struct SceneNode *n = new_node(s, s->root, ENTITY_PLAYER, pos, rot, scale);
  if (!n) return;
  struct Collider c;
  memset(&c, 0, sizeof(c));
  c.type = COLLIDER_CAPSULE;
  c.data.capsule.segment_A[1] = 0.25f;
  c.data.capsule.segment_B[1] = 1.75f;
  c.data.capsule.radius = 0.25f;
  n->entity->physics_body = physics_add_player(s->world, n, n->entity, c);
}

CruxScene *crux_scene_load(const char *path, int with_player) {
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "Error: failed to read scene file %s\n", path); return NULL; }
  fseek(f, 0, SEEK_END);
  long size = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *text = (char *)malloc((size_t)size + 1);
  if (fread(text, 1, (size_t)size, f) != (size_t)size) { fclose(f); free(text); return NULL; }
  fclose(f);
  text[size] = 0;
  JParser P = {text, 1};
  JVal *doc = j_value(&P);
  free(text);
  if (!doc || !P.ok || doc->type != J_OBJ) {
    fprintf(stderr, "Error: failed to parse scene JSON %s\n", path);
    j_free(doc);
    return NULL;
  }
  CruxScene *s = (CruxScene *)calloc(1, sizeof(CruxScene));
  s->world = physics_world_create();
  JVal *nodes = j_get(doc, "nodes"), *meshes = j_get(doc, "meshes");
  if (nodes && nodes->type == J_OBJ) {
    s->schema_generation = 2;
    load_node(s, nodes, NULL);
  } else if (meshes && meshes->type == J_OBJ) {
    s->schema_generation = 1;
    load_gen1(s, meshes);
  } else {
    fprintf(stderr, "Error: scene %s has neither \"nodes\" nor \"meshes\"\n", path);
  }
  if (!s->root) { j_free(doc); crux_scene_free(s); return NULL; }
  if (with_player) { add_player(s); s->with_player = 1; }
  j_free(doc);
  return s;
}

void crux_scene_update(CruxScene *scene, float delta_time) {
  physics_step(scene->world, delta_time);          /* scene.c:401 */
  crux_scene_node_update(scene->root);             /* scene.c:408 */
}

void crux_scene_update_n(CruxScene *scene, float delta_time, int n_steps) {
#ifdef CRUX_USE_ENGINE_HEADERS
  for (int k = 0; k < n_steps; k++) crux_scene_update(scene, delta_time);   /* the reference has no device-resident loop */
#else
  physics_step_n(scene->world, delta_time, n_steps);
  crux_scene_node_update(scene->root);
#endif
}

void crux_scene_free(CruxScene *scene) {
  if (!scene) return;
#ifndef CRUX_USE_ENGINE_HEADERS
  physics_world_release_gpu(scene->world);
#endif
  for (int i = 0; i < scene->n_nodes; i++) {
    free(scene->nodes[i]->children);
    free(scene->nodes[i]);
    free(scene->entities[i]);
  }
  free(scene->nodes); free(scene->entities);
  if (scene->world) {                              /* scene.c:569-571 */
    free(scene->world->static_bodies); free(scene->world->dynamic_bodies); free(scene->world->player_bodies);
    free(scene->world);
  }
  free(scene);
}
