/*
 * crux_headless.c -- run a Crux scene file through the physics-world API without a window:
 *
 *     crux_headless <scene.json> <steps> <out.bin> [--player] [--device-steps] [--script pickup|jump]
 *
 * Loads the scene (crux_scene.c), steps it `steps` frames at dt = 1/60 exactly like scene_update
 * (physics_step + scene_node_update per frame), records every contact event, and dumps
 *     int32 n_static, n_dynamic, n_player, n_events
 *     per body (static, dynamic, player order): position[3] velocity[3] at_rest(float) node_xform[16]
 *     per event: int32 type, a_index_tag, b_index_tag      (tag = low 4 bytes of the entity id)
 * --script plays the game layer's two write paths into the physics world between frames:
 *     pickup   three dynamic bodies become ITEM entities; a player/item contact is a TRIGGER whose event,
 *              processed after the frame like game_event_queue_process does (event.c:118-125), removes the
 *              item's body: scene_remove_entity -> physics_remove_body (scene.c:1203, world.c:147-172).  Two
 *              more removals are scripted (first and last dynamic body: both ends of the swap-and-pop).
 *     jump     the player is poked the way player.c does it: player_jump (at_rest = false, velocity[1] = 3;
 *              player.c:160-173), walking (position += forward * speed * dt, then scene_node_update of the
 *              player node; player.c:11-50) and turning (rotation[1] write; player.c:155-157).
 * The same source is linked twice: against libcrux_physics.so (this repo's GPU drop-in) and, by
 * oracle/Makefile, against the unmodified reference physics objects -- the two dumps must be
 * identical byte for byte (tests/test_gpu_dropin.py, config C1 of BASELINE.json).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "crux_scene.h"
#ifdef CRUX_HAVE_UPLOAD_STATS
#include "physics_gpu.h"      /* PgDebugDraw (the GPU drop-in build only) */
#endif

static int *g_ev = NULL;
static int g_nev = 0, g_cap = 0;
/* item pickups of the frame being stepped, handled after it (game_event_queue_process) */
static unsigned char g_pick[64][16];
static int g_npick = 0;

EventType get_event_type(EntityType a, EntityType b) {       /* event.c:14-20, 139-141 */
  if ((a == ENTITY_ITEM && b == ENTITY_PLAYER) || (a == ENTITY_PLAYER && b == ENTITY_ITEM)) return EVENT_PLAYER_ITEM_PICKUP;
  return EVENT_COLLISION;
}

void game_event_queue_enqueue(struct GameEvent ev) {           /* event.c:62-72, recording instead of queueing */
  if (g_nev == g_cap) { g_cap = g_cap ? g_cap * 2 : 4096; g_ev = (int *)realloc(g_ev, sizeof(int) * 3 * (size_t)g_cap); }
  int a, b;
  if (ev.type == EVENT_COLLISION) { memcpy(&a, ev.data.collision.entity_A_id, 4); memcpy(&b, ev.data.collision.entity_B_id, 4); }
  else {
    memcpy(&a, ev.data.item_pickup.item_entity_id, 4); memcpy(&b, ev.data.item_pickup.player_entity_id, 4);
    if (g_npick < 64) memcpy(g_pick[g_npick++], ev.data.item_pickup.item_entity_id, 16);
  }
  g_ev[3 * g_nev] = (int)ev.type; g_ev[3 * g_nev + 1] = a; g_ev[3 * g_nev + 2] = b;
  g_nev++;
}

static void dump_class(FILE *f, struct PhysicsBody *arr, unsigned int n) {
  for (unsigned int i = 0; i < n; i++) {
    float rec[23];
    memcpy(rec, arr[i].position, 12);
    memcpy(rec + 3, arr[i].velocity, 12);
    rec[6] = arr[i].at_rest ? 1.0f : 0.0f;
    if (arr[i].scene_node) memcpy(rec + 7, arr[i].scene_node->world_transform, 64); else memset(rec + 7, 0, 64);
    fwrite(rec, sizeof(float), 23, f);
  }
}

/* scene_remove_entity, physics part (scene.c:1196-1204): the body goes, the entity forgets it */
static int remove_entity_body(CruxScene *s, const unsigned char *entity_id) {
  for (int i = 0; i < s->n_nodes; i++) {
    struct Entity *e = s->entities[i];
    if (memcmp(e->id, entity_id, 16) != 0 || !e->physics_body) continue;
    physics_remove_body(s->world, e->physics_body);
    e->physics_body = NULL;
    return 1;
  }
  return 0;
}

static struct ItemComponent g_items[3];
static int g_removed = 0, g_jumps = 0;

static void script_before_frame(CruxScene *s, const char *script, int frame, float dt) {
  if (!script) return;
  if (!strcmp(script, "pickup")) {
    if (frame == 0) {
      const unsigned int pick[3] = {1, 4, 6};
      for (int k = 0; k < 3; k++) {
        if (pick[k] >= s->world->num_dynamic_bodies) continue;
        struct Entity *e = s->world->dynamic_bodies[pick[k]].entity;
        e->type = ENTITY_ITEM;
        g_items[k].id = 100 + k; g_items[k].count = 1 + k;
        e->item = &g_items[k];
      }
    }
    /* the player walks under the first item still in the world (a position write per frame, player.c:27-45) */
    if (s->world->num_player_bodies > 0) {
      struct PhysicsBody *body = &s->world->player_bodies[0];
      for (unsigned int i = 0; i < s->world->num_dynamic_bodies; i++) {
        const struct PhysicsBody *it = &s->world->dynamic_bodies[i];
        if (it->entity->type != ENTITY_ITEM) continue;
        body->position[0] = it->position[0];
        body->position[2] = it->position[2];
        if (body->scene_node) crux_scene_node_update(body->scene_node);
        break;
      }
    }
    if (frame == 120 && s->world->num_dynamic_bodies > 0)
      g_removed += remove_entity_body(s, s->world->dynamic_bodies[0].entity->id);
    if (frame == 240 && s->world->num_dynamic_bodies > 0)
      g_removed += remove_entity_body(s, s->world->dynamic_bodies[s->world->num_dynamic_bodies - 1].entity->id);
  } else if (!strcmp(script, "jump") && s->world->num_player_bodies > 0) {
    struct PhysicsBody *body = &s->world->player_bodies[0];
    if (frame == 30 || frame == 120 || frame == 121 || frame == 300) {     /* player_jump */
      body->at_rest = false;
      body->velocity[1] = 3.0f;
      g_jumps++;
    }
    if (frame >= 60 && frame < 90) {                                        /* CAMERA_FORWARD, speed 2.5, front = (0.6, 0, -0.8) */
      const float velocity = (float)(2.5f * dt);
      body->position[0] += 0.6f * velocity;
      body->position[2] += -0.8f * velocity;
      if (body->scene_node) crux_scene_node_update(body->scene_node);
    }
    if (frame == 200) { body->rotation[0] = 0.0f; body->rotation[1] = -35.0f + 90.0f; body->rotation[2] = 0.0f; }
  }
}

static void script_after_frame(CruxScene *s, const char *script) {
  if (script && !strcmp(script, "pickup"))
    for (int k = 0; k < g_npick; k++) g_removed += remove_entity_body(s, g_pick[k]);
  g_npick = 0;
}

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s scene.json steps out.bin [--player] [--device-steps] [--script pickup|jump]\n", argv[0]); return 2; }
  int steps = atoi(argv[2]), player = 0, device_steps = 0;
  const char *script = NULL, *debug_path = NULL;
  for (int i = 4; i < argc; i++) {
    if (!strcmp(argv[i], "--player")) player = 1;
    if (!strcmp(argv[i], "--device-steps")) device_steps = 1;
    if (!strcmp(argv[i], "--script") && i + 1 < argc) script = argv[++i];
    if (!strcmp(argv[i], "--debug-geometry") && i + 1 < argc) debug_path = argv[++i];
  }
  CruxScene *s = crux_scene_load(argv[1], player);
  if (!s) return 1;
  const float dt = 1.0f / 60.0f;
  struct timespec t_begin, t_end;
  clock_gettime(CLOCK_MONOTONIC, &t_begin);
  if (device_steps && !script) crux_scene_update_n(s, dt, steps);
  else for (int k = 0; k < steps; k++) {
    if (k == 100) clock_gettime(CLOCK_MONOTONIC, &t_begin);      /* (device and library start-up stay out of the figure) */
    script_before_frame(s, script, k, dt);
#ifdef CRUX_HAVE_UPLOAD_STATS
    const unsigned long long up0 = physics_world_uploaded_records(s->world);
#endif
    crux_scene_update(s, dt);
#ifdef CRUX_HAVE_UPLOAD_STATS
    if (script) printf("frame %d uploaded %llu\n", k, physics_world_uploaded_records(s->world) - up0);
#endif
    script_after_frame(s, script);
  }
  clock_gettime(CLOCK_MONOTONIC, &t_end);
  if (script) printf("script %s: %d bodies removed, %d jumps\n", script, g_removed, g_jumps);
  {
    const int timed = (device_steps && !script) ? steps : (steps > 100 ? steps - 100 : steps);
    printf("frame loop: %.1f us per frame (%d frames timed)\n",
           ((t_end.tv_sec - t_begin.tv_sec) * 1e6 + (t_end.tv_nsec - t_begin.tv_nsec) * 1e-3) / (timed > 0 ? timed : 1), timed);
  }
#ifdef CRUX_HAVE_UPLOAD_STATS
  if (debug_path) {
    /* what the debug renderer would draw now (debug_renderer.c:11-213), after one more host-side poke of the player
     * that the device has not seen yet: physics_debug_geometry has to send it up before the kernel runs */
    if (s->world->num_player_bodies > 0) {
      struct PhysicsBody *body = &s->world->player_bodies[0];
      body->position[0] += 0.25f;
      if (body->scene_node) crux_scene_node_update(body->scene_node);
    }
    const int cap = (int)(s->world->num_static_bodies + s->world->num_dynamic_bodies + s->world->num_player_bodies);
    PgDebugDraw *recs = (PgDebugDraw *)calloc((size_t)(cap > 0 ? cap : 1), sizeof(PgDebugDraw));
    const int n = physics_debug_geometry(s->world, recs, 0, cap);
    FILE *df = fopen(debug_path, "wb");
    if (!df || n < 0) return 1;
    fwrite(recs, sizeof(PgDebugDraw), (size_t)n, df);
    fclose(df);
    free(recs);
    printf("debug geometry: %d records\n", n);
  }
#else
  (void)debug_path;
#endif
  FILE *f = fopen(argv[3], "wb");
  if (!f) return 1;
  int hdr[4] = {(int)s->world->num_static_bodies, (int)s->world->num_dynamic_bodies, (int)s->world->num_player_bodies, g_nev};
  fwrite(hdr, sizeof(int), 4, f);
  dump_class(f, s->world->static_bodies, s->world->num_static_bodies);
  dump_class(f, s->world->dynamic_bodies, s->world->num_dynamic_bodies);
  dump_class(f, s->world->player_bodies, s->world->num_player_bodies);
  fwrite(g_ev, sizeof(int), 3 * (size_t)g_nev, f);
  fclose(f);
  printf("%s: generation %d, %d static %d dynamic %d player bodies, %d steps, %d events\n", argv[1], s->schema_generation, hdr[0], hdr[1], hdr[2],
         steps, g_nev);
#ifdef CRUX_HAVE_UPLOAD_STATS
  /* the GPU drop-in only: how many body records travelled host -> device (dirty ranges, crux_world.c) */
  printf("uploaded body records: %llu\n", physics_world_uploaded_records(s->world));
#endif
  crux_scene_free(s);
  return 0;
}
