/*
 * crux_headless.c -- run a Crux scene file through the physics-world API without a window:
 *
 *     crux_headless <scene.json> <steps> <out.bin> [--player] [--device-steps]
 *
 * Loads the scene (crux_scene.c), steps it `steps` frames at dt = 1/60 exactly like scene_update
 * (physics_step + scene_node_update per frame), records every contact event, and dumps
 *     int32 n_static, n_dynamic, n_player, n_events
 *     per body (static, dynamic, player order): position[3] velocity[3] at_rest(float) node_xform[16]
 *     per event: int32 type, a_index_tag, b_index_tag      (tag = low 4 bytes of the entity id)
 * The same source is linked twice: against libcrux_physics.so (this repo's GPU drop-in) and, by
 * oracle/Makefile, against the unmodified reference physics objects -- the two dumps must be
 * identical byte for byte (tests/test_gpu_dropin.py, config C1 of BASELINE.json).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crux_scene.h"

static int *g_ev = NULL;
static int g_nev = 0, g_cap = 0;

EventType get_event_type(EntityType a, EntityType b) {       /* event.c:14-20, 139-141 */
  if ((a == ENTITY_ITEM && b == ENTITY_PLAYER) || (a == ENTITY_PLAYER && b == ENTITY_ITEM)) return EVENT_PLAYER_ITEM_PICKUP;
  return EVENT_COLLISION;
}

void game_event_queue_enqueue(struct GameEvent ev) {           /* event.c:62-72, recording instead of queueing */
  if (g_nev == g_cap) { g_cap = g_cap ? g_cap * 2 : 4096; g_ev = (int *)realloc(g_ev, sizeof(int) * 3 * (size_t)g_cap); }
  int a, b;
  if (ev.type == EVENT_COLLISION) { memcpy(&a, ev.data.collision.entity_A_id, 4); memcpy(&b, ev.data.collision.entity_B_id, 4); }
  else { memcpy(&a, ev.data.item_pickup.item_entity_id, 4); memcpy(&b, ev.data.item_pickup.player_entity_id, 4); }
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

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s scene.json steps out.bin [--player] [--device-steps]\n", argv[0]); return 2; }
  int steps = atoi(argv[2]), player = 0, device_steps = 0;
  for (int i = 4; i < argc; i++) { if (!strcmp(argv[i], "--player")) player = 1; if (!strcmp(argv[i], "--device-steps")) device_steps = 1; }
  CruxScene *s = crux_scene_load(argv[1], player);
  if (!s) return 1;
  const float dt = 1.0f / 60.0f;
  if (device_steps) crux_scene_update_n(s, dt, steps);
  else for (int k = 0; k < steps; k++) crux_scene_update(s, dt);
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
