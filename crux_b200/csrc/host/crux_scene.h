/*
 * crux_scene.h -- headless scene loading and per-frame update for the physics path
 * (SURVEY.md section 8f rank 1): what /root/reference/src/scene.c does between a scenes/*.json file
 * and physics_add_body, and between physics_step and the next frame, without GL / audio / assimp.
 */
#ifndef CRUX_SCENE_H
#define CRUX_SCENE_H

#include "crux_physics.h"

typedef struct CruxScene {
  struct PhysicsWorld *world;
  struct SceneNode *root;
  struct SceneNode **nodes;      /* every node, creation (pre-order) order */
  struct Entity **entities;
  int n_nodes, cap_nodes;
  int schema_generation;         /* 1 flat "meshes", 2 "nodes" without components, 3 current */
  int with_player;
} CruxScene;

/* Load scenes/<name>.json.  add_player=1 also creates the player entity, node and capsule body the
 * way scene_load does (scene.c:303-332).  Returns NULL (and prints why) on unreadable input. */
CruxScene *crux_scene_load(const char *path, int add_player);
/* scene_update without the game layer: physics_step, then scene_node_update of the whole graph
 * (scene.c:395-408, 1051-1079).  Events are delivered through game_event_queue_enqueue. */
void crux_scene_update(CruxScene *scene, float delta_time);
/* The same n frames with the bodies resident on the device (physics_step_n). */
void crux_scene_update_n(CruxScene *scene, float delta_time, int n_steps);
void crux_scene_free(CruxScene *scene);
/* scene_node_update for one node and, recursively, its children */
void crux_scene_node_update(struct SceneNode *node);

#endif
