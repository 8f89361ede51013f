/*
 * oracle/oracle_abi.h -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * Flat, pointer-free records through which the tests drive BOTH checkers:
 *   - oracle/_ref/libcrux_ref.so   : the unmodified reference sources (the eight .c files of /root/reference/src/physics)
 *                                    compiled by oracle/Makefile, wrapped by oracle/ref_harness.c;
 *   - oracle/_build/libcrux_oracle.so : the plain-C restatement in oracle/crux_oracle.c.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything under oracle/.
 *
 * OraBody mirrors the fields of the reference's `struct PhysicsBody`
 * (/root/reference/include/physics/world.h:8-26) plus what the physics path reads through
 * `body->scene_node` (world_transform, parent_node->world_transform:
 * /root/reference/src/physics/distance.c:26-49,345-347) and `body->entity->type`
 * (/root/reference/src/physics/world.c:504-506).
 */
#ifndef ORACLE_ABI_H
#define ORACLE_ABI_H

#include <stdint.h>

enum { ORA_AABB = 0, ORA_SPHERE = 1, ORA_CAPSULE = 2, ORA_PLANE = 3 };   /* collider.h:18-24 */
enum { ORA_ENT_GROUPING = 0, ORA_ENT_WORLD = 1, ORA_ENT_ITEM = 2, ORA_ENT_PLAYER = 3 }; /* types.h:3-9 */
enum { ORA_CLASS_STATIC = 0, ORA_CLASS_DYNAMIC = 1, ORA_CLASS_PLAYER = 2 };

typedef struct OraBody {
  float position[3];
  float rotation[3];      /* degrees, as stored in PhysicsBody.rotation */
  float scale[3];
  float velocity[3];
  float restitution;
  float shape[7];         /* AABB: center[3],extents[3] | sphere: center[3],radius |
                             capsule: segment_A[3],segment_B[3],radius | plane: normal[3],distance */
  int32_t collider_type;  /* ORA_AABB .. ORA_PLANE */
  int32_t entity_type;    /* ORA_ENT_* */
  int32_t dynamic;
  int32_t at_rest;
  int32_t has_node;       /* body->scene_node != NULL */
  int32_t has_parent;     /* body->scene_node->parent_node != NULL */
  float node_xform[16];   /* scene_node->world_transform, cglm column-major mat4 */
  float parent_xform[16]; /* scene_node->parent_node->world_transform */
} OraBody;                /* 232 bytes */

/* One contact = one call the reference makes to game_event_queue_enqueue (world.c:259,352,457,540). */
typedef struct OraEvent {
  int32_t step;           /* 0-based step index within the ora_/ref_ world_step call sequence */
  int32_t event_type;     /* 0 = EVENT_COLLISION, 1 = EVENT_PLAYER_ITEM_PICKUP (event.h:8-11) */
  int32_t a_class, a_index;   /* body_A after the type-order swap (world.c:482-488) */
  int32_t b_class, b_index;
} OraEvent;

#endif /* ORACLE_ABI_H */
