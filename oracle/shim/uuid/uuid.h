/*
 * oracle shim (test infrastructure): the two libuuid names the physics path touches
 * (/root/reference/src/physics/world.c:149-165 uses uuid_compare on Entity.id).
 */
#ifndef ORACLE_SHIM_UUID_H
#define ORACLE_SHIM_UUID_H
#include <string.h>
typedef unsigned char uuid_t[16];
static inline int uuid_compare(const uuid_t a, const uuid_t b) { return memcmp(a, b, 16); }
#endif
