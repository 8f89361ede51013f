/* oracle shim (test infrastructure): only the struct name that include/utils.h:4,21-22 mentions. */
#ifndef ORACLE_SHIM_ASSIMP_MATRIX4X4_H
#define ORACLE_SHIM_ASSIMP_MATRIX4X4_H
struct aiMatrix4x4 { float a1, a2, a3, a4, b1, b2, b3, b4, c1, c2, c3, c4, d1, d2, d3, d4; };
#endif
