/*
 * oracle/shim/cglm/cglm.h -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * Stand-in for the cglm maths library, which the reference links as an un-vendored, un-pinned
 * system package (`-lcglm`, /root/reference/makefile:9; README.MD:11,21).  Only the 30-odd entry
 * points that /root/reference/src/physics/*.c, src/scene.c:666-689,1051-1079 and
 * test/physics/test_aabb.c call are provided.  Each one restates cglm's published *scalar*
 * definition (cglm 0.9.x, headers vec3.h / mat3.h / mat4.h / affine.h / affine-mat.h / euler.h /
 * util.h): plain IEEE binary32 operations, products summed left to right, no FMA (the build passes
 * -ffp-contract=off).  Upstream x86 builds may take an SSE2 path for mat4*vec4 and mat4*mat4 whose
 * summation order differs; the physics path only feeds those with (0,0,0,1) or identity parents,
 * where the order is immaterial, except for capsule end points (noted in DESIGN.md, "parity
 * unpinned" there).
 */
#ifndef ORACLE_SHIM_CGLM_H
#define ORACLE_SHIM_CGLM_H

#include <float.h>
#include <math.h>
#include <string.h>

#define CGLM_INLINE static inline __attribute__((always_inline))

typedef float vec2[2];
typedef float vec3[3];
typedef __attribute__((aligned(16))) float vec4[4];
typedef vec3 mat3[3];
typedef __attribute__((aligned(16))) vec4 mat4[4];

#define GLM_PIf 3.14159265358979323846264338327950288f
#define GLM_MAT4_IDENTITY_INIT {{1.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 1.0f, 0.0f, 0.0f}, \
                                {0.0f, 0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 1.0f}}

/* ---- util.h ---- */
CGLM_INLINE float glm_rad(float deg) { return deg * GLM_PIf / 180.0f; }
CGLM_INLINE float glm_deg(float rad) { return rad * 180.0f / GLM_PIf; }
CGLM_INLINE float glm_min(float a, float b) { if (a < b) return a; return b; }
CGLM_INLINE float glm_max(float a, float b) { if (a > b) return a; return b; }
CGLM_INLINE float glm_clamp(float val, float minVal, float maxVal) {
  return glm_min(glm_max(val, minVal), maxVal);
}

/* ---- vec3.h ---- */
CGLM_INLINE void glm_vec3(vec4 v4, vec3 dest) { dest[0] = v4[0]; dest[1] = v4[1]; dest[2] = v4[2]; }
CGLM_INLINE void glm_vec3_copy(vec3 a, vec3 dest) { dest[0] = a[0]; dest[1] = a[1]; dest[2] = a[2]; }
CGLM_INLINE void glm_vec3_zero(vec3 v) { v[0] = v[1] = v[2] = 0.0f; }
CGLM_INLINE float glm_vec3_dot(vec3 a, vec3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
CGLM_INLINE float glm_dot(vec3 a, vec3 b) { return glm_vec3_dot(a, b); }
CGLM_INLINE float glm_vec3_norm2(vec3 v) { return glm_vec3_dot(v, v); }
CGLM_INLINE float glm_vec3_norm(vec3 v) { return sqrtf(glm_vec3_norm2(v)); }
CGLM_INLINE void glm_vec3_add(vec3 a, vec3 b, vec3 dest) {
  dest[0] = a[0] + b[0]; dest[1] = a[1] + b[1]; dest[2] = a[2] + b[2];
}
CGLM_INLINE void glm_vec3_adds(vec3 v, float s, vec3 dest) {
  dest[0] = v[0] + s; dest[1] = v[1] + s; dest[2] = v[2] + s;
}
CGLM_INLINE void glm_vec3_sub(vec3 a, vec3 b, vec3 dest) {
  dest[0] = a[0] - b[0]; dest[1] = a[1] - b[1]; dest[2] = a[2] - b[2];
}
CGLM_INLINE void glm_vec3_mul(vec3 a, vec3 b, vec3 dest) {
  dest[0] = a[0] * b[0]; dest[1] = a[1] * b[1]; dest[2] = a[2] * b[2];
}
CGLM_INLINE void glm_vec3_scale(vec3 v, float s, vec3 dest) {
  dest[0] = v[0] * s; dest[1] = v[1] * s; dest[2] = v[2] * s;
}
CGLM_INLINE void glm_vec3_muladds(vec3 a, float s, vec3 dest) {
  dest[0] += a[0] * s; dest[1] += a[1] * s; dest[2] += a[2] * s;
}
CGLM_INLINE void glm_vec3_mulsubs(vec3 a, float s, vec3 dest) {
  dest[0] -= a[0] * s; dest[1] -= a[1] * s; dest[2] -= a[2] * s;
}
CGLM_INLINE void glm_vec3_negate(vec3 v) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }
CGLM_INLINE void glm_vec3_normalize(vec3 v) {
  float norm = glm_vec3_norm(v);
  if (norm < FLT_EPSILON) { glm_vec3_zero(v); return; }
  glm_vec3_scale(v, 1.0f / norm, v);
}
CGLM_INLINE void glm_vec3_cross(vec3 a, vec3 b, vec3 dest) {
  vec3 c;
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
  glm_vec3_copy(c, dest);
}
/* from + t * (to - from), t not clamped (glm_vec3_lerpc is the clamping variant) */
CGLM_INLINE void glm_vec3_lerp(vec3 from, vec3 to, float t, vec3 dest) {
  vec3 s, v;
  s[0] = s[1] = s[2] = t;
  glm_vec3_sub(to, from, v);
  glm_vec3_mul(s, v, v);
  glm_vec3_add(from, v, dest);
}

/* ---- vec4 helpers used by the affine routines ---- */
CGLM_INLINE void glm_vec4(vec3 v3, float last, vec4 dest) {
  dest[0] = v3[0]; dest[1] = v3[1]; dest[2] = v3[2]; dest[3] = last;
}
CGLM_INLINE void glm_vec4_copy(vec4 v, vec4 dest) {
  dest[0] = v[0]; dest[1] = v[1]; dest[2] = v[2]; dest[3] = v[3];
}
CGLM_INLINE void glm_vec4_scale(vec4 v, float s, vec4 dest) {
  dest[0] = v[0] * s; dest[1] = v[1] * s; dest[2] = v[2] * s; dest[3] = v[3] * s;
}
CGLM_INLINE void glm_vec4_add(vec4 a, vec4 b, vec4 dest) {
  dest[0] = a[0] + b[0]; dest[1] = a[1] + b[1]; dest[2] = a[2] + b[2]; dest[3] = a[3] + b[3];
}

/* ---- mat3.h ---- */
CGLM_INLINE void glm_mat3_scale(mat3 m, float s) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] *= s;
}
CGLM_INLINE void glm_mat3_mulv(mat3 m, vec3 v, vec3 dest) {
  vec3 res;
  res[0] = m[0][0] * v[0] + m[1][0] * v[1] + m[2][0] * v[2];
  res[1] = m[0][1] * v[0] + m[1][1] * v[1] + m[2][1] * v[2];
  res[2] = m[0][2] * v[0] + m[1][2] * v[1] + m[2][2] * v[2];
  glm_vec3_copy(res, dest);
}

/* ---- mat4.h ---- */
CGLM_INLINE void glm_mat4_copy(mat4 mat, mat4 dest) { memcpy(dest, mat, sizeof(mat4)); }
CGLM_INLINE void glm_mat4_identity(mat4 mat) {
  mat4 t = GLM_MAT4_IDENTITY_INIT;
  glm_mat4_copy(t, mat);
}
CGLM_INLINE void glm_mat4_pick3(mat4 mat, mat3 dest) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dest[i][j] = mat[i][j];
}
CGLM_INLINE void glm_mat4_mulv(mat4 m, vec4 v, vec4 dest) {
  vec4 res;
  res[0] = m[0][0] * v[0] + m[1][0] * v[1] + m[2][0] * v[2] + m[3][0] * v[3];
  res[1] = m[0][1] * v[0] + m[1][1] * v[1] + m[2][1] * v[2] + m[3][1] * v[3];
  res[2] = m[0][2] * v[0] + m[1][2] * v[1] + m[2][2] * v[2] + m[3][2] * v[3];
  res[3] = m[0][3] * v[0] + m[1][3] * v[1] + m[2][3] * v[2] + m[3][3] * v[3];
  glm_vec4_copy(res, dest);
}
CGLM_INLINE void glm_mat4_mulv3(mat4 m, vec3 v, float last, vec3 dest) {
  vec4 res;
  glm_vec4(v, last, res);
  glm_mat4_mulv(m, res, res);
  glm_vec3(res, dest);
}
CGLM_INLINE void glm_mat4_mul(mat4 m1, mat4 m2, mat4 dest) {
  mat4 r;
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++)
      r[c][k] = m1[0][k] * m2[c][0] + m1[1][k] * m2[c][1] + m1[2][k] * m2[c][2] + m1[3][k] * m2[c][3];
  glm_mat4_copy(r, dest);
}

/* ---- affine-mat.h: product of two affine matrices (bottom row 0 0 0 1 assumed) ---- */
CGLM_INLINE void glm_mul(mat4 m1, mat4 m2, mat4 dest) {
  mat4 r;
  for (int c = 0; c < 3; c++) {
    for (int k = 0; k < 3; k++)
      r[c][k] = m1[0][k] * m2[c][0] + m1[1][k] * m2[c][1] + m1[2][k] * m2[c][2];
    r[c][3] = 0.0f;
  }
  for (int k = 0; k < 3; k++)
    r[3][k] = m1[0][k] * m2[3][0] + m1[1][k] * m2[3][1] + m1[2][k] * m2[3][2] + m1[3][k];
  r[3][3] = 1.0f;
  glm_mat4_copy(r, dest);
}
/* rotation-only right factor: columns 0..2 as glm_mul, column 3 copied from m1 */
CGLM_INLINE void glm_mul_rot(mat4 m1, mat4 m2, mat4 dest) {
  mat4 r;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 4; k++)
      r[c][k] = m1[0][k] * m2[c][0] + m1[1][k] * m2[c][1] + m1[2][k] * m2[c][2];
  glm_vec4_copy(m1[3], r[3]);
  glm_mat4_copy(r, dest);
}

/* ---- affine.h ---- */
CGLM_INLINE void glm_translate(mat4 m, vec3 v) {
  vec4 v1, v2, v3;
  glm_vec4_scale(m[0], v[0], v1);
  glm_vec4_scale(m[1], v[1], v2);
  glm_vec4_scale(m[2], v[2], v3);
  glm_vec4_add(v1, m[3], m[3]);
  glm_vec4_add(v2, m[3], m[3]);
  glm_vec4_add(v3, m[3], m[3]);
}
CGLM_INLINE void glm_translate_make(mat4 m, vec3 v) {
  glm_mat4_identity(m);
  m[3][0] = v[0]; m[3][1] = v[1]; m[3][2] = v[2];
}
CGLM_INLINE void glm_scale(mat4 m, vec3 v) {
  glm_vec4_scale(m[0], v[0], m[0]);
  glm_vec4_scale(m[1], v[1], m[1]);
  glm_vec4_scale(m[2], v[2], m[2]);
}
CGLM_INLINE void glm_scale_make(mat4 m, vec3 v) {
  glm_mat4_identity(m);
  m[0][0] = v[0]; m[1][1] = v[1]; m[2][2] = v[2];
}
CGLM_INLINE void glm_rotate_x(mat4 m, float angle, mat4 dest) {
  mat4 t = GLM_MAT4_IDENTITY_INIT;
  float c = cosf(angle), s = sinf(angle);
  t[1][1] = c; t[1][2] = s; t[2][1] = -s; t[2][2] = c;
  glm_mul_rot(m, t, dest);
}
CGLM_INLINE void glm_rotate_y(mat4 m, float angle, mat4 dest) {
  mat4 t = GLM_MAT4_IDENTITY_INIT;
  float c = cosf(angle), s = sinf(angle);
  t[0][0] = c; t[0][2] = -s; t[2][0] = s; t[2][2] = c;
  glm_mul_rot(m, t, dest);
}
CGLM_INLINE void glm_rotate_z(mat4 m, float angle, mat4 dest) {
  mat4 t = GLM_MAT4_IDENTITY_INIT;
  float c = cosf(angle), s = sinf(angle);
  t[0][0] = c; t[0][1] = s; t[1][0] = -s; t[1][1] = c;
  glm_mul_rot(m, t, dest);
}
/* rotation by `angle` about `axis` (affine.h glm_rotate_make: axis normalised, Rodrigues form built column by column) */
CGLM_INLINE void glm_vec3_normalize_to(vec3 v, vec3 dest) {
  float norm = glm_vec3_norm(v);
  if (norm < FLT_EPSILON) { glm_vec3_zero(dest); return; }
  glm_vec3_scale(v, 1.0f / norm, dest);
}
CGLM_INLINE void glm_rotate_make(mat4 m, float angle, vec3 axis) {
  vec3 axisn, v, vs;
  float c = cosf(angle);
  glm_vec3_normalize_to(axis, axisn);
  glm_vec3_scale(axisn, 1.0f - c, v);
  glm_vec3_scale(axisn, sinf(angle), vs);
  glm_vec3_scale(axisn, v[0], m[0]);
  glm_vec3_scale(axisn, v[1], m[1]);
  glm_vec3_scale(axisn, v[2], m[2]);
  m[0][0] += c;       m[1][0] -= vs[2];   m[2][0] += vs[1];
  m[0][1] += vs[2];   m[1][1] += c;       m[2][1] -= vs[0];
  m[0][2] -= vs[1];   m[1][2] += vs[0];   m[2][2] += c;
  m[0][3] = m[1][3] = m[2][3] = m[3][0] = m[3][1] = m[3][2] = 0.0f;
  m[3][3] = 1.0f;
}
CGLM_INLINE void glm_decompose_scalev(mat4 m, vec3 s) {
  s[0] = glm_vec3_norm(m[0]);
  s[1] = glm_vec3_norm(m[1]);
  s[2] = glm_vec3_norm(m[2]);
}

/* ---- euler.h: extrinsic x-y-z, angles in radians ---- */
CGLM_INLINE void glm_euler_xyz(vec3 angles, mat4 dest) {
  float cx, cy, cz, sx, sy, sz, czsx, cxcz, sysz;
  sx = sinf(angles[0]); cx = cosf(angles[0]);
  sy = sinf(angles[1]); cy = cosf(angles[1]);
  sz = sinf(angles[2]); cz = cosf(angles[2]);
  czsx = cz * sx;
  cxcz = cx * cz;
  sysz = sy * sz;
  dest[0][0] =  cy * cz;
  dest[0][1] =  czsx * sy + cx * sz;
  dest[0][2] = -cxcz * sy + sx * sz;
  dest[1][0] = -cy * sz;
  dest[1][1] =  cxcz - sx * sysz;
  dest[1][2] =  czsx + cx * sysz;
  dest[2][0] =  sy;
  dest[2][1] = -cy * sx;
  dest[2][2] =  cx * cy;
  dest[0][3] = 0.0f; dest[1][3] = 0.0f; dest[2][3] = 0.0f;
  dest[3][0] = 0.0f; dest[3][1] = 0.0f; dest[3][2] = 0.0f; dest[3][3] = 1.0f;
}

#endif /* ORACLE_SHIM_CGLM_H */
