// pg_debug.cu -- debug-geometry interop.
//
// The reference's physics_debug_render (src/physics/debug_renderer.c:11-213) walks the three body arrays on the host
// every frame and, per body, either recomputes a world-space box and uploads its 8 corners into the body's VBO
// (physics_debug_AABB_render, :655-696) or builds a model matrix for a unit mesh (spheres :184-197, capsules :45-60).
// With the state resident on the GPU those reads would be a full read-back per frame; instead one kernel turns the
// resident records into the same numbers, written straight into a buffer the GL host has shared with CUDA.
// Same arithmetic as the reference (cglm scalar order, no FMA: the translation unit is built with -fmad=false), so the
// records are bit-identical to what the unmodified debug_renderer.c hands to glBufferSubData / shader_set_mat4
// (tests/test_gpu_debug.py, against a recording GL stub linked with the reference's own file).
//
// The init-time meshes (:311-337, :371-432, :469-616) and the model matrix of a capsule without a scene node use
// sinf/cosf/acosf and are built on the host with libm, like every other trigonometric quantity of this library.
#include "pg_common.cuh"
#include "pg_math.cuh"

#include <cfloat>
#include <cmath>
#include <vector>

namespace {

using pgm::V3;
using pgm::v3;

// glm_mat4_identity; glm_translate(m, t); glm_scale(m, s)  (affine.h), on a column-major float[16].
// Written out as the reference's calls do it so that signed zeros and non-finite inputs come out the same.
__device__ void model_translate_scale(float *m, const float *t, const float *s) {
#pragma unroll
  for (int k = 0; k < 16; k++) m[k] = (k % 5 == 0) ? 1.0f : 0.0f;
  // glm_translate: m[3] = m[0]*t0 + m[3]; m[3] = m[1]*t1 + m[3]; m[3] = m[2]*t2 + m[3]  (three separate vec4 adds)
  float v1[4], v2[4], v3_[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { v1[k] = m[0 + k] * t[0]; v2[k] = m[4 + k] * t[1]; v3_[k] = m[8 + k] * t[2]; }
#pragma unroll
  for (int k = 0; k < 4; k++) { m[12 + k] = v1[k] + m[12 + k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) { m[12 + k] = v2[k] + m[12 + k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) { m[12 + k] = v3_[k] + m[12 + k]; }
  // glm_scale: columns 0..2 times s0..s2, all four components
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 4; k++) m[c * 4 + k] = m[c * 4 + k] * s[c];
}

// the 8 corners in the order of debug_renderer.c:677-688 (matching the 24 line indices of :321-337)
__device__ void box_corners(const pgm::Box &b, float *out) {
  const float sx[8] = {+1, +1, +1, +1, -1, -1, -1, -1};
  const float sy[8] = {+1, +1, -1, -1, -1, -1, +1, +1};
  const float sz[8] = {+1, -1, -1, +1, -1, +1, +1, -1};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    out[3 * k + 0] = sx[k] > 0.0f ? b.c.x + b.e.x : b.c.x - b.e.x;
    out[3 * k + 1] = sy[k] > 0.0f ? b.c.y + b.e.y : b.c.y - b.e.y;
    out[3 * k + 2] = sz[k] > 0.0f ? b.c.z + b.e.z : b.c.z - b.e.z;
  }
}

// One thread per body; record order = the reference's draw order: players, statics, dynamics.
__global__ void k_debug_geometry(WorldsView v, int world, int sphere_mode, PgDebugDraw *out) {
  const int np = v.counts[world * 3 + 2], ns = v.counts[world * 3 + 0], nd = v.counts[world * 3 + 1];
  const int total = np + ns + nd;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
    int cls, idx;
    if (k < np) { cls = PG_CLASS_PLAYER; idx = k; }
    else if (k < np + ns) { cls = PG_CLASS_STATIC; idx = k - np; }
    else { cls = PG_CLASS_DYNAMIC; idx = k - np - ns; }
    PgDebugDraw d;
    d.kind = PG_DEBUG_NONE; d.cls = cls; d.index = idx; d.n_indices = 0;
#pragma unroll
    for (int q = 0; q < 24; q++) d.data[q] = 0.0f;
    if (sphere_mode) {
      // sphere batches: statics are planes (never drawn), dynamics are spheres with centre 0 and unit scale
      if (cls == PG_CLASS_DYNAMIC) {
        const float4 p = v.pos_rest[(size_t)world * v.cap[1] + idx];
        const float t[3] = {p.x + 0.0f, p.y + 0.0f, p.z + 0.0f};       // glm_vec3_add(position, centre = 0), :186
        const float s[3] = {1.0f, 1.0f, 1.0f};
        model_translate_scale(d.data, t, s);
        d.kind = PG_DEBUG_SPHERE_MODEL; d.n_indices = 2280;
      }
    } else {
      const PgBody &b = v.bodies[cls][(size_t)world * v.cap[cls] + idx];
      switch (b.collider_type) {
        case PG_AABB:
          if (b.has_node) {
            // physics_debug_AABB_render(box, ctx, scene_node->world_transform), :662-676
            pgm::NodeFrame f; pgm::node_decompose(b.node_xform, f);
            const pgm::Box wb = pgm::aabb_update(b.shape, f.rot, f.pos, f.scl);
            box_corners(wb, d.data);
            d.kind = PG_DEBUG_AABB_LINES; d.n_indices = 24;
          }
          break;
        case PG_SPHERE: {
          // players and statics translate by the position (:39, :100); dynamics by position + centre (:184-187)
          float t[3] = {b.position[0], b.position[1], b.position[2]};
          if (cls == PG_CLASS_DYNAMIC) { t[0] = b.position[0] + b.shape[0]; t[1] = b.position[1] + b.shape[1]; t[2] = b.position[2] + b.shape[2]; }
          model_translate_scale(d.data, t, b.scale);
          d.kind = PG_DEBUG_SPHERE_MODEL; d.n_indices = 2280;
          break;
        }
        case PG_CAPSULE:
          // players with a scene node (:49-51) and every dynamic capsule (:206) draw with the node's world transform
          if ((cls == PG_CLASS_PLAYER || cls == PG_CLASS_DYNAMIC) && b.has_node) {
#pragma unroll
            for (int q = 0; q < 16; q++) d.data[q] = b.node_xform[q];
            d.kind = PG_DEBUG_CAPSULE_MODEL;
          } else if (cls != PG_CLASS_DYNAMIC) {
            d.kind = PG_DEBUG_CAPSULE_HOST;        // T R_y R_x R_z S from libm on the host: pg_debug_capsule_model
          }
          d.n_indices = d.kind != PG_DEBUG_NONE ? 4920 : 0;
          break;
        default:
          break;                                   // planes are not drawn (:61, :119, :208)
      }
    }
    out[k] = d;
  }
}

// ---- host: cglm's scalar definitions of the handful of calls the init functions make ----
inline float h_norm(const float *v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
inline void h_normalize(float *v) {                               // glm_vec3_normalize
  const float n = h_norm(v);
  if (n < FLT_EPSILON) { v[0] = v[1] = v[2] = 0.0f; return; }
  const float s = 1.0f / n;
  v[0] *= s; v[1] *= s; v[2] *= s;
}
// glm_rotate_make (affine.h): m = rotation by `angle` about `axis` (normalised here), column-major float[16]
void h_rotate_make(float *m, float angle, const float *axis) {
  float axisn[3] = {axis[0], axis[1], axis[2]}, v[3], vs[3];
  const float c = cosf(angle);
  h_normalize(axisn);
  const float omc = 1.0f - c, s = sinf(angle);
  for (int k = 0; k < 3; k++) { v[k] = axisn[k] * omc; vs[k] = axisn[k] * s; }
  for (int col = 0; col < 3; col++)
    for (int k = 0; k < 3; k++) m[col * 4 + k] = axisn[k] * v[col];
  m[0] += c;      m[4] -= vs[2];  m[8] += vs[1];
  m[1] += vs[2];  m[5] += c;      m[9] -= vs[0];
  m[2] -= vs[1];  m[6] += vs[0];  m[10] += c;
  m[3] = m[7] = m[11] = m[12] = m[13] = m[14] = 0.0f;
  m[15] = 1.0f;
}
// glm_mul_rot(m, t, m): columns 0..2 of m times the rotation t (all four rows), column 3 kept
void h_mul_rot(float *m, const float *t) {
  float r[16];
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 4; k++) r[c * 4 + k] = m[0 + k] * t[c * 4 + 0] + m[4 + k] * t[c * 4 + 1] + m[8 + k] * t[c * 4 + 2];
  for (int k = 0; k < 4; k++) r[12 + k] = m[12 + k];
  memcpy(m, r, sizeof(r));
}
inline float h_rad(float deg) { return deg * 3.14159265358979323846264338327950288f / 180.0f; }   // glm_rad

}  // namespace

static int pg_launch_debug_geometry(PgWorlds *w, int world, int total, PgDebugDraw *d_out) {
  if (total <= 0) return PG_OK;
  const int threads = 128;
  k_debug_geometry<<<(total + threads - 1) / threads, threads, 0, w->stream>>>(w->v, world, w->mode == PG_MODE_SPHERES ? 1 : 0, d_out);
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}

extern "C" {

int pg_worlds_debug_geometry(PgWorlds *w, int world, void *dst, int dst_is_device, int max_records) {
  if (!w || world < 0 || world >= w->v.n_worlds || (!dst && max_records > 0) || max_records < 0) {
    pg_set_error("pg_worlds_debug_geometry: bad argument");
    return PG_ERR_ARG;
  }
  const int total = w->h_counts[world * 3 + 0] + w->h_counts[world * 3 + 1] + w->h_counts[world * 3 + 2];
  if (total > max_records) { pg_set_error("pg_worlds_debug_geometry: need room for %d records", total); return PG_ERR_CAPACITY; }
  if (total == 0) return 0;
  if (w->mode == PG_MODE_GENERAL && !w->v.bodies[0]) { pg_set_error("pg_worlds_debug_geometry: the world holds no records"); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  if (dst_is_device) {
    const int rc = pg_launch_debug_geometry(w, world, total, static_cast<PgDebugDraw *>(dst));
    return rc == PG_OK ? total : rc;
  }
  PgDebugDraw *d_tmp = nullptr;
  PG_CUDA(cudaMallocAsync(&d_tmp, sizeof(PgDebugDraw) * (size_t)total, w->stream));
  int rc = pg_launch_debug_geometry(w, world, total, d_tmp);
  if (rc == PG_OK) {
    cudaError_t e = cudaMemcpyAsync(dst, d_tmp, sizeof(PgDebugDraw) * (size_t)total, cudaMemcpyDeviceToHost, w->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) { pg_set_error("pg_worlds_debug_geometry: %s", cudaGetErrorString(e)); rc = PG_ERR_CUDA; }
  }
  cudaFreeAsync(d_tmp, w->stream);
  return rc == PG_OK ? total : rc;
}

// Unit meshes of the debug renderer.  Both are latitude rings of kRing + 1 vertices (the seam vertex is doubled) joined
// by two triangles per quad; the arithmetic types follow debug_renderer.c to the letter (float steps from double
// constants, `double - float` for the latitude, float products), because the vertices are compared bit for bit.
constexpr int kRing = 20;        // sectors per ring and stacks per hemisphere (debug_renderer.c:373-374, :477-478)

// ring of vertices at height z and radius rad around `centre`, optionally rotated: returns the next free vertex slot
static int ring_vertices(float *out, int at, float rad, float z, const float *rot16_or_null, const float *centre3_or_null) {
  const float dtheta = (2 * M_PI) / kRing;                       // float <- double, as upstream
  for (int col = 0; col <= kRing; col++) {
    const float theta = col * dtheta;
    const float x = rad * cosf(theta), y = rad * sinf(theta);
    float *v = out + 3 * at++;
    if (!rot16_or_null) { v[0] = x; v[1] = y; v[2] = z; continue; }
    // glm_mat4_mulv3(rot, (x, y, z), 1.0f) + centre: four products per row summed left to right
    for (int k = 0; k < 3; k++)
      v[k] = (rot16_or_null[0 + k] * x + rot16_or_null[4 + k] * y + rot16_or_null[8 + k] * z + rot16_or_null[12 + k] * 1.0f) + centre3_or_null[k];
  }
  return at;
}

// sphere of physics_debug_sphere_init (debug_renderer.c:371-432): 21 rings from the +z pole down, 2280 indices (the two
// polar rows contribute one triangle per quad)
void pg_debug_sphere_mesh(float radius, float *vertices, uint32_t *indices) {
  const float dphi = M_PI / kRing;
  int at = 0;
  for (int ring = 0; ring <= kRing; ring++) {
    const float phi = (M_PI / 2) - (ring * dphi);                  // double arithmetic on the float product
    at = ring_vertices(vertices, at, radius * cosf(phi), radius * sinf(phi), nullptr, nullptr);
  }
  uint32_t *w = indices;
  for (int ring = 0; ring < kRing; ring++)
    for (int col = 0; col < kRing; col++) {
      const uint32_t upper = (uint32_t)(ring * (kRing + 1) + col), lower = upper + kRing + 1;
      if (ring != 0) { *w++ = upper; *w++ = lower; *w++ = upper + 1; }
      if (ring != kRing - 1) { *w++ = upper + 1; *w++ = lower; *w++ = lower + 1; }
    }
}

// capsule of physics_debug_capsule_init (debug_renderer.c:469-616): a hemisphere of 20 rings below end point A, the two
// equators, a hemisphere above end point B, all built about +z and turned onto the segment's axis; 42 rings, 4920 indices
void pg_debug_capsule_mesh(const float *A, const float *B, float radius, float *vertices, uint32_t *indices) {
  float axis[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
  h_normalize(axis);
  // rotation taking +z to the axis: about z x axis by acos(z . axis); a half turn about x when the axis is -z
  float about[3] = {0.0f * axis[2] - 1.0f * axis[1], 1.0f * axis[0] - 0.0f * axis[2], 0.0f * axis[1] - 0.0f * axis[0]};
  const float cos_angle = 0.0f * axis[0] + 0.0f * axis[1] + 1.0f * axis[2];
  const float angle = acosf(cos_angle);
  float rot[16];
  if (h_norm(about) > 0.0f) {
    h_normalize(about);
    h_rotate_make(rot, angle, about);
  } else {
    for (int k = 0; k < 16; k++) rot[k] = (k % 5 == 0) ? 1.0f : 0.0f;
    if (cos_angle < 0) { const float x_axis[3] = {1.0f, 0.0f, 0.0f}; h_rotate_make(rot, M_PI, x_axis); }
  }
  const int n_rings = 2 * kRing + 2;
  int at = 0;
  for (int ring = 0; ring < n_rings; ring++) {
    const bool lower_half = ring <= kRing;
    const int from_equator = lower_half ? ring : ring - (kRing + 1);          // 0 at ... see below
    float rad = radius, z = 0.0f;
    if (ring < kRing) {                                                         // below A: t = ring / 20
      const float t = ((float)ring / kRing);
      rad = radius * cosf(t * M_PI / 2.0f); z = -radius * sinf(t * M_PI / 2.0f);
    } else if (ring > kRing + 1) {                                              // above B
      const float t = ((float)from_equator / kRing);
      rad = radius * cosf(t * M_PI / 2.0f); z = radius * sinf(t * M_PI / 2.0f);
    }
    at = ring_vertices(vertices, at, rad, z, rot, lower_half ? A : B);
  }
  uint32_t *w = indices;
  for (int ring = 0; ring + 1 < n_rings; ring++)
    for (int col = 0; col < kRing; col++) {
      const uint32_t upper = (uint32_t)(ring * (kRing + 1) + col), lower = upper + kRing + 1;
      *w++ = upper; *w++ = upper + 1; *w++ = lower;
      *w++ = upper + 1; *w++ = lower + 1; *w++ = lower;
    }
}

// physics_debug_AABB_init, debug_renderer.c:311-337: local corners and the 12 edges as 24 line indices
void pg_debug_box_mesh(const float *c, const float *e, float *vertices, uint32_t *indices) {
  static const int sgn[8][3] = {{+1, +1, +1}, {+1, +1, -1}, {+1, -1, -1}, {+1, -1, +1}, {-1, -1, -1}, {-1, -1, +1}, {-1, +1, +1}, {-1, +1, -1}};
  static const uint32_t idx[24] = {0, 1, 1, 2, 2, 3, 3, 0, 0, 6, 1, 7, 2, 4, 3, 5, 4, 5, 5, 6, 6, 7, 7, 4};
  for (int k = 0; k < 8; k++)
    for (int a = 0; a < 3; a++) vertices[3 * k + a] = sgn[k][a] > 0 ? c[a] + e[a] : c[a] - e[a];
  for (int k = 0; k < 24; k++) indices[k] = idx[k];
}

// model of a capsule drawn without a scene node (debug_renderer.c:53-59, :109-116):
// identity, glm_translate(position), glm_rotate_y, glm_rotate_x, glm_rotate_z (angles through glm_rad), glm_scale
void pg_debug_capsule_model(const PgBody *b, float *m) {
  for (int k = 0; k < 16; k++) m[k] = (k % 5 == 0) ? 1.0f : 0.0f;
  float v1[4], v2[4], v3_[4];
  for (int k = 0; k < 4; k++) { v1[k] = m[0 + k] * b->position[0]; v2[k] = m[4 + k] * b->position[1]; v3_[k] = m[8 + k] * b->position[2]; }
  for (int k = 0; k < 4; k++) m[12 + k] = v1[k] + m[12 + k];
  for (int k = 0; k < 4; k++) m[12 + k] = v2[k] + m[12 + k];
  for (int k = 0; k < 4; k++) m[12 + k] = v3_[k] + m[12 + k];
  float t[16];
  auto ident = [&]() { for (int k = 0; k < 16; k++) t[k] = (k % 5 == 0) ? 1.0f : 0.0f; };
  { const float a = h_rad(b->rotation[1]), c = cosf(a), s = sinf(a); ident(); t[0] = c; t[2] = -s; t[8] = s; t[10] = c; h_mul_rot(m, t); }
  { const float a = h_rad(b->rotation[0]), c = cosf(a), s = sinf(a); ident(); t[5] = c; t[6] = s; t[9] = -s; t[10] = c; h_mul_rot(m, t); }
  { const float a = h_rad(b->rotation[2]), c = cosf(a), s = sinf(a); ident(); t[0] = c; t[1] = s; t[4] = -s; t[5] = c; h_mul_rot(m, t); }
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < 4; k++) m[c * 4 + k] = m[c * 4 + k] * b->scale[c];
}

}  // extern "C"
