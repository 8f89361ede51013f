// physics_gpu.cu -- C-ABI entry points of libphysics_gpu.so (include/physics_gpu.h): handle
// lifetime, host<->device state exchange, launches.  Kernels live in pg_general.cu (any collider
// mix, block per world), pg_spheres.cu (sphere batches, warp per world) and pg_large.cu (one big
// world).  No CPU fallback: without a usable CUDA device every call fails with PG_ERR_CUDA.
#include "pg_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <ctime>
#include <vector>

static thread_local char g_err[512] = "";

void pg_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Enumerate the leaves of world.c:557-582's recursion for the root (t0, t1): same float midpoints,
// same `end - start < 1e-6` stop rule.  2^15 leaves for dt = 0.01666; cached per (t0, t1).
static void leaf_rec(float a, float b, float *lo, float *hi, int depth) {
  const float len = b - a;
  if (len < 0.000001f || depth > 40 || !(len == len)) { if (len < *lo) *lo = len; if (len > *hi) *hi = len; return; }
  const float mid = (a + b) * 0.5f;
  leaf_rec(a, mid, lo, hi, depth + 1);
  leaf_rec(mid, b, lo, hi, depth + 1);
}
void pg_leaf_lengths(float t0, float t1, float *lo, float *hi) {
  static thread_local float c_t0 = -1.0f, c_t1 = -1.0f, c_lo = 0.0f, c_hi = 0.0f;
  if (!(t0 == c_t0 && t1 == c_t1)) {
    float l = 1.0e30f, h = -1.0e30f;
    if (t1 - t0 < 1.0f && t1 > t0) leaf_rec(t0, t1, &l, &h, 0);
    if (!(l <= h) || l <= 0.0f) { l = 0.4990e-6f; h = 1.0e-6f; }   // degenerate root: generic bounds
    c_t0 = t0; c_t1 = t1; c_lo = l; c_hi = h;
  }
  *lo = c_lo; *hi = c_hi;
}

namespace {

// glm_euler_xyz upper 3x3 (cglm euler.h), host libm sinf/cosf -- same library the reference links.
void euler_xyz(const float *a, float *rot) {
  float sx = sinf(a[0]), cx = cosf(a[0]);
  float sy = sinf(a[1]), cy = cosf(a[1]);
  float sz = sinf(a[2]), cz = cosf(a[2]);
  float czsx = cz * sx, cxcz = cx * cz, sysz = sy * sz;
  rot[0] = cy * cz;
  rot[1] = czsx * sy + cx * sz;
  rot[2] = -cxcz * sy + sx * sz;
  rot[3] = -cy * sz;
  rot[4] = cxcz - sx * sysz;
  rot[5] = czsx + cx * sysz;
  rot[6] = sy;
  rot[7] = -cy * sx;
  rot[8] = cx * cy;
}

inline float to_rad(float deg) { return deg * 3.14159265358979323846264338327950288f / 180.0f; }   // glm_rad

bool valid_world(const PgWorlds *w, int world) { return w && world >= 0 && world < w->v.n_worlds; }

// Can a resolver write to a STATIC body?  (sphere-sphere always moves both, box-capsule always
// moves the capsule, ...; see pg_general.cu.)  Pass 3 may run one thread per dynamic body only if not.
bool statics_read_only(const PgBody *st, int n) {
  for (int i = 0; i < n; i++)
    if (st[i].collider_type != PG_AABB && st[i].collider_type != PG_PLANE) return false;
  return true;
}

int refresh_parallel_flag(PgWorlds *w) {
  // conservative: inspect every world's statics and dynamics on the host copy
  std::vector<PgBody> tmp;
  bool ok = true;
  for (int wd = 0; wd < w->v.n_worlds && ok; wd++) {
    int ns = w->h_counts[wd * 3 + 0], nd = w->h_counts[wd * 3 + 1];
    if (ns) {
      tmp.resize(ns);
      PG_CUDA(cudaMemcpy(tmp.data(), w->v.bodies[0] + (size_t)wd * w->v.cap[0], sizeof(PgBody) * ns, cudaMemcpyDeviceToHost));
      ok = ok && statics_read_only(tmp.data(), ns);
    }
    if (nd && ok) {
      tmp.resize(nd);
      PG_CUDA(cudaMemcpy(tmp.data(), w->v.bodies[1] + (size_t)wd * w->v.cap[1], sizeof(PgBody) * nd, cudaMemcpyDeviceToHost));
      for (int i = 0; i < nd; i++) if (tmp[i].collider_type == PG_PLANE) ok = false;   // AABB-plane resolver writes the AABB
    }
  }
  w->static_pass_parallel_ok = ok;
  return PG_OK;
}

}  // namespace

extern "C" {

const char *pg_last_error(void) { return g_err; }

int pg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void pg_body_prepare(PgBody *b, int n) {
  for (int i = 0; i < n; i++) {
    euler_xyz(b[i].rotation, b[i].euler_raw);
    float rad[3] = {to_rad(b[i].rotation[0]), to_rad(b[i].rotation[1]), to_rad(b[i].rotation[2])};
    euler_xyz(rad, b[i].euler_node);
  }
}

void pg_node_transform(const float *pos, const float *rot_deg, const float *scale,
                       const float *parent, float *world) {
  float rad[3] = {to_rad(rot_deg[0]), to_rad(rot_deg[1]), to_rad(rot_deg[2])};
  float r[9];
  euler_xyz(rad, r);
  float local[16];
  for (int c = 0; c < 3; c++) {
    for (int k = 0; k < 3; k++) local[c * 4 + k] = r[c * 3 + k] * scale[c];
    local[c * 4 + 3] = 0.0f * scale[c];
  }
  local[12] = pos[0]; local[13] = pos[1]; local[14] = pos[2]; local[15] = 1.0f;
  if (!parent) { memcpy(world, local, sizeof(local)); return; }
  float out[16];
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++)
      out[c * 4 + k] = parent[0 * 4 + k] * local[c * 4 + 0] + parent[1 * 4 + k] * local[c * 4 + 1] +
                       parent[2 * 4 + k] * local[c * 4 + 2] + parent[3 * 4 + k] * local[c * 4 + 3];
  memcpy(world, out, sizeof(out));
}

void *pg_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { pg_set_error("cudaHostAlloc(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
  return p;
}
void pg_host_free(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------ batched worlds

PgWorlds *pg_worlds_create(int n_worlds, int max_static, int max_dynamic, int max_player,
                           int max_events_per_world, int device) {
  if (n_worlds <= 0 || max_static < 0 || max_dynamic < 0 || max_player < 0 || max_events_per_world < 0) {
    pg_set_error("pg_worlds_create: bad argument");
    return nullptr;
  }
  PG_CUDA_NULL(cudaSetDevice(device));
  PgWorlds *w = new PgWorlds();
  memset(w, 0, sizeof(*w));
  w->device = device;
  w->mode = PG_MODE_GENERAL;
  w->v.n_worlds = n_worlds;
  w->v.cap[0] = std::max(max_static, 1);
  w->v.cap[1] = std::max(max_dynamic, 1);
  w->v.cap[2] = std::max(max_player, 1);
  w->v.max_events = max_events_per_world;
  w->static_pass_parallel_ok = true;
  cudaDeviceProp prop;
  PG_CUDA_NULL(cudaGetDeviceProperties(&prop, device));
  w->sm_count = prop.multiProcessorCount;
  PG_CUDA_NULL(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
  PG_CUDA_NULL(cudaEventCreate(&w->ev0));
  PG_CUDA_NULL(cudaEventCreate(&w->ev1));
  PG_CUDA_NULL(cudaMalloc(&w->v.counts, sizeof(int) * 3 * (size_t)n_worlds));
  PG_CUDA_NULL(cudaMemset(w->v.counts, 0, sizeof(int) * 3 * (size_t)n_worlds));
  PG_CUDA_NULL(cudaMalloc(&w->v.n_events, sizeof(int) * (size_t)n_worlds));
  PG_CUDA_NULL(cudaMemset(w->v.n_events, 0, sizeof(int) * (size_t)n_worlds));
  if (max_events_per_world > 0)
    PG_CUDA_NULL(cudaMalloc(&w->v.events, sizeof(PgEvent) * (size_t)n_worlds * max_events_per_world));
  PG_CUDA_NULL(cudaMalloc(&w->d_stats, 8 * sizeof(double)));
  PG_CUDA_NULL(cudaMalloc(&w->d_stats_sum, 8 * sizeof(double)));
  PG_CUDA_NULL(cudaMemset(w->d_stats, 0, 8 * sizeof(double)));
  PG_CUDA_NULL(cudaMemset(w->d_stats_sum, 0, 8 * sizeof(double)));
  PG_CUDA_NULL(cudaMalloc(&w->d_work_counter, sizeof(int)));
  PG_CUDA_NULL(cudaMalloc(&w->d_cost, sizeof(unsigned) * (size_t)n_worlds));
  PG_CUDA_NULL(cudaMalloc(&w->d_order, sizeof(int) * (size_t)n_worlds));
  w->cost_valid = false;
  w->h_counts = (int *)calloc((size_t)n_worlds * 3, sizeof(int));
  return w;
}

void pg_worlds_destroy(PgWorlds *w) {
  if (!w) return;
  cudaSetDevice(w->device);
  cudaStreamSynchronize(w->stream);
  for (int c = 0; c < 3; c++) if (w->v.bodies[c]) cudaFree(w->v.bodies[c]);
  if (w->v.counts) cudaFree(w->v.counts);
  if (w->v.events) cudaFree(w->v.events);
  if (w->v.n_events) cudaFree(w->v.n_events);
  if (w->v.pos_rest) cudaFree(w->v.pos_rest);
  if (w->v.vel_rad) cudaFree(w->v.vel_rad);
  if (w->v.planes) cudaFree(w->v.planes);
  if (w->d_stats) cudaFree(w->d_stats);
  if (w->d_stats_sum) cudaFree(w->d_stats_sum);
  if (w->d_work_counter) cudaFree(w->d_work_counter);
  if (w->d_cost) cudaFree(w->d_cost);
  if (w->d_order) cudaFree(w->d_order);
  if (w->pipe_ready) {
    for (int k = 0; k < 3; k++) cudaStreamDestroy(w->pipe_stream[k]);
    for (int k = 0; k < PG_PIPE_DMA; k++) { cudaStreamDestroy(w->pipe_up[k]); cudaStreamDestroy(w->pipe_down[k]); }
    for (int k = 0; k < PG_PIPE_MAX_CHUNKS; k++) { cudaEventDestroy(w->pipe_up_ev[k]); cudaEventDestroy(w->pipe_k_ev[k]); }
    for (int k = 0; k < 2 * PG_PIPE_DMA; k++) cudaEventDestroy(w->pipe_done[k]);
    for (int k = 0; k < 3; k++) cudaEventDestroy(w->pipe_join[k]);
    cudaEventDestroy(w->pipe_fork);
    if (w->pipe_graph) cudaGraphExecDestroy(w->pipe_graph);
    cudaFree(w->d_pipe_counters);
  }
  if (w->d_stage_pos) { cudaFree(w->d_stage_pos); cudaFree(w->d_stage_vel); cudaFree(w->d_stage_rest); }
  if (w->d_stage_packed) cudaFree(w->d_stage_packed);
  if (w->h_frame) cudaFreeHost(w->h_frame);
  for (int c = 0; c < 3; c++) free(w->h_types[c]);
  cudaEventDestroy(w->ev0);
  cudaEventDestroy(w->ev1);
  cudaStreamDestroy(w->stream);
  free(w->h_counts);
  delete w;
}

int pg_worlds_count(const PgWorlds *w) { return w ? w->v.n_worlds : 0; }

static int ensure_records(PgWorlds *w) {
  for (int c = 0; c < 3; c++) {
    if (!w->v.bodies[c]) {
      size_t bytes = sizeof(PgBody) * (size_t)w->v.n_worlds * w->v.cap[c];
      PG_CUDA(cudaMalloc(&w->v.bodies[c], bytes));
      PG_CUDA(cudaMemset(w->v.bodies[c], 0, bytes));
    }
  }
  return PG_OK;
}

static int push_counts(PgWorlds *w, int world) {
  PG_CUDA(cudaMemcpyAsync(w->v.counts + world * 3, w->h_counts + world * 3, 3 * sizeof(int), cudaMemcpyHostToDevice, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  return PG_OK;
}

int pg_worlds_set_bodies(PgWorlds *w, int world, int cls, const PgBody *bodies, int n) {
  if (!valid_world(w, world) || cls < 0 || cls > 2 || n < 0) { pg_set_error("pg_worlds_set_bodies: bad argument"); return PG_ERR_ARG; }
  if (w->mode != PG_MODE_GENERAL) { pg_set_error("handle is in sphere mode"); return PG_ERR_ARG; }
  if (n > w->v.cap[cls]) { pg_set_error("class %d: %d bodies > capacity %d", cls, n, w->v.cap[cls]); return PG_ERR_CAPACITY; }
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  int rc = ensure_records(w);
  if (rc) return rc;
  if (n)   // verbatim: the add-time field rules (world.c:41-142) belong to pg_worlds_add_body
    PG_CUDA(cudaMemcpy(w->v.bodies[cls] + (size_t)world * w->v.cap[cls], bodies, sizeof(PgBody) * (size_t)n, cudaMemcpyHostToDevice));
  w->h_counts[world * 3 + cls] = n;
  rc = push_counts(w, world);
  if (rc) return rc;
  return refresh_parallel_flag(w);
}

int pg_worlds_update_bodies(PgWorlds *w, int world, int cls, int first, const PgBody *bodies, int n) {
  if (!valid_world(w, world) || cls < 0 || cls > 2 || first < 0 || n < 0 || (n > 0 && !bodies)) { pg_set_error("pg_worlds_update_bodies: bad argument"); return PG_ERR_ARG; }
  if (w->mode != PG_MODE_GENERAL) { pg_set_error("handle is in sphere mode"); return PG_ERR_ARG; }
  if (first + n > w->h_counts[world * 3 + cls]) { pg_set_error("pg_worlds_update_bodies: range [%d,%d) exceeds the class length %d", first, first + n, w->h_counts[world * 3 + cls]); return PG_ERR_ARG; }
  if (n == 0) return PG_OK;
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  PG_CUDA(cudaMemcpy(w->v.bodies[cls] + (size_t)world * w->v.cap[cls] + first, bodies, sizeof(PgBody) * (size_t)n, cudaMemcpyHostToDevice));
  // the uploaded records can only withdraw the "statics are read-only in pass 3" licence, never grant it
  if (cls == PG_CLASS_STATIC && !statics_read_only(bodies, n)) w->static_pass_parallel_ok = false;
  if (cls == PG_CLASS_DYNAMIC) for (int i = 0; i < n; i++) if (bodies[i].collider_type == PG_PLANE) w->static_pass_parallel_ok = false;
  return PG_OK;
}

int pg_worlds_add_body(PgWorlds *w, int world, const PgBody *body, int as_player) {
  if (!valid_world(w, world) || !body) { pg_set_error("pg_worlds_add_body: bad argument"); return PG_ERR_ARG; }
  if (w->mode != PG_MODE_GENERAL) { pg_set_error("handle is in sphere mode"); return PG_ERR_ARG; }
  // collider type check of world.c:56 / :125 (`> COLLIDER_COUNT`, so 4 slips through upstream; rejected here)
  if (body->collider_type < 0 || body->collider_type > PG_PLANE) {
    pg_set_error("Error: collider type provided to physics_add_body is invalid");
    return PG_ERR_ARG;
  }
  int cls = as_player ? PG_CLASS_PLAYER : (body->dynamic ? PG_CLASS_DYNAMIC : PG_CLASS_STATIC);
  int n = w->h_counts[world * 3 + cls];
  if (n >= w->v.cap[cls]) { pg_set_error("class %d full (%d)", cls, n); return PG_ERR_CAPACITY; }
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  int rc = ensure_records(w);
  if (rc) return rc;
  PgBody tmp = *body;
  if (cls == PG_CLASS_STATIC) { tmp.velocity[0] = tmp.velocity[1] = tmp.velocity[2] = 0.0f; }   // world.c:51
  if (cls == PG_CLASS_PLAYER) { tmp.restitution = 0.0f; tmp.dynamic = 0; }                           // world.c:137 (+calloc)
  PG_CUDA(cudaMemcpy(w->v.bodies[cls] + (size_t)world * w->v.cap[cls] + n, &tmp, sizeof(PgBody), cudaMemcpyHostToDevice));
  w->h_counts[world * 3 + cls] = n + 1;
  rc = push_counts(w, world);
  if (rc) return rc;
  rc = refresh_parallel_flag(w);
  return rc ? rc : n;
}

int pg_worlds_remove_body(PgWorlds *w, int world, int cls, int index) {
  if (!valid_world(w, world) || cls < 0 || cls > 2) { pg_set_error("pg_worlds_remove_body: bad argument"); return PG_ERR_ARG; }
  if (w->mode != PG_MODE_GENERAL) { pg_set_error("handle is in sphere mode"); return PG_ERR_ARG; }
  int n = w->h_counts[world * 3 + cls];
  if (index < 0 || index >= n) { pg_set_error("pg_worlds_remove_body: index %d out of %d", index, n); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  PgBody *base = w->v.bodies[cls] + (size_t)world * w->v.cap[cls];
  if (index != n - 1) PG_CUDA(cudaMemcpy(base + index, base + (n - 1), sizeof(PgBody), cudaMemcpyDeviceToDevice));
  w->h_counts[world * 3 + cls] = n - 1;
  return push_counts(w, world);
}

int pg_worlds_body_count(const PgWorlds *w, int world, int cls) {
  if (!valid_world(w, world) || cls < 0 || cls > 2) return PG_ERR_ARG;
  return w->h_counts[world * 3 + cls];
}

int pg_worlds_set_spheres(PgWorlds *w, int first, int count, const PgBody *statics, int n_static,
                          int n_spheres, const float *pos, const float *vel, float radius,
                          float restitution) {
  if (!w || first < 0 || count <= 0 || first + count > w->v.n_worlds || n_static < 0 || n_spheres < 0 || !pos || !vel) {
    pg_set_error("pg_worlds_set_spheres: bad argument");
    return PG_ERR_ARG;
  }
  if (n_static > w->v.cap[0] || n_spheres > w->v.cap[1]) { pg_set_error("pg_worlds_set_spheres: over capacity"); return PG_ERR_CAPACITY; }
  for (int i = 0; i < n_static; i++)
    if (statics[i].collider_type != PG_PLANE) { pg_set_error("sphere mode: static bodies must be planes"); return PG_ERR_UNSUPPORTED; }
  if (w->v.bodies[0] || w->v.bodies[1] || w->v.bodies[2]) { pg_set_error("handle already holds general bodies"); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  const size_t nd_all = (size_t)w->v.n_worlds * w->v.cap[1];
  if (!w->v.pos_rest) {
    PG_CUDA(cudaMalloc(&w->v.pos_rest, sizeof(float4) * nd_all));
    PG_CUDA(cudaMalloc(&w->v.vel_rad, sizeof(float4) * nd_all));
    PG_CUDA(cudaMalloc(&w->v.planes, sizeof(float4) * (size_t)w->v.n_worlds * w->v.cap[0]));
    PG_CUDA(cudaMemset(w->v.pos_rest, 0, sizeof(float4) * nd_all));
    PG_CUDA(cudaMemset(w->v.vel_rad, 0, sizeof(float4) * nd_all));
  }
  w->mode = PG_MODE_SPHERES;
  w->stats_valid = false;
  w->v.restitution = restitution;
  if (w->pipe_graph) { cudaGraphExecDestroy(w->pipe_graph); w->pipe_graph = nullptr; }   // the captured kernels carry the old view by value
  std::vector<float4> hp((size_t)count * w->v.cap[1]), hv((size_t)count * w->v.cap[1]), pl((size_t)count * w->v.cap[0]);
  for (int wd = 0; wd < count; wd++) {
    for (int i = 0; i < n_spheres; i++) {
      const float *p = pos + ((size_t)wd * n_spheres + i) * 3, *q = vel + ((size_t)wd * n_spheres + i) * 3;
      hp[(size_t)wd * w->v.cap[1] + i] = make_float4(p[0], p[1], p[2], 0.0f);
      hv[(size_t)wd * w->v.cap[1] + i] = make_float4(q[0], q[1], q[2], radius * 1.0f);   // radius * scale[0], scale = 1
    }
    for (int i = 0; i < n_static; i++)
      pl[(size_t)wd * w->v.cap[0] + i] = make_float4(statics[i].shape[0], statics[i].shape[1], statics[i].shape[2], statics[i].shape[3]);
    w->h_counts[(first + wd) * 3 + 0] = n_static;
    w->h_counts[(first + wd) * 3 + 1] = n_spheres;
    w->h_counts[(first + wd) * 3 + 2] = 0;
  }
  PG_CUDA(cudaMemcpy(w->v.pos_rest + (size_t)first * w->v.cap[1], hp.data(), sizeof(float4) * hp.size(), cudaMemcpyHostToDevice));
  PG_CUDA(cudaMemcpy(w->v.vel_rad + (size_t)first * w->v.cap[1], hv.data(), sizeof(float4) * hv.size(), cudaMemcpyHostToDevice));
  PG_CUDA(cudaMemcpy(w->v.planes + (size_t)first * w->v.cap[0], pl.data(), sizeof(float4) * pl.size(), cudaMemcpyHostToDevice));
  PG_CUDA(cudaMemcpy(w->v.counts + first * 3, w->h_counts + first * 3, sizeof(int) * 3 * (size_t)count, cudaMemcpyHostToDevice));
  return PG_OK;
}

int pg_worlds_get_bodies(PgWorlds *w, int world, int cls, PgBody *out, int cap) {
  if (!valid_world(w, world) || cls < 0 || cls > 2 || !out) { pg_set_error("pg_worlds_get_bodies: bad argument"); return PG_ERR_ARG; }
  int n = w->h_counts[world * 3 + cls];
  if (n > cap) { pg_set_error("pg_worlds_get_bodies: need room for %d records", n); return PG_ERR_CAPACITY; }
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  if (n == 0) return 0;
  if (w->mode == PG_MODE_GENERAL) {
    PG_CUDA(cudaMemcpy(out, w->v.bodies[cls] + (size_t)world * w->v.cap[cls], sizeof(PgBody) * (size_t)n, cudaMemcpyDeviceToHost));
    return n;
  }
  // sphere mode: synthesise records from the SoA state
  memset(out, 0, sizeof(PgBody) * (size_t)n);
  if (cls == PG_CLASS_STATIC) {
    std::vector<float4> pl(n);
    PG_CUDA(cudaMemcpy(pl.data(), w->v.planes + (size_t)world * w->v.cap[0], sizeof(float4) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
      out[i].collider_type = PG_PLANE; out[i].entity_type = PG_ENT_WORLD;
      out[i].scale[0] = out[i].scale[1] = out[i].scale[2] = 1.0f;
      out[i].shape[0] = pl[i].x; out[i].shape[1] = pl[i].y; out[i].shape[2] = pl[i].z; out[i].shape[3] = pl[i].w;
    }
  } else if (cls == PG_CLASS_DYNAMIC) {
    std::vector<float4> hp(n), hv(n);
    PG_CUDA(cudaMemcpy(hp.data(), w->v.pos_rest + (size_t)world * w->v.cap[1], sizeof(float4) * n, cudaMemcpyDeviceToHost));
    PG_CUDA(cudaMemcpy(hv.data(), w->v.vel_rad + (size_t)world * w->v.cap[1], sizeof(float4) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
      out[i].collider_type = PG_SPHERE; out[i].entity_type = PG_ENT_WORLD; out[i].dynamic = 1;
      out[i].scale[0] = out[i].scale[1] = out[i].scale[2] = 1.0f;
      out[i].restitution = w->v.restitution;
      out[i].position[0] = hp[i].x; out[i].position[1] = hp[i].y; out[i].position[2] = hp[i].z;
      out[i].velocity[0] = hv[i].x; out[i].velocity[1] = hv[i].y; out[i].velocity[2] = hv[i].z;
      out[i].shape[3] = hv[i].w;
      out[i].at_rest = hp[i].w != 0.0f;
    }
  }
  pg_body_prepare(out, n);
  return n;
}

static int sphere_range_count(PgWorlds *w, int first, int count, int *n_out) {
  const int n = w->h_counts[first * 3 + 1];
  for (int wd = 1; wd < count; wd++)
    if (w->h_counts[(first + wd) * 3 + 1] != n) { pg_set_error("sphere mode state exchange needs the same sphere count in every world of the range"); return PG_ERR_ARG; }
  *n_out = n;
  return PG_OK;
}
static int ensure_stage(PgWorlds *w, size_t bodies) {
  if (bodies <= w->stage_bodies) return PG_OK;
  if (w->d_stage_pos) { cudaFree(w->d_stage_pos); cudaFree(w->d_stage_vel); cudaFree(w->d_stage_rest); }
  PG_CUDA(cudaMalloc(&w->d_stage_pos, sizeof(float) * 3 * bodies));
  PG_CUDA(cudaMalloc(&w->d_stage_vel, sizeof(float) * 3 * bodies));
  PG_CUDA(cudaMalloc(&w->d_stage_rest, bodies));
  w->stage_bodies = bodies;
  return PG_OK;
}

// SoA state exchange.  In sphere mode this is two strided copies; in general mode a gather/scatter
// through a host staging buffer (game-sized worlds).
int pg_worlds_upload_state(PgWorlds *w, int first, int count, const float *pos, const float *vel,
                           const uint8_t *at_rest) {
  if (!w || first < 0 || count <= 0 || first + count > w->v.n_worlds || !pos || !vel) { pg_set_error("pg_worlds_upload_state: bad argument"); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  const int cap = w->v.cap[1];
  if (w->mode == PG_MODE_SPHERES) {
    int n = 0;
    int rc = sphere_range_count(w, first, count, &n);
    if (rc) return rc;
    const size_t total = (size_t)count * n;
    rc = ensure_stage(w, total);
    if (rc) return rc;
    w->stats_valid = false;
    // one H2D copy per array (async when the caller's buffers are pinned), then a pack kernel
    PG_CUDA(cudaMemcpyAsync(w->d_stage_pos, pos, sizeof(float) * 3 * total, cudaMemcpyHostToDevice, w->stream));
    PG_CUDA(cudaMemcpyAsync(w->d_stage_vel, vel, sizeof(float) * 3 * total, cudaMemcpyHostToDevice, w->stream));
    if (at_rest) PG_CUDA(cudaMemcpyAsync(w->d_stage_rest, at_rest, total, cudaMemcpyHostToDevice, w->stream));
    return pg_launch_pack_state(w, first, count, n, w->d_stage_pos, w->d_stage_vel, at_rest ? w->d_stage_rest : nullptr);
  }
  std::vector<PgBody> tmp((size_t)count * cap);
  PG_CUDA(cudaMemcpy(tmp.data(), w->v.bodies[1] + (size_t)first * cap, sizeof(PgBody) * tmp.size(), cudaMemcpyDeviceToHost));
  size_t k = 0;
  for (int wd = 0; wd < count; wd++) {
    const int n = w->h_counts[(first + wd) * 3 + 1];
    for (int i = 0; i < n; i++, k++) {
      PgBody &b = tmp[(size_t)wd * cap + i];
      memcpy(b.position, pos + 3 * k, 12);
      memcpy(b.velocity, vel + 3 * k, 12);
      if (at_rest) b.at_rest = at_rest[k] != 0;
    }
  }
  PG_CUDA(cudaMemcpy(w->v.bodies[1] + (size_t)first * cap, tmp.data(), sizeof(PgBody) * tmp.size(), cudaMemcpyHostToDevice));
  return PG_OK;
}

int pg_worlds_download_state(PgWorlds *w, int first, int count, float *pos, float *vel, uint8_t *at_rest) {
  if (!w || first < 0 || count <= 0 || first + count > w->v.n_worlds || !pos || !vel) { pg_set_error("pg_worlds_download_state: bad argument"); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  const int cap = w->v.cap[1];
  size_t k = 0;
  if (w->mode == PG_MODE_SPHERES) {
    int n = 0;
    int rc = sphere_range_count(w, first, count, &n);
    if (rc) return rc;
    const size_t total = (size_t)count * n;
    rc = ensure_stage(w, total);
    if (rc) return rc;
    rc = pg_launch_unpack_state(w, first, count, n, w->d_stage_pos, w->d_stage_vel, at_rest ? w->d_stage_rest : nullptr);
    if (rc) return rc;
    PG_CUDA(cudaMemcpyAsync(pos, w->d_stage_pos, sizeof(float) * 3 * total, cudaMemcpyDeviceToHost, w->stream));
    PG_CUDA(cudaMemcpyAsync(vel, w->d_stage_vel, sizeof(float) * 3 * total, cudaMemcpyDeviceToHost, w->stream));
    if (at_rest) PG_CUDA(cudaMemcpyAsync(at_rest, w->d_stage_rest, total, cudaMemcpyDeviceToHost, w->stream));
    PG_CUDA(cudaStreamSynchronize(w->stream));
    return PG_OK;
  }
  PG_CUDA(cudaStreamSynchronize(w->stream));
  std::vector<PgBody> tmp((size_t)count * cap);
  PG_CUDA(cudaMemcpy(tmp.data(), w->v.bodies[1] + (size_t)first * cap, sizeof(PgBody) * tmp.size(), cudaMemcpyDeviceToHost));
  for (int wd = 0; wd < count; wd++) {
    const int n = w->h_counts[(first + wd) * 3 + 1];
    for (int i = 0; i < n; i++, k++) {
      const PgBody &b = tmp[(size_t)wd * cap + i];
      memcpy(pos + 3 * k, b.position, 12);
      memcpy(vel + 3 * k, b.velocity, 12);
      if (at_rest) at_rest[k] = b.at_rest != 0;
    }
  }
  return PG_OK;
}

int pg_worlds_step(PgWorlds *w, float dt, int n_steps, int refresh_nodes) {
  if (!w || n_steps < 0) { pg_set_error("pg_worlds_step: bad argument"); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaMemsetAsync(w->v.n_events, 0, sizeof(int) * (size_t)w->v.n_worlds, w->stream));
  PG_CUDA(cudaEventRecord(w->ev0, w->stream));
  {
    float cdt = dt;
    if ((double)cdt > 0.01666) cdt = (float)0.01666;                 // world.c:175-177
    pg_leaf_lengths(0.0f, cdt, &w->v.leaf_lo, &w->v.leaf_hi);
  }
  int rc = PG_OK;
  if (n_steps > 0) {
    if (w->mode == PG_MODE_SPHERES) rc = pg_launch_sphere_step(w, dt, n_steps);
    else {
      rc = ensure_records(w);
      if (!rc) rc = pg_launch_general_step(w, dt, n_steps, refresh_nodes);
    }
  }
  PG_CUDA(cudaEventRecord(w->ev1, w->stream));
  w->last_steps = n_steps;
  return rc;
}

// Host-to-host step: state in from the caller's (pinned) arrays, n_steps steps, state back out into the same
// arrays.  The batch is cut into chunks of worlds; chunk c's upload, chunk c-1's kernel and chunk c-2's download
// overlap (PCIe is full duplex).  The step kernel itself reads the staging buffer the upload landed in and writes
// the one the download leaves from (StageView): no pack / unpack launches, one kernel per chunk.  Uploads and
// downloads are spread over two queues per direction, so that the set-up latency of one DMA hides behind the
// transfer of another.  The whole pipeline is captured once and replayed as one CUDA graph.
struct HostState {
  unsigned char *pos, *vel, *rest;     // host addresses (rest may be null; packed: vel/rest lie inside the record pos points to)
  size_t stride, rest_stride;          // bytes per world
  bool packed;                         // one contiguous record per world: a chunk is ONE copy each way
};

static int step_host_impl(PgWorlds *w, float dt, int n_steps, const HostState &h, int n) {
  const int nw = w->v.n_worlds;
  static const int kChunksEnv = getenv("PG_HOST_CHUNKS") ? atoi(getenv("PG_HOST_CHUNKS")) : 4;      // measured (1 M bodies, packed): 4 -> 0.847, 6 -> 0.860, 8 -> 0.853, 16 -> 0.946 ms
  static const int kDmaEnv = getenv("PG_HOST_DMA_QUEUES") ? atoi(getenv("PG_HOST_DMA_QUEUES")) : 2;
  const int kChunks = std::max(1, std::min(PG_PIPE_MAX_CHUNKS, std::min(kChunksEnv, nw)));
  const int kDma = std::max(1, std::min(PG_PIPE_DMA, kDmaEnv));
  if (!w->pipe_ready) {
    for (int k = 0; k < 3; k++) PG_CUDA(cudaStreamCreateWithFlags(&w->pipe_stream[k], cudaStreamNonBlocking));
    for (int k = 0; k < PG_PIPE_DMA; k++) {
      PG_CUDA(cudaStreamCreateWithFlags(&w->pipe_up[k], cudaStreamNonBlocking));
      PG_CUDA(cudaStreamCreateWithFlags(&w->pipe_down[k], cudaStreamNonBlocking));
      PG_CUDA(cudaEventCreateWithFlags(&w->pipe_done[2 * k], cudaEventDisableTiming));
      PG_CUDA(cudaEventCreateWithFlags(&w->pipe_done[2 * k + 1], cudaEventDisableTiming));
    }
    for (int k = 0; k < PG_PIPE_MAX_CHUNKS; k++) {
      PG_CUDA(cudaEventCreateWithFlags(&w->pipe_up_ev[k], cudaEventDisableTiming));
      PG_CUDA(cudaEventCreateWithFlags(&w->pipe_k_ev[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < 3; k++) PG_CUDA(cudaEventCreateWithFlags(&w->pipe_join[k], cudaEventDisableTiming));
    PG_CUDA(cudaEventCreateWithFlags(&w->pipe_fork, cudaEventDisableTiming));
    PG_CUDA(cudaMalloc(&w->d_pipe_counters, sizeof(int) * PG_PIPE_MAX_CHUNKS));
    w->pipe_ready = true;
  }
  // device staging mirrors the host layout byte for byte
  StageView sv;
  memset(&sv, 0, sizeof(sv));
  sv.stride = h.stride; sv.rest_stride = h.rest_stride; sv.load = 1; sv.store = 1;
  if (h.packed) {
    const size_t bytes = (size_t)nw * h.stride;
    if (bytes > w->stage_packed_bytes) {
      if (w->d_stage_packed) cudaFree(w->d_stage_packed);
  if (w->h_frame) cudaFreeHost(w->h_frame);
  for (int c = 0; c < 3; c++) free(w->h_types[c]);
      w->d_stage_packed = nullptr; w->stage_packed_bytes = 0;
      PG_CUDA(cudaMalloc(&w->d_stage_packed, bytes));
      w->stage_packed_bytes = bytes;
    }
    sv.pos = w->d_stage_packed;
    sv.vel = w->d_stage_packed + (h.vel - h.pos);
    sv.rest = w->d_stage_packed + (h.rest - h.pos);
  } else {
    int rc = ensure_stage(w, (size_t)nw * n);
    if (rc) return rc;
    sv.pos = (unsigned char *)w->d_stage_pos; sv.vel = (unsigned char *)w->d_stage_vel;
    sv.rest = h.rest ? w->d_stage_rest : nullptr;
  }
  {
    float cdt = dt;
    if ((double)cdt > 0.01666) cdt = (float)0.01666;
    pg_leaf_lengths(0.0f, cdt, &w->v.leaf_lo, &w->v.leaf_hi);
  }
  const bool same = w->pipe_graph && w->pipe_key_ptr[0] == h.pos && w->pipe_key_ptr[1] == h.vel && w->pipe_key_ptr[2] == h.rest &&
                    w->pipe_key_dt == dt && w->pipe_key_steps == n_steps && w->pipe_key_chunks == kChunks && w->pipe_key_packed == (h.packed ? 1 : 0);
  static const bool use_graph = getenv("PG_HOST_NO_GRAPH") == nullptr;
  if (use_graph && same) {
    PG_CUDA(cudaEventRecord(w->ev0, w->stream));
    PG_CUDA(cudaGraphLaunch(w->pipe_graph, w->stream));
    PG_CUDA(cudaEventRecord(w->ev1, w->stream));
    w->launches += w->pipe_graph_kernels;
    w->last_steps = n_steps;
    w->cost_valid = false;
    w->stats_valid = true;
    PG_CUDA(cudaStreamSynchronize(w->stream));
    return PG_OK;
  }
  if (w->pipe_graph) { cudaGraphExecDestroy(w->pipe_graph); w->pipe_graph = nullptr; }
  const long long launches_before = w->launches;
  int rc = PG_OK;
  if (use_graph) {
    rc = pg_launch_sphere_range(w, dt, 0, 0, 0, w->stream, w->d_work_counter, nullptr);      // kernel attributes are not capturable
    if (rc) return rc;
    PG_CUDA(cudaStreamSynchronize(w->stream));
    PG_CUDA(cudaEventRecord(w->ev0, w->stream));
    PG_CUDA(cudaStreamBeginCapture(w->stream, cudaStreamCaptureModeThreadLocal));
  } else {
    PG_CUDA(cudaEventRecord(w->ev0, w->stream));
  }
  PG_CUDA(cudaMemsetAsync(w->v.n_events, 0, sizeof(int) * (size_t)nw, w->stream));
  PG_CUDA(cudaMemsetAsync(w->d_stats, 0, 8 * sizeof(double), w->stream));      // every chunk's launch adds its partials
  PG_CUDA(cudaEventRecord(w->pipe_fork, w->stream));
  for (int k = 0; k < 3; k++) PG_CUDA(cudaStreamWaitEvent(w->pipe_stream[k], w->pipe_fork, 0));
  for (int k = 0; k < kDma; k++) {
    PG_CUDA(cudaStreamWaitEvent(w->pipe_up[k], w->pipe_fork, 0));
    PG_CUDA(cudaStreamWaitEvent(w->pipe_down[k], w->pipe_fork, 0));
  }
  // Chunk sizes: what cannot overlap is the first chunk's upload (nothing to download yet) and the last
  // chunk's kernel + download (nothing left to upload), so both ends are small and the middle is large
  // (weights 1, 2, 3, ... up to the middle and back down; PG_HOST_EVEN_CHUNKS for equal sizes).
  int chunk_first[PG_PIPE_MAX_CHUNKS + 1];
  {
    static const bool even = getenv("PG_HOST_EVEN_CHUNKS") != nullptr;
    long long wsum = 0, acc = 0;
    auto weight = [&](int c) { return even ? 2 : 1 + 2 * std::min(c, kChunks - 1 - c); };
    for (int c = 0; c < kChunks; c++) wsum += weight(c);
    for (int c = 0; c < kChunks; c++) { chunk_first[c] = (int)((long long)nw * acc / wsum); acc += weight(c); }
    chunk_first[kChunks] = nw;
  }
  for (int c = 0; c < kChunks; c++) {
    const int first = chunk_first[c], count = chunk_first[c + 1] - first;
    if (count <= 0) continue;
    cudaStream_t st = w->pipe_stream[c % 3], up = w->pipe_up[c % kDma], down = w->pipe_down[c % kDma];
    const size_t off = (size_t)first * h.stride, bytes = (size_t)count * h.stride;
    const size_t roff = (size_t)first * h.rest_stride, rbytes = (size_t)count * h.rest_stride;
    if (h.packed) {
      PG_CUDA(cudaMemcpyAsync(sv.pos + off, h.pos + off, bytes, cudaMemcpyHostToDevice, up));
    } else {
      PG_CUDA(cudaMemcpyAsync(sv.pos + off, h.pos + off, bytes, cudaMemcpyHostToDevice, up));
      PG_CUDA(cudaMemcpyAsync(sv.vel + off, h.vel + off, bytes, cudaMemcpyHostToDevice, up));
      if (h.rest) PG_CUDA(cudaMemcpyAsync(sv.rest + roff, h.rest + roff, rbytes, cudaMemcpyHostToDevice, up));
    }
    PG_CUDA(cudaEventRecord(w->pipe_up_ev[c], up));
    PG_CUDA(cudaStreamWaitEvent(st, w->pipe_up_ev[c], 0));
    rc = pg_launch_sphere_range(w, dt, n_steps, first, count, st, w->d_pipe_counters + c, w->d_stats, &sv);
    if (rc) return rc;
    PG_CUDA(cudaEventRecord(w->pipe_k_ev[c], st));
    PG_CUDA(cudaStreamWaitEvent(down, w->pipe_k_ev[c], 0));
    if (h.packed) {
      PG_CUDA(cudaMemcpyAsync(h.pos + off, sv.pos + off, bytes, cudaMemcpyDeviceToHost, down));
    } else {
      PG_CUDA(cudaMemcpyAsync(h.pos + off, sv.pos + off, bytes, cudaMemcpyDeviceToHost, down));
      PG_CUDA(cudaMemcpyAsync(h.vel + off, sv.vel + off, bytes, cudaMemcpyDeviceToHost, down));
      if (h.rest) PG_CUDA(cudaMemcpyAsync(h.rest + roff, sv.rest + roff, rbytes, cudaMemcpyDeviceToHost, down));
    }
  }
  for (int k = 0; k < kDma; k++) {               // every forked queue joins
    PG_CUDA(cudaEventRecord(w->pipe_done[2 * k], w->pipe_down[k]));
    PG_CUDA(cudaStreamWaitEvent(w->stream, w->pipe_done[2 * k], 0));
    PG_CUDA(cudaEventRecord(w->pipe_done[2 * k + 1], w->pipe_up[k]));
    PG_CUDA(cudaStreamWaitEvent(w->stream, w->pipe_done[2 * k + 1], 0));
  }
  for (int k = 0; k < 3; k++) {
    PG_CUDA(cudaEventRecord(w->pipe_join[k], w->pipe_stream[k]));
    PG_CUDA(cudaStreamWaitEvent(w->stream, w->pipe_join[k], 0));
  }
  if (use_graph) {
    cudaGraph_t graph = nullptr;
    PG_CUDA(cudaStreamEndCapture(w->stream, &graph));
    PG_CUDA(cudaGraphInstantiate(&w->pipe_graph, graph, 0));
    cudaGraphDestroy(graph);
    w->pipe_key_ptr[0] = h.pos; w->pipe_key_ptr[1] = h.vel; w->pipe_key_ptr[2] = h.rest;
    w->pipe_key_dt = dt; w->pipe_key_steps = n_steps; w->pipe_key_chunks = kChunks; w->pipe_key_packed = h.packed ? 1 : 0;
    w->pipe_graph_kernels = w->launches - launches_before;
    PG_CUDA(cudaGraphLaunch(w->pipe_graph, w->stream));
  }
  PG_CUDA(cudaEventRecord(w->ev1, w->stream));
  w->last_steps = n_steps;
  w->cost_valid = false;
  w->stats_valid = true;
  PG_CUDA(cudaStreamSynchronize(w->stream));
  return PG_OK;
}

int pg_worlds_step_host(PgWorlds *w, float dt, int n_steps, float *pos, float *vel, uint8_t *at_rest) {
  if (!w || n_steps < 0 || !pos || !vel) { pg_set_error("pg_worlds_step_host: bad argument"); return PG_ERR_ARG; }
  if (w->mode != PG_MODE_SPHERES) { pg_set_error("pg_worlds_step_host: handle must be in sphere mode"); return PG_ERR_UNSUPPORTED; }
  PG_CUDA(cudaSetDevice(w->device));
  int n = 0;
  int rc = sphere_range_count(w, 0, w->v.n_worlds, &n);
  if (rc) return rc;
  HostState h;
  h.pos = (unsigned char *)pos; h.vel = (unsigned char *)vel; h.rest = at_rest;
  h.stride = (size_t)n * 12; h.rest_stride = (size_t)n; h.packed = false;
  return step_host_impl(w, dt, n_steps, h, n);
}

size_t pg_worlds_packed_world_bytes(const PgWorlds *w) {
  if (!w || w->mode != PG_MODE_SPHERES) return 0;
  const size_t n = (size_t)w->h_counts[1];
  return (n * 25 + 15) & ~(size_t)15;
}

int pg_worlds_step_host_packed(PgWorlds *w, float dt, int n_steps, void *state) {
  if (!w || n_steps < 0 || !state) { pg_set_error("pg_worlds_step_host_packed: bad argument"); return PG_ERR_ARG; }
  if (w->mode != PG_MODE_SPHERES) { pg_set_error("pg_worlds_step_host_packed: handle must be in sphere mode"); return PG_ERR_UNSUPPORTED; }
  PG_CUDA(cudaSetDevice(w->device));
  int n = 0;
  int rc = sphere_range_count(w, 0, w->v.n_worlds, &n);
  if (rc) return rc;
  HostState h;
  h.pos = (unsigned char *)state; h.vel = h.pos + (size_t)n * 12; h.rest = h.pos + (size_t)n * 24;
  h.stride = pg_worlds_packed_world_bytes(w); h.rest_stride = h.stride; h.packed = true;
  return step_host_impl(w, dt, n_steps, h, n);
}

int pg_worlds_sync(PgWorlds *w) {
  if (!w) return PG_ERR_ARG;
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  return PG_OK;
}

static inline long long event_key_row(const PgEvent &e) {
  const int pass = e.pad_ >> 8;
  const bool swapped = e.pad_ & 1;
  // the ROW body of the pass: player in passes 1-2, dynamic in 3-4 (world.c:180,277,369,473)
  int row, col;
  if (pass == 1 || pass == 2) { row = e.a_class == PG_CLASS_PLAYER ? e.a_index : e.b_index; col = e.a_class == PG_CLASS_PLAYER ? e.b_index : e.a_index; }
  else if (pass == 3) { row = e.a_class == PG_CLASS_DYNAMIC ? e.a_index : e.b_index; col = e.a_class == PG_CLASS_DYNAMIC ? e.b_index : e.a_index; }
  else { row = swapped ? e.b_index : e.a_index; col = swapped ? e.a_index : e.b_index; }
  return ((long long)row << 32) | (unsigned)col;
}

// restore solve order: step, pass, row, column (device slots are claimed atomically); pad_ -> pass number
static void sort_events(PgEvent *buf, int n) {
  std::stable_sort(buf, buf + n, [](const PgEvent &a, const PgEvent &b) {
    if (a.step != b.step) return a.step < b.step;
    const int pa = a.pad_ >> 8, pb = b.pad_ >> 8;
    if (pa != pb) return pa < pb;
    return event_key_row(a) < event_key_row(b);
  });
  for (int i = 0; i < n; i++) buf[i].pad_ >>= 8;
}

int pg_worlds_frame(PgWorlds *w, int world, const int32_t counts[3], const PgRange *ranges, int n_ranges,
                    const PgBody *records, float dt, int n_steps, int refresh_nodes,
                    PgBody *out_static, PgBody *out_dynamic, PgBody *out_player,
                    PgEvent *events, long long events_cap, long long *n_events, int *overflowed) {
  if (!valid_world(w, world) || !counts || n_ranges < 0 || (n_ranges > 0 && (!ranges || !records)) || n_steps < 0) { pg_set_error("pg_worlds_frame: bad argument"); return PG_ERR_ARG; }
  if (w->mode != PG_MODE_GENERAL) { pg_set_error("handle is in sphere mode"); return PG_ERR_ARG; }
  for (int c = 0; c < 3; c++) if (counts[c] < 0 || counts[c] > w->v.cap[c]) { pg_set_error("pg_worlds_frame: class %d: %d bodies > capacity %d", c, counts[c], w->v.cap[c]); return PG_ERR_CAPACITY; }
  size_t n_up = 0;
  for (int r = 0; r < n_ranges; r++) {
    const PgRange &g = ranges[r];
    if (g.cls < 0 || g.cls > 2 || g.first < 0 || g.count < 0 || g.first + g.count > counts[g.cls]) { pg_set_error("pg_worlds_frame: bad range %d", r); return PG_ERR_ARG; }
    n_up += (size_t)g.count;
  }
  static const bool frame_timing = getenv("PG_FRAME_TIMING") != nullptr;
  static double t_acc[5] = {0, 0, 0, 0, 0};
  static long n_acc = 0;
  auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
  const double t0 = frame_timing ? now() : 0.0;
  PG_CUDA(cudaSetDevice(w->device));
  int rc = ensure_records(w);
  if (rc) return rc;
  // pinned staging: [ dirty records in | records of the three classes out | event counter | events out ]
  const size_t cap_all = (size_t)w->v.cap[0] + w->v.cap[1] + w->v.cap[2];
  const size_t off_out = sizeof(PgBody) * cap_all, off_cnt = off_out + sizeof(PgBody) * cap_all, off_ev = off_cnt + 64;
  const size_t need = off_ev + sizeof(PgEvent) * (size_t)std::max(w->v.max_events, 1);
  if (need > w->h_frame_bytes) {
    if (w->h_frame) cudaFreeHost(w->h_frame);
    w->h_frame = nullptr; w->h_frame_bytes = 0;
    PG_CUDA(cudaHostAlloc((void **)&w->h_frame, need, cudaHostAllocDefault));
    w->h_frame_bytes = need;
    for (int c = 0; c < 3; c++) { free(w->h_types[c]); w->h_types[c] = (unsigned char *)calloc((size_t)w->v.cap[c], 1); }
  }
  if (n_up > cap_all) { pg_set_error("pg_worlds_frame: ranges overlap (more records than the world holds)"); return PG_ERR_ARG; }
  // A game-sized world's frame needs no copy at all: the kernel reads the dirty records out of the pinned block and
  // writes the records, the event count and the events back into it (SmallFrame; pinned memory is mapped into the device's
  // address space).  Larger worlds: one asynchronous copy per range and per class.
  const bool zero_copy = n_steps > 0 && n_ranges <= PG_FRAME_MAX_RANGES && pg_small_frame_fits(w, counts);
  // dirty records: one memcpy into the pinned block, one asynchronous copy per range out of it
  PgBody *h_in = reinterpret_cast<PgBody *>(w->h_frame);
  if (n_up) memcpy(h_in, records, sizeof(PgBody) * n_up);
  SmallFrame fr;
  memset(&fr, 0, sizeof(fr));
  fr.world = world;
  size_t at = 0;
  for (int r = 0; r < n_ranges; r++) {
    const PgRange &g = ranges[r];
    if (zero_copy) { fr.cls[r] = g.cls; fr.first[r] = g.first; fr.count[r] = g.count; }
    if (g.count == 0) continue;
    if (!zero_copy)
      PG_CUDA(cudaMemcpyAsync(w->v.bodies[g.cls] + (size_t)world * w->v.cap[g.cls] + g.first, h_in + at, sizeof(PgBody) * (size_t)g.count,
                              cudaMemcpyHostToDevice, w->stream));
    for (int i = 0; i < g.count; i++) w->h_types[g.cls][g.first + i] = (unsigned char)h_in[at + i].collider_type;
    at += (size_t)g.count;
  }
  bool counts_changed = false;
  for (int c = 0; c < 3; c++) if (w->h_counts[world * 3 + c] != counts[c]) { w->h_counts[world * 3 + c] = counts[c]; counts_changed = true; }
  int *h_cnt = reinterpret_cast<int *>(w->h_frame + off_cnt);
  if (counts_changed && !zero_copy) {
    for (int c = 0; c < 3; c++) h_cnt[4 + c] = counts[c];
    PG_CUDA(cudaMemcpyAsync(w->v.counts + world * 3, h_cnt + 4, 3 * sizeof(int), cudaMemcpyHostToDevice, w->stream));
  }
  // pass 3 may run one thread per dynamic body only if no resolver can write a static body (see pg_general.cu)
  bool ok = true;
  for (int i = 0; i < counts[0] && ok; i++) ok = w->h_types[0][i] == PG_AABB || w->h_types[0][i] == PG_PLANE;
  for (int i = 0; i < counts[1] && ok; i++) ok = w->h_types[1][i] != PG_PLANE;
  w->static_pass_parallel_ok = ok && w->v.n_worlds == 1;
  if (w->v.n_worlds > 1) { rc = refresh_parallel_flag(w); if (rc) return rc; }
  const double t1 = frame_timing ? now() : 0.0;
  PgBody *h_out = reinterpret_cast<PgBody *>(w->h_frame + off_out);
  PgEvent *h_ev = reinterpret_cast<PgEvent *>(w->h_frame + off_ev);
  if (!zero_copy) PG_CUDA(cudaMemsetAsync(w->v.n_events, 0, sizeof(int) * (size_t)w->v.n_worlds, w->stream));
  PG_CUDA(cudaEventRecord(w->ev0, w->stream));
  {
    float cdt = dt;
    if ((double)cdt > 0.01666) cdt = (float)0.01666;
    pg_leaf_lengths(0.0f, cdt, &w->v.leaf_lo, &w->v.leaf_hi);
  }
  if (zero_copy) {
    for (int c = 0; c < 3; c++) fr.counts[c] = counts[c];
    fr.n_ranges = n_ranges;
    void *d_frame = nullptr;
    PG_CUDA(cudaHostGetDevicePointer(&d_frame, w->h_frame, 0));
    unsigned char *db = static_cast<unsigned char *>(d_frame);
    fr.in = reinterpret_cast<const PgBody *>(db);
    fr.out = reinterpret_cast<PgBody *>(db + off_out);
    fr.cnt_out = reinterpret_cast<int *>(db + off_cnt);
    fr.ev_out = reinterpret_cast<PgEvent *>(db + off_ev);
    rc = pg_launch_general_step(w, dt, n_steps, refresh_nodes, &fr);
    if (rc) return rc;
  } else if (n_steps > 0) { rc = pg_launch_general_step(w, dt, n_steps, refresh_nodes); if (rc) return rc; }
  PG_CUDA(cudaEventRecord(w->ev1, w->stream));
  w->last_steps = n_steps;
  const double t2 = frame_timing ? now() : 0.0;
  // everything back into the pinned block, then ONE synchronisation
  PgBody *outs[3] = {out_static, out_dynamic, out_player};
  size_t out_at[3] = {0, (size_t)w->v.cap[0], (size_t)w->v.cap[0] + w->v.cap[1]};
  if (!zero_copy) {
    for (int c = 0; c < 3; c++)
      if (outs[c] && counts[c] > 0)
        PG_CUDA(cudaMemcpyAsync(h_out + out_at[c], w->v.bodies[c] + (size_t)world * w->v.cap[c], sizeof(PgBody) * (size_t)counts[c], cudaMemcpyDeviceToHost, w->stream));
    PG_CUDA(cudaMemcpyAsync(h_cnt, w->v.n_events + world, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    if (w->v.max_events > 0 && events)
      PG_CUDA(cudaMemcpyAsync(h_ev, w->v.events + (size_t)world * w->v.max_events, sizeof(PgEvent) * (size_t)w->v.max_events, cudaMemcpyDeviceToHost, w->stream));
  }
  const double t3 = frame_timing ? now() : 0.0;
  PG_CUDA(cudaStreamSynchronize(w->stream));
  const double t4 = frame_timing ? now() : 0.0;
  if (frame_timing) {
    float kms = 0.0f;
    cudaEventElapsedTime(&kms, w->ev0, w->ev1);
    t_acc[0] += t1 - t0; t_acc[1] += t2 - t1; t_acc[2] += t3 - t2; t_acc[3] += t4 - t3; t_acc[4] += kms * 1e3;
    if (++n_acc % 1000 == 0)
      fprintf(stderr, "[pg_worlds_frame] per call (us): upload issue %.1f, launch %.1f, download issue %.1f, sync wait %.1f; kernel %.1f\n",
              t_acc[0] / n_acc, t_acc[1] / n_acc, t_acc[2] / n_acc, t_acc[3] / n_acc, t_acc[4] / n_acc);
  }
  for (int c = 0; c < 3; c++) if (outs[c] && counts[c] > 0) memcpy(outs[c], h_out + out_at[c], sizeof(PgBody) * (size_t)counts[c]);
  int n = h_cnt[0], over = 0;
  if (n > w->v.max_events) { over = 1; n = w->v.max_events; }
  if (events && n > 0) {
    sort_events(h_ev, n);
    for (int i = 0; i < n && i < events_cap; i++) events[i] = h_ev[i];
  }
  if (n_events) *n_events = n;
  if (overflowed) *overflowed = over;
  return PG_OK;
}

long long pg_worlds_events(PgWorlds *w, PgEvent *out, long long cap, int *overflowed) {
  if (!w) { pg_set_error("pg_worlds_events: bad handle"); return PG_ERR_ARG; }
  if (cudaSetDevice(w->device) != cudaSuccess || cudaStreamSynchronize(w->stream) != cudaSuccess) { pg_set_error("pg_worlds_events: cuda failure"); return PG_ERR_CUDA; }
  std::vector<int> cnt(w->v.n_worlds);
  if (cudaMemcpy(cnt.data(), w->v.n_events, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost) != cudaSuccess) { pg_set_error("pg_worlds_events: copy failed"); return PG_ERR_CUDA; }
  long long total = 0;
  int over = 0;
  std::vector<PgEvent> buf;
  for (int wd = 0; wd < w->v.n_worlds; wd++) {
    int n = cnt[wd];
    if (n > w->v.max_events) { over = 1; n = w->v.max_events; }
    if (n > 0 && out) {
      buf.resize(n);
      if (cudaMemcpy(buf.data(), w->v.events + (size_t)wd * w->v.max_events, sizeof(PgEvent) * n, cudaMemcpyDeviceToHost) != cudaSuccess) { pg_set_error("pg_worlds_events: copy failed"); return PG_ERR_CUDA; }
      sort_events(buf.data(), n);
      for (int i = 0; i < n; i++) if (total + i < cap) out[total + i] = buf[i];
    }
    total += n;
  }
  if (overflowed) *overflowed = over;
  return total;
}

int pg_worlds_stats(PgWorlds *w, PgStats *host_out, void *dev_out8) {
  if (!w) return PG_ERR_ARG;
  PG_CUDA(cudaSetDevice(w->device));
  if (w->mode == PG_MODE_GENERAL) { int rc = ensure_records(w); if (rc) return rc; }
  int rc = pg_launch_worlds_stats(w);
  if (rc) return rc;
  if (dev_out8) PG_CUDA(cudaMemcpyAsync(dev_out8, w->d_stats, 8 * sizeof(double), cudaMemcpyDeviceToDevice, w->stream));
  if (host_out) {
    double h[8];
    PG_CUDA(cudaMemcpyAsync(h, w->d_stats, sizeof(h), cudaMemcpyDeviceToHost, w->stream));
    PG_CUDA(cudaStreamSynchronize(w->stream));
    host_out->bodies = h[0]; host_out->kinetic = h[1];
    host_out->momentum[0] = h[2]; host_out->momentum[1] = h[3]; host_out->momentum[2] = h[4];
    host_out->contacts = h[5]; host_out->at_rest = h[6]; host_out->nonfinite = h[7];
  }
  return PG_OK;
}

long long pg_worlds_launches(const PgWorlds *w) { return w ? w->launches : 0; }

float pg_worlds_last_step_ms(PgWorlds *w) {
  if (!w) return -1.0f;
  float ms = -1.0f;
  if (cudaSetDevice(w->device) != cudaSuccess) return -1.0f;
  if (cudaEventSynchronize(w->ev1) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, w->ev0, w->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

void *pg_worlds_stream(PgWorlds *w) { return w ? (void *)w->stream : nullptr; }

}  // extern "C"
