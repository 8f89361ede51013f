// pg_general.cu -- the Crux physics step for arbitrary collider mixes, reference solve order kept.
//
// One thread block per world.  physics_step (/root/reference/src/physics/world.c:174-555) is four
// nested loops that mutate bodies in place; a visit (A,B) only ever changes A and B.  The block
// reproduces the loops row by row: for row body R the threads test all columns in parallel on the
// current state, the LOWEST hitting column is resolved by the thread that found it (exactly the
// visit the sequential loop would have reached first), and the columns after it are re-tested
// against R's new state.  Rows themselves stay sequential.  Pass 3 (dynamic x static) runs one
// thread per dynamic body when no static body can be written by a resolver.
//
// This kernel is the drop-in path for game-sized scenes (<= a few hundred bodies per world, any
// collider type).  Sphere-only batches use pg_spheres.cu; one huge world uses pg_large.cu.
#include "pg_common.cuh"
#include "pg_math.cuh"

#include <climits>

namespace {

constexpr int kThreads = 128;

struct EventSink {
  PgEvent *buf;
  int *counter;
  int cap;
  int world;
  int step;
};

// pad_ carries what the host needs to restore solve order: bit0 = "A/B were swapped into collider
// type order", bits 8.. = pass number (1..4).
__device__ void emit(const EventSink &s, int pass, bool swapped, int etype, int ac, int ai, int bc, int bi) {
  int slot = atomicAdd(s.counter, 1);
  if (slot >= s.cap) return;
  PgEvent e;
  e.world = s.world; e.step = s.step; e.event_type = etype;
  e.a_class = ac; e.a_index = ai; e.b_class = bc; e.b_index = bi;
  e.pad_ = (pass << 8) | (swapped ? 1 : 0);
  s.buf[slot] = e;
}

// Read-only half of a visit: would (x, y) fire?  (world.c:481-502)
__device__ bool test_visit(const PgBody &x, const PgBody &y, float dt, float *t_hit) {
  const bool swap = x.collider_type > y.collider_type;
  const PgBody &A = swap ? y : x;
  const PgBody &B = swap ? x : y;
  float h;
  if (!pgm::interval_hit(A, B, 0.0f, dt, &h)) return false;
  bool has;
  pgm::Contact c = pgm::narrow(A, B, dt, &has);
  if (!has || !(c.hit && c.t >= 0)) return false;
  *t_hit = c.t;
  return true;
}

// Mutating half (world.c:503-541): behaviour lookup, resolution, event.
__device__ void commit_visit(PgBody &x, int xc, int xi, PgBody &y, int yc, int yi, float t,
                             const EventSink &sink, int pass) {
  const bool swap = x.collider_type > y.collider_type;
  PgBody &A = swap ? y : x;
  PgBody &B = swap ? x : y;
  if (pgm::collision_behavior(A.entity_type, B.entity_type) == 0) pgm::resolve(A, B, t);
  emit(sink, pass, swap, pgm::event_type(A.entity_type, B.entity_type),
       swap ? yc : xc, swap ? yi : xi, swap ? xc : yc, swap ? xi : yi);
}

// One row of a sequential pass: row body vs columns [0, ncols) \ {skip}, in increasing column order.
__device__ void sweep_row(PgBody *row, int rc, int ri, PgBody *cols, int cc, int ncols, int skip,
                          float dt, const EventSink &sink, int pass, int *s_min) {
  int cursor = 0;
  for (;;) {
    if (threadIdx.x == 0) *s_min = INT_MAX;
    __syncthreads();
    int found = INT_MAX;
    float ft = 0.0f;
    for (int j = cursor + (int)threadIdx.x; j < ncols; j += kThreads) {
      if (j == skip) continue;
      float t;
      if (test_visit(*row, cols[j], dt, &t)) { found = j; ft = t; break; }
    }
    if (found != INT_MAX) atomicMin(s_min, found);
    __syncthreads();
    const int jmin = *s_min;
    if (jmin == INT_MAX) break;
    if (found == jmin) commit_visit(*row, rc, ri, cols[jmin], cc, jmin, ft, sink, pass);
    __syncthreads();
    cursor = jmin + 1;
  }
}

__device__ void integrate_body(PgBody &b, float dt) {   // world.c:362-366, 549-553
  if (b.at_rest) return;
  pgm::V3 p = pgm::v3p(b.position), v = pgm::v3p(b.velocity);
  pgm::integrate(p, v, dt);
  pgm::store(b.position, p);
  pgm::store(b.velocity, v);
}

__global__ void __launch_bounds__(kThreads)
k_step_general(WorldsView v, float dt_in, int n_steps, int refresh_nodes, int static_parallel_ok) {
  __shared__ int s_min;
  const int w = blockIdx.x;
  if (w >= v.n_worlds) return;
  PgBody *S = v.bodies[0] + (size_t)w * v.cap[0];
  PgBody *D = v.bodies[1] + (size_t)w * v.cap[1];
  PgBody *P = v.bodies[2] + (size_t)w * v.cap[2];
  const int ns = v.counts[w * 3 + 0], nd = v.counts[w * 3 + 1], np = v.counts[w * 3 + 2];
  const float dt = pgm::clamp_dt(dt_in);
  EventSink sink;
  sink.buf = v.events + (size_t)w * v.max_events;
  sink.counter = v.n_events + w;
  sink.cap = v.max_events;
  sink.world = w;

  for (int step = 0; step < n_steps; step++) {
    sink.step = step;
    // pass 1: players x statics (world.c:180-274)
    for (int i = 0; i < np; i++) sweep_row(&P[i], PG_CLASS_PLAYER, i, S, PG_CLASS_STATIC, ns, -1, dt, sink, 1, &s_min);
    // pass 2: players x dynamics, then integrate the player (world.c:277-367)
    for (int i = 0; i < np; i++) {
      sweep_row(&P[i], PG_CLASS_PLAYER, i, D, PG_CLASS_DYNAMIC, nd, -1, dt, sink, 2, &s_min);
      if (threadIdx.x == 0) integrate_body(P[i], dt);
      __syncthreads();
    }
    // pass 3: dynamics x statics (world.c:369-471)
    if (static_parallel_ok) {
      for (int i = threadIdx.x; i < nd; i += kThreads) {
        for (int j = 0; j < ns; j++) {
          float t;
          if (test_visit(D[i], S[j], dt, &t)) commit_visit(D[i], PG_CLASS_DYNAMIC, i, S[j], PG_CLASS_STATIC, j, t, sink, 3);
        }
      }
      __syncthreads();
    } else {
      for (int i = 0; i < nd; i++) sweep_row(&D[i], PG_CLASS_DYNAMIC, i, S, PG_CLASS_STATIC, ns, -1, dt, sink, 3, &s_min);
    }
    // pass 4: dynamics x dynamics, integrate row i after its sweep (world.c:473-554)
    for (int i = 0; i < nd; i++) {
      sweep_row(&D[i], PG_CLASS_DYNAMIC, i, D, PG_CLASS_DYNAMIC, nd, i, dt, sink, 4, &s_min);
      if (threadIdx.x == 0) integrate_body(D[i], dt);
      __syncthreads();
    }
    // scene_node_update of every node (scene.c:1051-1079), as scene_update does after the step
    if (refresh_nodes) {
      for (int i = threadIdx.x; i < ns; i += kThreads) if (S[i].has_node) pgm::node_refresh(S[i]);
      for (int i = threadIdx.x; i < nd; i += kThreads) if (D[i].has_node) pgm::node_refresh(D[i]);
      for (int i = threadIdx.x; i < np; i += kThreads) if (P[i].has_node) pgm::node_refresh(P[i]);
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------- per-pair kernels (pg_pair_*)

__global__ void k_pair_max_move(const PgBody *a, int n, float t0, float t1, float *out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = pgm::max_move(a[k], t0, t1);
}
__global__ void k_pair_min_dist(const PgBody *a, const PgBody *b, int n, const float *t, float *out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = pgm::min_dist(a[k], b[k], t[k]);
}
__global__ void k_pair_interval(const PgBody *a, const PgBody *b, int n, float t0, float t1, int *hit, float *ht) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  float h = -1.0f;
  hit[k] = pgm::interval_hit(a[k], b[k], t0, t1, &h) ? 1 : 0;
  ht[k] = h;
}
__global__ void k_pair_narrow(const PgBody *a, const PgBody *b, int n, float dt, int *col, float *ht) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  bool has;
  pgm::Contact c = pgm::narrow(a[k], b[k], dt, &has);
  col[k] = has ? (c.hit ? 1 : 0) : -1;
  ht[k] = has ? c.t : 0.0f;
}
__global__ void k_pair_resolve(PgBody *a, PgBody *b, int n, const float *ht, float) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) pgm::resolve(a[k], b[k], ht[k]);
}
__global__ void k_pair_visit(PgBody *a, PgBody *b, int n, float dt, int *ev) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const bool swap = a[k].collider_type > b[k].collider_type;
  PgBody &A = swap ? b[k] : a[k];
  PgBody &B = swap ? a[k] : b[k];
  ev[k] = pgm::visit_ordered(A, B, dt) ? 1 : 0;
}

// The two collider pairs the reference leaves as empty functions (distance.c:301-303,323-325;
// narrow_phase.c:522-524,590-592; resolution.c:572-574,624-626): their table entries return undefined values
// upstream, so a direct table-level query is rejected at the boundary.  Inside a step such
// a visit never fires (pgm::visit_ordered), which pg_pair_visit reports as "no event".
bool undefined_pair(const PgBody *a, const PgBody *b, int n) {
  if (!a || !b) return false;
  for (int k = 0; k < n; k++) {
    const int ta = a[k].collider_type, tb = b[k].collider_type;
    if ((ta == PG_CAPSULE && (tb == PG_SPHERE || tb == PG_CAPSULE)) || (ta == PG_SPHERE && tb == PG_CAPSULE)) {
      pg_set_error("pair %d: sphere-capsule / capsule-capsule is undefined in the reference (empty functions)", k);
      return true;
    }
  }
  return false;
}

template <typename F>
int with_pairs(const PgBody *a, const PgBody *b, int n, PgBody **da, PgBody **db, F &&body) {
  *da = nullptr; *db = nullptr;
  if (n <= 0) return PG_ERR_ARG;
  PG_CUDA(cudaMalloc(da, sizeof(PgBody) * (size_t)n));
  PG_CUDA(cudaMemcpy(*da, a, sizeof(PgBody) * (size_t)n, cudaMemcpyHostToDevice));
  if (b) {
    PG_CUDA(cudaMalloc(db, sizeof(PgBody) * (size_t)n));
    PG_CUDA(cudaMemcpy(*db, b, sizeof(PgBody) * (size_t)n, cudaMemcpyHostToDevice));
  }
  int rc = body();
  cudaFree(*da);
  if (*db) cudaFree(*db);
  return rc;
}

template <typename T>
struct DevBuf {
  T *p = nullptr;
  int alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * n) == cudaSuccess ? 0 : -1; }
  ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

int pg_launch_general_step(PgWorlds *w, float dt, int n_steps, int refresh_nodes) {
  k_step_general<<<w->v.n_worlds, kThreads, 0, w->stream>>>(w->v, dt, n_steps, refresh_nodes,
                                                             w->static_pass_parallel_ok ? 1 : 0);
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}

extern "C" {

int pg_pair_max_move(const PgBody *a, int n, float t0, float t1, float *out) {
  PgBody *da, *db;
  return with_pairs(a, nullptr, n, &da, &db, [&]() -> int {
    DevBuf<float> o; if (o.alloc(n)) return PG_ERR_CUDA;
    k_pair_max_move<<<(n + 127) / 128, 128>>>(da, n, t0, t1, o.p);
    PG_CUDA(cudaMemcpy(out, o.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return PG_OK;
  });
}

int pg_pair_min_dist(const PgBody *a, const PgBody *b, int n, const float *t, float *out) {
  if (n > 0 && undefined_pair(a, b, n)) return PG_ERR_UNSUPPORTED;
  PgBody *da, *db;
  return with_pairs(a, b, n, &da, &db, [&]() -> int {
    DevBuf<float> o, dt_; if (o.alloc(n) || dt_.alloc(n)) return PG_ERR_CUDA;
    PG_CUDA(cudaMemcpy(dt_.p, t, sizeof(float) * n, cudaMemcpyHostToDevice));
    k_pair_min_dist<<<(n + 127) / 128, 128>>>(da, db, n, dt_.p, o.p);
    PG_CUDA(cudaMemcpy(out, o.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return PG_OK;
  });
}

int pg_pair_interval_collision(const PgBody *a, const PgBody *b, int n, float t0, float t1,
                               int32_t *out_hit, float *out_hit_time) {
  if (n > 0 && undefined_pair(a, b, n)) return PG_ERR_UNSUPPORTED;
  PgBody *da, *db;
  return with_pairs(a, b, n, &da, &db, [&]() -> int {
    DevBuf<int> h; DevBuf<float> t; if (h.alloc(n) || t.alloc(n)) return PG_ERR_CUDA;
    k_pair_interval<<<(n + 127) / 128, 128>>>(da, db, n, t0, t1, h.p, t.p);
    PG_CUDA(cudaMemcpy(out_hit, h.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    PG_CUDA(cudaMemcpy(out_hit_time, t.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return PG_OK;
  });
}

int pg_pair_narrow_phase(const PgBody *a, const PgBody *b, int n, float dt, int32_t *out_colliding,
                         float *out_hit_time) {
  if (n > 0 && undefined_pair(a, b, n)) return PG_ERR_UNSUPPORTED;
  PgBody *da, *db;
  return with_pairs(a, b, n, &da, &db, [&]() -> int {
    DevBuf<int> h; DevBuf<float> t; if (h.alloc(n) || t.alloc(n)) return PG_ERR_CUDA;
    k_pair_narrow<<<(n + 127) / 128, 128>>>(da, db, n, dt, h.p, t.p);
    PG_CUDA(cudaMemcpy(out_colliding, h.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    PG_CUDA(cudaMemcpy(out_hit_time, t.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return PG_OK;
  });
}

int pg_pair_resolve(PgBody *a, PgBody *b, int n, const float *hit_time, float dt) {
  if (n > 0 && undefined_pair(a, b, n)) return PG_ERR_UNSUPPORTED;
  PgBody *da, *db;
  return with_pairs(a, b, n, &da, &db, [&]() -> int {
    DevBuf<float> t; if (t.alloc(n)) return PG_ERR_CUDA;
    PG_CUDA(cudaMemcpy(t.p, hit_time, sizeof(float) * n, cudaMemcpyHostToDevice));
    k_pair_resolve<<<(n + 127) / 128, 128>>>(da, db, n, t.p, dt);
    PG_CUDA(cudaMemcpy(a, da, sizeof(PgBody) * (size_t)n, cudaMemcpyDeviceToHost));
    PG_CUDA(cudaMemcpy(b, db, sizeof(PgBody) * (size_t)n, cudaMemcpyDeviceToHost));
    return PG_OK;
  });
}

int pg_pair_visit(PgBody *a, PgBody *b, int n, float dt, int32_t *out_event) {
  PgBody *da, *db;
  return with_pairs(a, b, n, &da, &db, [&]() -> int {
    DevBuf<int> e; if (e.alloc(n)) return PG_ERR_CUDA;
    k_pair_visit<<<(n + 127) / 128, 128>>>(da, db, n, dt, e.p);
    PG_CUDA(cudaMemcpy(out_event, e.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    PG_CUDA(cudaMemcpy(a, da, sizeof(PgBody) * (size_t)n, cudaMemcpyDeviceToHost));
    PG_CUDA(cudaMemcpy(b, db, sizeof(PgBody) * (size_t)n, cudaMemcpyDeviceToHost));
    return PG_OK;
  });
}

}  // extern "C"
