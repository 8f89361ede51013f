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

#include <algorithm>
#include <climits>
#include <cstdlib>

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

// ---------------------------------------------------------------- game-sized worlds: speculate, then commit in order
// k_step_general walks the sweep row by row with two block barriers per row: 140 us per physics_step for the 15
// bodies of scenes/bouncehouse.json, nearly all of it barrier and load latency.  A game-sized world has a few hundred
// visits per step and about one contact, so: evaluate EVERY visit of the step at once, one per thread, on the state
// it will see if no contact intervenes (a body integrated or not according to where its own row ends in the sweep);
// then find the earliest visit that fires, bring the real state up to it (the integrations that lie before it),
// commit it exactly as the sequential loop would, and re-evaluate only the later visits of the two bodies it
// changed.  One round per contact; a step without contacts is one round.  Same visits, same order, same arithmetic:
// the result is the sequential one.
constexpr int kSmallMaxVisits = 2048;
constexpr int kSmallBudget = 48;         // nodes of interval_collision's recursion a single thread explores before the visit goes to the block
                                         // (bouncehouse.json, us per frame over 3000 / 6000 frames: 6 -> 86 / 203, 24 -> 63 / 98, 48 -> 60 / 94, 96 -> 58 / 96, 256 -> 60 / 103)
constexpr int kSmallStageBodies = 48;     // worlds of at most this many bodies keep their records in shared memory (14.6 KB)
constexpr int kSmallStack = 2048;        // open nodes the block-wide search can hold

// interval_collision (world.c:557-582) with a node budget: 0 = false, 1 = true, 2 = still undecided after
// `budget` nodes.  Same tree, same end-point tests as pgm::interval_hit.
__device__ int interval_hit_budget(const PgBody &A, const PgBody &B, float t_begin, float t_end, int budget, const pgm::BoxPre *na = nullptr,
                                   const pgm::BoxPre *nb = nullptr) {
  unsigned long long path = 0;
  int depth = 0;
  float t0 = t_begin, t1 = t_end;
  const pgm::PairPre pre = pgm::pair_pre(A, B, na, nb);      // everything of the distance function that does not depend on the time
  // The left child of a node shares its t0: the distance there is carried down instead of evaluated again (a touching
  // pair is found by walking straight down the left edge of the tree, where that is every second evaluation).
  float d0 = 0.0f;
  bool have_d0 = false;
  for (;;) {
    if (--budget < 0) return 2;
    const float len = t1 - t0;
    const float mm = pre.nA * len + pre.nB * len;       // maximum_object_movement_over_time of both bodies (world.c:559-560)
    if (!have_d0) d0 = pgm::min_dist_pre(A, B, pre, t0);
    const bool open = !(d0 > mm) && !(pgm::min_dist_pre(A, B, pre, t1) > mm);
    if (open) {
      if (t1 - t0 < pgm::kIntervalEps) return 1;
      if (depth >= 62) return 0;
      path <<= 1; depth++;
      t1 = (t0 + t1) * 0.5f;
      have_d0 = true;
      continue;
    }
    have_d0 = false;
    while (depth > 0 && (path & 1ull)) { path >>= 1; depth--; }
    if (depth == 0) return 0;
    path |= 1ull;
    t0 = t_begin; t1 = t_end;
    for (int k = depth - 1; k >= 0; k--) {
      const float mid = (t0 + t1) * 0.5f;
      if ((path >> k) & 1ull) t0 = mid; else t1 = mid;
    }
  }
}

// The same search by the whole block (all threads, identical arguments).  Only the boolean "some leaf survives with
// all its ancestors" is consumed (hit_time is never read, world.c:196-198), so the ORDER in which the tree is
// explored is free: the open nodes sit on a shared stack, every round the threads take the top ones (deepest
// first), test them exactly as the recursion would -- end points re-derived from the root by the same
// (t0+t1)*0.5f sequence -- and push the children of those that pass.  A grazing pair's hundreds to thousands of
// nodes take tens of rounds instead of one thread's milliseconds.
struct BlockSearch {
  unsigned long long stack[kSmallStack];          // node = depth << 56 | path
  int top, found, overflow;
};
__device__ bool interval_hit_block(const PgBody &A, const PgBody &B, float t_begin, float t_end, BlockSearch &bs, const pgm::BoxPre *na = nullptr,
                                   const pgm::BoxPre *nb = nullptr) {
  __syncthreads();
  if (threadIdx.x == 0) { bs.stack[0] = 0ull; bs.top = 1; bs.found = 0; bs.overflow = 0; }
  __syncthreads();
  const pgm::PairPre pre = pgm::pair_pre(A, B, na, nb);
  for (;;) {
    const int n = bs.top;
    const bool stop = n == 0 || bs.found || bs.overflow;
    __syncthreads();
    if (stop) break;
    const int take = min(n, (int)blockDim.x), base = n - take;
    unsigned long long node = 0ull;
    if ((int)threadIdx.x < take) node = bs.stack[base + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) bs.top = base;
    __syncthreads();
    if ((int)threadIdx.x < take) {
      const int depth = (int)(node >> 56);
      const unsigned long long path = node & 0x00ffffffffffffffull;
      float t0 = t_begin, t1 = t_end;
      for (int k = depth - 1; k >= 0; k--) {
        const float mid = (t0 + t1) * 0.5f;
        if ((path >> k) & 1ull) t0 = mid; else t1 = mid;
      }
      if (pgm::interval_node_open(A, B, pre, t0, t1)) {
        if (t1 - t0 < pgm::kIntervalEps) atomicExch(&bs.found, 1);
        else if (depth >= 55) atomicExch(&bs.overflow, 1);
        else {
          const int pos = atomicAdd(&bs.top, 2);
          if (pos + 2 <= kSmallStack) {
            bs.stack[pos] = ((unsigned long long)(depth + 1) << 56) | (path << 1) | 1ull;      // right child below,
            bs.stack[pos + 1] = ((unsigned long long)(depth + 1) << 56) | (path << 1);        // left child on top
          } else atomicExch(&bs.overflow, 1);
        }
      }
    }
    __syncthreads();
  }
  const bool overflow = bs.overflow != 0, found = bs.found != 0;
  __syncthreads();
  if (overflow) { float h; return pgm::interval_hit(A, B, t_begin, t_end, &h); }      // (every thread alone: slow, but the same answer)
  return found;
}

struct SmallPlan {
  int ns, nd, np;
  int v1, v2, v3, v4;          // visits per pass (world.c:180, 277, 369, 473)
  __device__ int total() const { return v1 + v2 + v3 + v4; }
  // visit k of the sweep -> row body (class, index), column body (class, index), pass
  __device__ void decode(int k, int &rc, int &ri, int &cc, int &ci, int &pass) const {
    if (k < v1) { pass = 1; rc = PG_CLASS_PLAYER; ri = k / ns; cc = PG_CLASS_STATIC; ci = k % ns; return; }
    k -= v1;
    if (k < v2) { pass = 2; rc = PG_CLASS_PLAYER; ri = k / nd; cc = PG_CLASS_DYNAMIC; ci = k % nd; return; }
    k -= v2;
    if (k < v3) { pass = 3; rc = PG_CLASS_DYNAMIC; ri = k / ns; cc = PG_CLASS_STATIC; ci = k % ns; return; }
    k -= v3;
    pass = 4; rc = PG_CLASS_DYNAMIC; cc = PG_CLASS_DYNAMIC;
    ri = k / (nd - 1);
    const int jj = k % (nd - 1);
    ci = jj + (jj >= ri ? 1 : 0);
  }
  // index of the first visit AFTER body (cls, i) has been integrated by its own row (world.c:362-366, 549-553);
  // statics never are
  __device__ int int_point(int cls, int i) const {
    if (cls == PG_CLASS_PLAYER) return v1 + (i + 1) * nd;
    if (cls == PG_CLASS_DYNAMIC) return v1 + v2 + v3 + (i + 1) * (nd - 1);
    return 0x7fffffff;
  }
};

__device__ __forceinline__ PgBody *small_body(PgBody *S, PgBody *D, PgBody *P, int cls, int i) {
  return cls == PG_CLASS_STATIC ? S + i : (cls == PG_CLASS_DYNAMIC ? D + i : P + i);
}

// Read-only half of a visit (world.c:481-502) with the interval search either bounded (bs == nullptr: one thread,
// 2 = undecided) or run by the whole block (bs != nullptr: all threads, identical arguments).
// px / py: box_pre_node of x / y where the kernel holds it (null otherwise)
__device__ int test_visit_small(const PgBody &x, const PgBody &y, float dt, float *t_hit, BlockSearch *bs, const pgm::BoxPre *px = nullptr,
                                const pgm::BoxPre *py = nullptr) {
  const bool swap = x.collider_type > y.collider_type;
  const PgBody &A = swap ? y : x;
  const PgBody &B = swap ? x : y;
  const pgm::BoxPre *na = swap ? py : px, *nb = swap ? px : py;
  if (bs) { if (!interval_hit_block(A, B, 0.0f, dt, *bs, na, nb)) return 0; }
  else {
    const int r = interval_hit_budget(A, B, 0.0f, dt, kSmallBudget, na, nb);
    if (r != 1) return r;
  }
  bool has;
  pgm::Contact c = pgm::narrow(A, B, dt, &has);
  if (!has || !(c.hit && c.t >= 0)) return 0;
  *t_hit = c.t;
  return 1;
}

// Would visit k fire, on the real state (integrations with point <= `applied` done) plus the integrations that lie
// between `applied` and k?  0 / 1, or 2 = undecided (bs == nullptr only).
__device__ int small_test(const SmallPlan &pl, PgBody *S, PgBody *D, PgBody *P, int k, int applied, float dt, float *t_hit, BlockSearch *bs,
                          PgBody *s_tmp, const pgm::BoxPre *node_pre) {
  int rc, ri, cc, ci, pass;
  pl.decode(k, rc, ri, cc, ci, pass);
  const PgBody *row = small_body(S, D, P, rc, ri);        // a row body is never integrated before its own visits
  const PgBody *col = small_body(S, D, P, cc, ci);
  // the bodies' boxes through their nodes, built once per step (the node matrix does not change during a step, and the
  // copy of a column advanced by its own row keeps it)
  auto slot = [&](int cls, int i) { return cls == PG_CLASS_STATIC ? i : (cls == PG_CLASS_DYNAMIC ? pl.ns + i : pl.ns + pl.nd + i); };
  const pgm::BoxPre *prow = node_pre && row->collider_type == PG_AABB ? node_pre + slot(rc, ri) : nullptr;
  const pgm::BoxPre *pcol = node_pre && col->collider_type == PG_AABB ? node_pre + slot(cc, ci) : nullptr;
  const int ip = pl.int_point(cc, ci);
  if (ip > applied && ip <= k && !col->at_rest) {
    // the column as its own row's integration will leave it
    if (bs) {
      __syncthreads();
      if (threadIdx.x == 0) { *s_tmp = *col; integrate_body(*s_tmp, dt); }
      __syncthreads();
      return test_visit_small(*row, *s_tmp, dt, t_hit, bs, prow, pcol);
    }
    PgBody tmp = *col;
    integrate_body(tmp, dt);
    return test_visit_small(*row, tmp, dt, t_hit, nullptr, prow, pcol);
  }
  return test_visit_small(*row, *col, dt, t_hit, bs, prow, pcol);
}

__global__ void __launch_bounds__(kThreads)
k_step_small(WorldsView v, float dt_in, int n_steps, int refresh_nodes, SmallFrame fr) {
  __shared__ unsigned char s_fire[kSmallMaxVisits];      // 0 / 1, 2 = the single-thread search ran out of budget
  __shared__ float s_t[kSmallMaxVisits];
  __shared__ int s_min;
  __shared__ BlockSearch s_bs;
  __shared__ PgBody s_tmp;
  const int w = blockIdx.x;
  if (w >= v.n_worlds) return;
  PgBody *S = v.bodies[0] + (size_t)w * v.cap[0];
  PgBody *D = v.bodies[1] + (size_t)w * v.cap[1];
  PgBody *P = v.bodies[2] + (size_t)w * v.cap[2];
  SmallPlan pl;
  pl.ns = v.counts[w * 3 + 0]; pl.nd = v.counts[w * 3 + 1]; pl.np = v.counts[w * 3 + 2];
  const bool framed = fr.world == w;       // this launch is one frame of the drop-in: no copies around it (SmallFrame)
  if (framed) {
    pl.ns = fr.counts[0]; pl.nd = fr.counts[1]; pl.np = fr.counts[2];
    if (threadIdx.x < 3) v.counts[w * 3 + threadIdx.x] = fr.counts[threadIdx.x];
    if (threadIdx.x == 0) v.n_events[w] = 0;
  }
  // A game scene's records live in shared memory for the launch (bouncehouse.json: 15 bodies, 4.5 KB): the kernel is one
  // block whose every visit is a chain of dependent reads of a few record fields, and those were L2 round trips.
  __shared__ __align__(16) unsigned char s_bodies_raw[kSmallStageBodies * sizeof(PgBody)];
  PgBody *const s_bodies = reinterpret_cast<PgBody *>(s_bodies_raw);
  PgBody *const gS = S, *const gD = D, *const gP = P;
  const bool in_smem = pl.ns + pl.nd + pl.np <= kSmallStageBodies;
  if (in_smem) {
    constexpr int kWords = (int)(sizeof(PgBody) / 16);
    static_assert(sizeof(PgBody) % 16 == 0, "PgBody is copied in 16-byte pieces");
    const int n_all = pl.ns + pl.nd + pl.np;
    for (int k = threadIdx.x; k < n_all * kWords; k += kThreads) {
      const int b = k / kWords, q = k % kWords;
      const PgBody *src = b < pl.ns ? gS + b : (b < pl.ns + pl.nd ? gD + (b - pl.ns) : gP + (b - pl.ns - pl.nd));
      reinterpret_cast<uint4 *>(s_bodies + b)[q] = reinterpret_cast<const uint4 *>(src)[q];
    }
    S = s_bodies; D = s_bodies + pl.ns; P = s_bodies + pl.ns + pl.nd;
    __syncthreads();
    if (framed) {
      // the records the host changed since the device last saw them, straight out of the pinned block
      int at = 0;
      for (int r = 0; r < fr.n_ranges; r++) {
        PgBody *dst = small_body(S, D, P, fr.cls[r], fr.first[r]);
        for (int k = threadIdx.x; k < fr.count[r] * kWords; k += kThreads)
          reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(fr.in + at)[k];
        at += fr.count[r];
      }
      __syncthreads();
    }
  }
  pl.v1 = pl.np * pl.ns; pl.v2 = pl.np * pl.nd; pl.v3 = pl.nd * pl.ns; pl.v4 = pl.nd * (pl.nd - 1);
  const int V = pl.total();
  const float dt = pgm::clamp_dt(dt_in);
  EventSink sink;
  sink.buf = v.events + (size_t)w * v.max_events;
  sink.counter = v.n_events + w;
  sink.cap = v.max_events;
  sink.world = w;
  // the integrations whose point lies in (applied, to] happen for real
  int applied = -1;
  auto advance = [&](int to) {
    for (int b = threadIdx.x; b < pl.np + pl.nd; b += kThreads) {
      const int cls = b < pl.np ? PG_CLASS_PLAYER : PG_CLASS_DYNAMIC, i = b < pl.np ? b : b - pl.np;
      const int ip = pl.int_point(cls, i);
      if (ip > applied && ip <= to) integrate_body(*small_body(S, D, P, cls, i), dt);
    }
    applied = to;
    __syncthreads();
  };
  __shared__ pgm::BoxPre s_node_pre[kSmallStageBodies];
  const pgm::BoxPre *const node_pre = in_smem ? s_node_pre : nullptr;
  for (int step = 0; step < n_steps; step++) {
    sink.step = step;
    int cursor = 0;
    applied = -1;
    if (in_smem) {
      // every box through its node, once per body and step (each of its ~2 n visits used to rebuild it)
      for (int b = threadIdx.x; b < pl.ns + pl.nd + pl.np; b += kThreads)
        if (s_bodies[b].collider_type == PG_AABB) s_node_pre[b] = pgm::box_pre_node(s_bodies[b]);
      __syncthreads();
    }
    for (int k = threadIdx.x; k < V; k += kThreads) {
      float t = 0.0f;
      s_fire[k] = (unsigned char)small_test(pl, S, D, P, k, -1, dt, &t, nullptr, nullptr, node_pre);
      s_t[k] = t;
    }
    __syncthreads();
    // visits whose interval search is a long one: the whole block decides them, one after the other
    auto settle_hard = [&](int from) {
      int mine = 0;                                      // (most steps have none: one vote instead of a scan by every thread)
      for (int k = from + (int)threadIdx.x; k < V; k += kThreads) mine |= s_fire[k] == 2;
      if (!__syncthreads_or(mine)) return;
      for (int k = from; k < V; k++) {
        if (s_fire[k] != 2) continue;                    // (uniform: shared memory, read between barriers)
        float t = 0.0f;
        const int r = small_test(pl, S, D, P, k, applied, dt, &t, &s_bs, &s_tmp, node_pre);
        __syncthreads();
        if (threadIdx.x == 0) { s_fire[k] = (unsigned char)r; s_t[k] = t; }
        __syncthreads();
      }
    };
    settle_hard(0);
    for (;;) {
      if (threadIdx.x == 0) s_min = INT_MAX;
      __syncthreads();
      int found = INT_MAX;
      for (int k = cursor + (int)threadIdx.x; k < V; k += kThreads) if (s_fire[k]) { found = k; break; }
      if (found != INT_MAX) atomicMin(&s_min, found);
      __syncthreads();
      const int kmin = s_min;
      __syncthreads();
      if (kmin == INT_MAX) break;
      advance(kmin);                            // the real state as visit kmin sees it
      int rc, ri, cc, ci, pass;
      pl.decode(kmin, rc, ri, cc, ci, pass);
      if (threadIdx.x == 0)
        commit_visit(*small_body(S, D, P, rc, ri), rc, ri, *small_body(S, D, P, cc, ci), cc, ci, s_t[kmin], sink, pass);
      __syncthreads();
      advance(kmin + 1);                        // a row that ended with this very visit integrates now
      cursor = kmin + 1;
      // the two bodies changed: their later visits are evaluated again (everything else stands)
      for (int k = cursor + (int)threadIdx.x; k < V; k += kThreads) {
        int rc2, ri2, cc2, ci2, p2;
        pl.decode(k, rc2, ri2, cc2, ci2, p2);
        const bool hit_row = (rc2 == rc && ri2 == ri) || (rc2 == cc && ri2 == ci);
        const bool hit_col = (cc2 == rc && ci2 == ri) || (cc2 == cc && ci2 == ci);
        if (!hit_row && !hit_col) continue;
        float t = 0.0f;
        s_fire[k] = (unsigned char)small_test(pl, S, D, P, k, applied, dt, &t, nullptr, nullptr, node_pre);
        s_t[k] = t;
      }
      __syncthreads();
      settle_hard(cursor);
    }
    advance(0x7ffffffe);                        // the integrations that are left
    // scene_node_update of every node (scene.c:1051-1079), as scene_update does after the step
    if (refresh_nodes) {
      for (int i = threadIdx.x; i < pl.ns; i += kThreads) if (S[i].has_node) pgm::node_refresh(S[i]);
      for (int i = threadIdx.x; i < pl.nd; i += kThreads) if (D[i].has_node) pgm::node_refresh(D[i]);
      for (int i = threadIdx.x; i < pl.np; i += kThreads) if (P[i].has_node) pgm::node_refresh(P[i]);
      __syncthreads();
    }
  }
  if (in_smem) {
    constexpr int kWords = (int)(sizeof(PgBody) / 16);
    const int n_all = pl.ns + pl.nd + pl.np;
    __syncthreads();
    for (int k = threadIdx.x; k < n_all * kWords; k += kThreads) {
      const int b = k / kWords, q = k % kWords;
      PgBody *dst = b < pl.ns ? gS + b : (b < pl.ns + pl.nd ? gD + (b - pl.ns) : gP + (b - pl.ns - pl.nd));
      const uint4 word = reinterpret_cast<const uint4 *>(s_bodies + b)[q];
      reinterpret_cast<uint4 *>(dst)[q] = word;
      if (framed) {
        // ... and into the pinned block the host reads after its one synchronisation
        PgBody *h = b < pl.ns ? fr.out + b : (b < pl.ns + pl.nd ? fr.out + v.cap[0] + (b - pl.ns) : fr.out + v.cap[0] + v.cap[1] + (b - pl.ns - pl.nd));
        reinterpret_cast<uint4 *>(h)[q] = word;
      }
    }
    if (framed) {
      const int n_ev = v.n_events[w];                    // (every emit lies before the barrier above)
      if (threadIdx.x == 0) *fr.cnt_out = n_ev;
      constexpr int kEvWords = (int)(sizeof(PgEvent) / 16);
      const uint4 *src = reinterpret_cast<const uint4 *>(v.events + (size_t)w * v.max_events);
      for (int k = threadIdx.x; k < min(n_ev, v.max_events) * kEvWords; k += kThreads) reinterpret_cast<uint4 *>(fr.ev_out)[k] = src[k];
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

// Can a frame of these class lengths run without copies (k_step_small with the records in shared memory)?
bool pg_small_frame_fits(const PgWorlds *w, const int32_t counts[3]) {
  static const bool no_small = getenv("PG_GENERAL_ROW_SWEEP") != nullptr, no_zero_copy = getenv("PG_FRAME_DMA") != nullptr;
  if (no_small || no_zero_copy || w->v.n_worlds != 1) return false;
  const long long ns = counts[0], nd = counts[1], np = counts[2];
  return ns + nd + np <= kSmallStageBodies && np * ns + np * nd + nd * ns + nd * (nd - 1) <= kSmallMaxVisits;
}

int pg_launch_general_step(PgWorlds *w, float dt, int n_steps, int refresh_nodes, const SmallFrame *frame) {
  SmallFrame fr;
  memset(&fr, 0, sizeof(fr));
  fr.world = -1;
  if (frame) fr = *frame;
  // game-sized worlds (every world of the handle has at most kSmallMaxVisits visits per step): speculate + commit
  static const bool no_small = getenv("PG_GENERAL_ROW_SWEEP") != nullptr;
  long long worst = 0;
  for (int wd = 0; wd < w->v.n_worlds; wd++) {
    const long long ns = w->h_counts[wd * 3 + 0], nd = w->h_counts[wd * 3 + 1], np = w->h_counts[wd * 3 + 2];
    worst = std::max(worst, np * ns + np * nd + nd * ns + nd * (nd - 1));
  }
  if (!no_small && worst <= kSmallMaxVisits) {
    k_step_small<<<w->v.n_worlds, kThreads, 0, w->stream>>>(w->v, dt, n_steps, refresh_nodes, fr);
    PG_CUDA(cudaGetLastError());
    w->launches++;
    return PG_OK;
  }
  if (frame) { pg_set_error("pg_launch_general_step: a copy-free frame needs the game-sized kernel"); return PG_ERR_ARG; }
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
