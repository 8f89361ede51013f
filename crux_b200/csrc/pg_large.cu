// pg_large.cu -- physics_step for ONE large world of dynamic spheres + static planes (configs C2, C4).
//
// The reference visits all N(N-1) ordered pairs in array order, mutating bodies in place
// (world.c:473-554).  Here the same RESULT is produced without the quadratic sweep and without
// giving up the solve order:
//
//   1. pass 3 (spheres x statics) touches only the sphere -> one thread per sphere: a streaming probe
//      kernel with the cheap gates, then the exact visits of the few bodies that pass one
//                                                                        [k_pass3, k_pass3_exact]
//   2. BROADPHASE OVER PHASE ENTRIES.  During pass 4 a body is seen in two phases: before its own row has
//      integrated it (phase A: by the visits of rows <= its own) and after (phase B: by the rows that follow;
//      a body at rest has no phase B).  A visit (row, col) pairs phase A of `row` with phase A of `col` when
//      col > row and with phase B of `col` when col < row -- never B with B.  In either phase the body moves
//      on a straight line for the length of the visit, x(t) = c + v t, t in [0, dt], so it stays inside the
//      ball around the MIDPOINT of that segment
//            mid = c + v dt/2,   rho = r + |v| dt/2   (+ rounding slack)
//      and a visit can only fire -- the closest approach of the two segments must come within r_a + r_b plus
//      what the interval search can still bridge, pgs::pair_gate_rejects -- if the two balls touch.  The
//      balls are what goes into the uniform grid: two entries per body, binned by midpoint.  Compared with one
//      ball per body that covers both phases and the predicate's own sum-of-speeds reach
//      (r + 2 (|v| + g dt) dt: the previous scheme) the search radius of a fast body drops by 4x, and a gas
//      that falls in parallel at 50 m/s lists a handful of candidates per body instead of hundreds.
//      Candidates are VISITS (row, col): entry pair (A_i, A_j), i < j -> visit (i, j) [and (j, i) too when i
//      rests]; (B_i, A_j), i < j -> visit (j, i).  Each carries the verdict of the closest-approach gate on
//      the untouched states, so the first iteration only evaluates the few that pass it.
//                                           [k_cell_count, scan, k_scatter (counting sort), k_pairs]
//   3. ORDER-PRESERVING FIXED POINT.  Each body keeps a timeline: the contacts it has suffered so
//      far, keyed by (row, column) = the position of the visit in the reference's sweep, with the
//      state they left behind.  A candidate visit is evaluated on the states the timelines imply AT THAT KEY
//      (including whether the body's own row, hence its integration, lies before the key).  The hits found
//      form the next timelines; iterate until nothing changes.  The earliest event of the true sequential
//      sweep is found in the first iteration and never changes again; by induction iteration k fixes the
//      k-th link of every dependency chain, so the fixed point IS the sequential result -- bit for bit,
//      since every visit runs the exact arithmetic of pg_math.cuh.  An iteration only looks at what can
//      have changed: k_begin copies the entries of bodies whose timeline did not change (with
//      partners that did not either) and lists the visits that need a look; k_eval evaluates them,
//      one thread per visit, and hands the rare long interval searches to k_eval_hard (a block per
//      visit); k_check compares the timelines of the bodies involved.
//                                                          [k_begin, k_eval, k_eval_hard, k_check]
//   4. a contact gives a body a state its phase entry does not describe (it leaves the contact point in a new
//      direction).  Every state that appears in a timeline gets a ball of its own -- same midpoint construction,
//      so it stays small -- which is (a) linked into a second, incremental cell structure over the same grid and
//      (b) queried against the cell-sorted untouched entries and against the state balls linked so far; entry
//      pairs that touch name the visits to add (deduplicated: visits the untouched entries already listed are
//      recognised by repeating their ball test, the others go through a hash set).  Key ranges are ignored, so
//      the list is a superset of every visit that could fire under ANY combination of the states seen so far,
//      and at the fixed point it contains every visit the final timelines could fire.
//                                                         [k_check, k_states_insert, k_states_query]
//   5. final state = timeline tail, integrated if the body's own row came after it; the fixed
//      point's per-body bookkeeping is put back for the bodies that used it.
//                                                                    [k_finalize, k_reset_touched]
//
// Steps 2 and 5 are the streaming, HBM-bound kernels of SURVEY.md section 8(d).
#include "pg_common.cuh"
#include "pg_math.cuh"
#include "pg_sphere_exact.cuh"

#include <algorithm>
#include <vector>

namespace {

using pgm::V3;
using pgm::v3;
using pgs::Sphere;

constexpr int kMaxPlanesL = 16;
constexpr int kTimeline = 24;           // contacts per body per step the timelines can hold (a sphere buried in a pile: 12 neighbours, each pair visited twice)
constexpr int kMaxIter = 256;       // a dependency chain of the sweep is resolved one link per iteration; later iterations only touch what changed
constexpr unsigned kFullMask = 0xffffffffu;

enum { K_PASS3 = 0, K_BOUNDS, K_CELLS, K_SCAN, K_SCATTER, K_PAIRS, K_EVAL, K_CHECK, K_FINAL, K_COUNT };
const char *kKernelNames[K_COUNT] = {"k_pass3", "k_states", "k_cell_count", "k_scan", "k_scatter", "k_pairs", "k_eval", "k_check", "k_finalize"};

struct GridParams {
  float org[3];
  float inv_h;
  int dim[3];
  int n_cells;
  int ok;
  float e_cap;            // ball radius the cell edge is sized for; entries beyond it search wider (k_pairs)
  float e_cap2;           // radius bound of the balls linked into the second cell structure (state balls, and the untouched entries
                          // beyond e_cap): a query searches that far; the few beyond it sit on a list every query walks
};

// Running statistics that keep the grid sized for the BULK of the bodies whatever a few stragglers do
// (a sphere that tunnelled out of the level falls for ever, faster and faster): centre and accepted
// half-widths carried from step to step; only bodies inside them contribute to the next ones.
struct Robust {
  float centre[3], half[3];
  int valid;
};

struct Flags {
  int changed;            // fixed point: some timeline differs from the previous iteration
  int overflow_balls;     // state-ball buffer full
  int overflow_vset;      // visit hash set full
  int overflow_huge;      // more balls beyond e_cap2 than their list holds
  unsigned tail_max;      // largest radius on the tail level of the second structure (ordered-uint encoding; 0: level empty)
  unsigned n_balls;       // state balls linked so far this step
  unsigned balls_begin;   // ... of which [balls_begin, n_balls) appeared in the current iteration
  unsigned n_huge;        // balls beyond e_cap2 (on the list every query walks)
  unsigned n_vset;        // visits in the hash set
  int overflow_pairs;
  int overflow_timeline;
  int unsafe;             // non-finite state or an envelope the grid cannot hold
  unsigned long long n_pairs;
  unsigned long long n_hits;
  unsigned long long n_carried;   // hits d_clear_carry copied this iteration; folded into n_hits by k_check
  unsigned n_touched;
  unsigned n_active;      // candidate pairs the selection (k_begin) handed to k_eval this iteration
  unsigned n_hard;        // of those, pairs k_eval passed on to k_eval_hard (long interval searches)
  unsigned n_pass3;       // bodies k_pass3 passed on to k_pass3_exact
  int stop;               // fixed point: the iterations still queued behind this one return at once (converged, or the host must act)
  int converged;
  unsigned iters_done;    // iterations that really ran this step
};

// One phase entry of the cell-ordered copy.  A whole 32-byte sector per entry: the counting sort's scattered
// store then never leaves a partial sector behind (16 B + 4 B stores into shared sectors made the L2
// fetch the rest from DRAM: 2.6x the algorithmic traffic), and a reader gets everything a candidate test
// needs -- the exact centre and velocity the phase's visits see (the gate's inputs), the ball radius, and the
// entry id (2 * body + phase) -- without a second look-up.
struct __align__(32) SortedBody { float4 crho; float vx, vy, vz; unsigned idx; };
// One state ball: midpoint and radius, entry id (2 body + phase) | kRestBit, next ball of its cell (0xffffffff: none)
// -- a whole sector, so that walking a cell's list costs one load per node.
struct __align__(32) BallRec { float4 c; unsigned ent, next, pad0, pad1; };

// Phase entries.  entry id e = 2 * body + phase (0 = A: before the body's own row integrates it, 1 = B: after).
__device__ __forceinline__ float3 entry_mid(float cx, float cy, float cz, float vx, float vy, float vz, float dt) {
  const float h = 0.5f * dt;
  return make_float3(__fmaf_rn(vx, h, cx), __fmaf_rn(vy, h, cy), __fmaf_rn(vz, h, cz));
}
// Radius of the ball around the midpoint that holds the body for the whole visit, plus the share of the gate's
// tolerance (pgs::pair_gate_rejects: delta = 1.1e-6 vsum + 4e-6 (cmax + vsum dt) + 1e-6) this body answers for,
// with a 5e-4 relative cushion over every rounding in here and in the float ball test.
__device__ __forceinline__ float entry_rho(float r, float cx, float cy, float cz, float vx, float vy, float vz, float dt) {
  const float sp = sqrtf(__fmaf_rn(vx, vx, __fmaf_rn(vy, vy, vz * vz)));
  const float cm = fmaxf(fmaxf(fabsf(cx), fabsf(cy)), fabsf(cz));
  return __fmaf_rn(0.5f * sp, dt, r) * 1.0005f + (2.0e-6f * sp + 8.0e-6f * (cm + sp * dt) + 4.0e-6f);
}
// State of an untouched body in phase `ph` from its post-pass-3 record: the record itself, or what its own
// row's integration makes of it (world.c:549-553; the very function the real integration uses).
__device__ __forceinline__ void entry_state(const float4 p, const float4 q, int ph, float dt, float *c, float *vel) {
  V3 pp = v3(p.x, p.y, p.z), vv = v3(q.x, q.y, q.z);
  if (ph) pgm::integrate(pp, vv, dt);
  c[0] = pp.x; c[1] = pp.y; c[2] = pp.z; vel[0] = vv.x; vel[1] = vv.y; vel[2] = vv.z;
}

struct LargeView {
  int n, ns;
  int stage_limit;                   // entries the per-block shared-memory stages may hold (tests shrink it to reach the overflow paths)
  int pass3_split_min;               // worlds of at least this many bodies run pass 3 as probe + dense exact kernel
  float4 *pos_rest, *vel_rad;        // state (in/out), body order = reference array order
  float4 *snap_pos, *snap_vel;       // the state as it was when the current step began (a step that fails puts it back)
  int *snap_n_events;
  float4 *planes;                    // the static PLANES, array order
  int *plane_sidx;                   // their index in the static array
  float restitution;
  // level geometry: static AABBs (world boxes precomputed on the host) in a uniform grid
  int n_static_total, n_boxes;
  float4 *box_c, *box_e;             // [n_static_total] world centre / extents (boxes only)
  unsigned char *static_type;        // [n_static_total] PG_AABB / PG_PLANE
  float sg_org[3], sg_inv_h, sg_reach;   // cell(p) = floor((p - org) * inv_h); boxes inserted inflated by sg_reach
  int sg_dim[3];
  unsigned *sg_start;                // [cells + 1]
  int *sg_items;                     // static indices, ascending inside a cell
  float *env;                        // [2n] ball radius per phase entry (entry e = 2 body + phase); < 0: no such entry (body at rest has no phase B)
  float r_uniform;                   // >= 0: every sphere has this world radius (no per-body look-up in k_pairs)
  float ecap_factor;                 // grid sizing: cells hold balls up to ecap_factor x the mean radius
  float ecap2_factor;                // k_requery's partner bound e_cap2 = ecap2_factor x e_cap
  // reductions
  unsigned *bbox;                    // 6 ordered-uint encodings: min xyz, max xyz ; [6] = max env
  float *rstat;                      // [8] sums over the bodies inside the previous robust box: x y z |dx| |dy| |dz| env count
  Robust *robust;
  // state balls (fixed point, step 4): append-only per step
  struct BallRec *balls;             // [balls_cap] one 32-byte record per ball
  int balls_cap;
  unsigned *head2, *head3;           // [cells_cap] first ball of every cell of the second structure (0xffffffff: empty): balls up to
                                     // e_cap / up to e_cap2
  int *huge_list;                    // [huge_cap] balls beyond grid->e_cap2
  int huge_cap;
  unsigned long long *vset;          // visit hash set: keys (row << 32 | col), 0xff..ff = empty slot
  unsigned long long vset_mask;
  GridParams *grid;
  // grid
  unsigned *cell_of;                 // [2n] (0xffffffff: entry not in the grid)
  unsigned *cell_cnt;                // [cells_cap + 1]  counts -> exclusive starts
  unsigned *block_sums;              // scan scratch
  int cells_cap;
  SortedBody *sorted;                // [2n] entries in cell order: centre, radius, velocity, entry id -- one 32-byte sector each
  // candidates
  int2 *pairs;                       // candidate VISITS: .x = row (~row: appended by k_requery, never evaluated yet), .y = column
                                     // (sign bit set: the closest-approach gate rejected the visit on the untouched states)
  long long pairs_cap;
  unsigned *active;                  // [2 pairs_cap] indices into pairs[]: the ones this iteration has to look at (k_begin), and from
                                     // pairs_cap on the ones k_eval deferred to k_eval_hard
  // timelines, double buffered
  unsigned char *tl_cnt[2];
  unsigned long long *tl_chg[2];      // smallest key of an entry that DISAPPEARED since the previous iteration (~0 = none);
                                      // new / changed entries carry a flag in tl_pos[..].w
  int *seen;                          // body already on the touched list (this step)
  unsigned *touched_bits;             // the same as a bitmap (n/32 words): the cheap test k_eval makes for every pair
  unsigned *dirty_bits[2];            // per timeline buffer: the body's timeline differs from the previous iteration's (k_check)
  unsigned *pushed_bits;              // the body received an entry from k_eval in the current iteration, or had one that k_eval
                                      // was to decide about again: k_check must compare its timelines
  int *touched;                       // bodies that ever received a timeline entry this step
  unsigned *cell_rank;                // [2n] rank of an entry inside its cell (returned by the counting atomic)
  unsigned long long *tl_key[2];
  float4 *tl_pos[2], *tl_vel[2];
  Flags *flags;
  // events
  PgEvent *events;
  int *n_events;
  int max_events;
  float leaf_lo, leaf_hi;
};

// order-preserving float <-> uint for atomicMin/Max
__device__ __forceinline__ unsigned f2o(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float o2f(unsigned o) {
  unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

// Contact events of the streaming kernels (k_pass3, k_finalize).  Tens of thousands per step used to take
// one atomic each on the single event counter, which serialised them for about as long as the kernel needs
// to stream the whole state; a block now stages its events in shared memory and appends them with one
// atomic when it is through (kEventInit / kEventFlush: called by ALL threads of the block).  The order in
// the buffer is irrelevant: the host restores solve order at read-back.
constexpr int kEventStage = 192;
enum { kEventAdd = 0, kEventFlush = 1, kEventInit = 2 };
// (the view's fields are passed by value: a reference to the kernel parameter would make every thread copy
// the whole 400-byte struct to its local stack)
__device__ __noinline__ void event_stage_impl(PgEvent *events, int *n_events, int max_events, int cap, int mode, int step, int pass, int ac, int ai, int bc, int bi) {
  __shared__ PgEvent s_ev[kEventStage];
  __shared__ unsigned s_cnt;
  __shared__ int s_base;
  if (mode == kEventInit) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    return;
  }
  if (mode == kEventAdd) {
    if (max_events <= 0) return;
    PgEvent e;
    e.world = 0; e.step = step; e.event_type = 0;
    e.a_class = ac; e.a_index = ai; e.b_class = bc; e.b_index = bi;
    e.pad_ = pass << 8;
    const unsigned st = atomicAdd(&s_cnt, 1u);
    if (st < (unsigned)cap) { s_ev[st] = e; return; }
    const int slot = atomicAdd(n_events, 1);
    if (slot < max_events) events[slot] = e;
    return;
  }
  __syncthreads();
  const int staged = (int)min(s_cnt, (unsigned)cap);
  if (threadIdx.x == 0 && staged) s_base = atomicAdd(n_events, staged);
  __syncthreads();
  for (int e = threadIdx.x; e < staged; e += blockDim.x)
    if (s_base + e < max_events) events[s_base + e] = s_ev[e];
  __syncthreads();
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
}
__device__ __forceinline__ void event_stage(const LargeView &v, int mode, int step, int pass, int ac, int ai, int bc, int bi) {
  event_stage_impl(v.events, v.n_events, v.max_events, min(v.stage_limit, kEventStage), mode, step, pass, ac, ai, bc, bi);
}
__device__ __forceinline__ void emit_event(const LargeView &v, int step, int pass, int ac, int ai, int bc, int bi) {
  event_stage(v, kEventAdd, step, pass, ac, ai, bc, bi);
}

// ---------------------------------------------------------------- 1. pass 3 + envelope + bounds
// Bodies that receive their first timeline entry join the touched list.  Tens of thousands do in the
// first iteration of a step, and one atomic each on the list's counter serialises them; they are staged
// per block instead and appended with one atomic when the block is through (kTouchInit / kTouchFlush:
// called by ALL threads of the block at the start / end of every phase that pushes timeline entries).
constexpr int kTouchStage = 512;
enum { kTouchAdd = 0, kTouchFlush = 1, kTouchInit = 2 };
__device__ __noinline__ void touched_stage_impl(int *touched, unsigned *n_touched, int cap, int b, int mode) {
  __shared__ int s_list[kTouchStage];
  __shared__ unsigned s_cnt, s_base;
  if (mode == kTouchInit) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    return;
  }
  if (mode == kTouchAdd) {
    const unsigned slot = atomicAdd(&s_cnt, 1u);
    if (slot < (unsigned)cap) s_list[slot] = b;
    else touched[atomicAdd(n_touched, 1u)] = b;
    return;
  }
  __syncthreads();
  const unsigned staged = min(s_cnt, (unsigned)cap);
  if (threadIdx.x == 0 && staged) s_base = atomicAdd(n_touched, staged);
  __syncthreads();
  for (unsigned e = threadIdx.x; e < staged; e += blockDim.x) touched[s_base + e] = s_list[e];
  __syncthreads();
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
}

__device__ __forceinline__ void touched_stage(const LargeView &v, int b, int mode) { touched_stage_impl(v.touched, &v.flags->n_touched, min(v.stage_limit, kTouchStage), b, mode); }
// Pass 3 runs as two kernels.  Probe = true (k_pass3, one thread per body, few registers, streaming): the
// cheap gates only; a body none of whose statics passes its gate is left as it is and gets its envelope,
// bookkeeping and bounds here, the others (a few per cent; 15 % in a level full of slabs) go on a list.
// Probe = false (k_pass3_exact, dense over that list): the whole visit sequence of the body -- gates again,
// exact predicate, narrowphase, resolver, statics in array order -- then the same envelope / bounds code.
// One kernel doing both ran every warp through the 124-register exact path with a few lanes alive.
// All = true (with Probe = false): the exact kernel over every body, in one launch -- for worlds small enough
// to be bound by their launches.
template <bool Probe, bool All = false>
__device__ __forceinline__ void pass3_impl(const LargeView &v, float dt, int step, int do_planes, int full_reset) {
  __shared__ float4 s_pl[kMaxPlanesL];
  __shared__ float s_red[8][15];                   // per-warp partials of the bounds and the running statistics
  if (threadIdx.x < v.ns) s_pl[threadIdx.x] = v.planes[threadIdx.x];
  event_stage(v, kEventInit, 0, 0, 0, 0, 0, 0);     // (includes the block barrier)
  touched_stage_impl(v.touched, &v.flags->n_pass3, kTouchStage, 0, kTouchInit);
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;
  const bool gate_ok = dt > 1.0e-9f;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f}, emax = 0.0f;
  float rs[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  const Robust rb = *v.robust;
  bool bad = false;
  const int n_items = (Probe || All) ? v.n : (int)v.flags->n_pass3;
  for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
    const int i = (Probe || All) ? item : v.touched[item];   // (the touched list's storage is free until k_eval)
    float4 p = v.pos_rest[i], q = v.vel_rad[i];
    bool needs = false;                                    // Probe: some static passed its gate
    if (do_planes) {
      bool touched = false;
      const V3 p0 = v3(p.x, p.y, p.z);        // scene-node translation: the position at step start
      auto do_plane = [&](int k) {
        const float4 pl = s_pl[k];
        const float rho = pgs::reach(q.w, q.x, q.y, q.z, dt);
        const float s0 = __fmaf_rn(p.x, pl.x, __fmaf_rn(p.y, pl.y, p.z * pl.z)) - pl.w;
        if (gate_ok && fabsf(s0) > rho * 1.0001f + 1.0e-5f * (fabsf(s0) + fabsf(pl.w) + 1.0f)) return;
        if (Probe) { needs = true; return; }
        Sphere A;
        A.px = p.x; A.py = p.y; A.pz = p.z; A.vx = q.x; A.vy = q.y; A.vz = q.z; A.r = q.w; A.rest = 0;
        if (pgs::sphere_plane_exact(A, pl, v.restitution, dt, leaf)) {
          p.x = A.px; p.y = A.py; p.z = A.pz; q.x = A.vx; q.y = A.vy; q.z = A.vz;
          if (A.rest) p.w = 1.0f;
          touched = true;
          emit_event(v, step, 3, PG_CLASS_DYNAMIC, i, PG_CLASS_STATIC, v.plane_sidx[k]);
        }
      };
      auto do_box = [&](int j) {
        pgm::BoxS bx;
        const float4 bc = v.box_c[j], be = v.box_e[j];
        bx.c = v3(bc.x, bc.y, bc.z); bx.e = v3(be.x, be.y, be.z);
        const V3 vel = v3(q.x, q.y, q.z);
        const float nv = pgm::norm(vel);
        // cheap gate on the root test: distance at t=0 must not exceed r + |v| dt
        const float rho = __fmaf_rn(nv, dt, q.w) * 1.0001f + 1.0e-6f;
        const float d2 = pgm::bb_d2(v3(0.0f + p0.x, 0.0f + p0.y, 0.0f + p0.z), bx);
        if (gate_ok && d2 > rho * rho * 1.0001f) return;
        if (Probe) { needs = true; return; }
        if (!pgm::bb_interval(p0, vel, nv, q.w, bx, 0.0f, dt, leaf)) return;
        const pgm::Contact c = pgm::bb_narrow(p0, vel, q.w, bx, dt);
        if (!(c.hit && c.t >= 0)) return;
        V3 pos = v3(p.x, p.y, p.z), vv = vel;
        pgm::bb_resolve(pos, vv, p0, q.w, v.restitution, bx, c.t);
        p.x = pos.x; p.y = pos.y; p.z = pos.z; q.x = vv.x; q.y = vv.y; q.z = vv.z;
        touched = true;
        emit_event(v, step, 3, PG_CLASS_STATIC, j, PG_CLASS_DYNAMIC, i);      // the box is body_A (lower collider type)
      };
      if (v.n_boxes == 0) {
        for (int k = 0; k < v.ns && !needs; k++) do_plane(k);
      } else {
        // planes and the boxes registered in this sphere's level-grid cell, merged in static-array order
        const float rho0 = pgs::reach(q.w, q.x, q.y, q.z, dt);
        int cx = (int)floorf((p0.x - v.sg_org[0]) * v.sg_inv_h), cy = (int)floorf((p0.y - v.sg_org[1]) * v.sg_inv_h),
            cz = (int)floorf((p0.z - v.sg_org[2]) * v.sg_inv_h);
        const bool inside = cx >= 0 && cy >= 0 && cz >= 0 && cx < v.sg_dim[0] && cy < v.sg_dim[1] && cz < v.sg_dim[2];
        if (rho0 * 1.01f + 1.0e-3f <= v.sg_reach) {
          unsigned ib = 0, ie = 0;
          if (inside) { const int cell = (cz * v.sg_dim[1] + cy) * v.sg_dim[0] + cx; ib = v.sg_start[cell]; ie = v.sg_start[cell + 1]; }
          int ip = 0;
          while ((ip < v.ns || ib < ie) && !needs) {
            const int jp = ip < v.ns ? v.plane_sidx[ip] : 0x7fffffff;
            const int jb = ib < ie ? v.sg_items[ib] : 0x7fffffff;
            if (jp < jb) { do_plane(ip); ip++; } else { do_box(jb); ib++; }
          }
        } else {
          // faster than the level grid was built for: walk the whole static array
          int ip = 0;
          for (int j = 0; j < v.n_static_total && !needs; j++) {
            if (v.static_type[j] == PG_PLANE) { do_plane(ip); ip++; } else do_box(j);
          }
        }
      }
      if (touched) { v.pos_rest[i] = p; v.vel_rad[i] = q; }
    }
    if (Probe && needs) { touched_stage_impl(v.touched, &v.flags->n_pass3, kTouchStage, i, kTouchAdd); continue; }
    // the body's two phase entries (midpoint balls of the untouched state, see the header comment)
    float cA[3], vA[3];
    entry_state(p, q, 0, dt, cA, vA);
    const float rhoA = entry_rho(q.w, cA[0], cA[1], cA[2], vA[0], vA[1], vA[2], dt);
    const float3 mA = entry_mid(cA[0], cA[1], cA[2], vA[0], vA[1], vA[2], dt);
    float rhoB = -1.0f;
    float3 mB = mA;
    if (p.w == 0.0f) {
      float cB[3], vB[3];
      entry_state(p, q, 1, dt, cB, vB);
      rhoB = entry_rho(q.w, cB[0], cB[1], cB[2], vB[0], vB[1], vB[2], dt);
      mB = entry_mid(cB[0], cB[1], cB[2], vB[0], vB[1], vB[2], dt);
    }
    reinterpret_cast<float2 *>(v.env)[i] = make_float2(rhoA, rhoB);
    const float e = fmaxf(rhoA, rhoB);
    if (full_reset) {
      // per-step bookkeeping of the fixed point: only bodies that get a timeline entry ever change it, and
      // k_reset_touched puts those back at the end of the step (23 B per body and step not written here)
      v.tl_cnt[0][i] = 0;
      v.tl_cnt[1][i] = 0;
      v.tl_chg[0][i] = 0xffffffffffffffffull;
      v.tl_chg[1][i] = 0xffffffffffffffffull;
      v.seen[i] = 0;
    }
    if (!(fabsf(mA.x) < 1.0e15f && fabsf(mA.y) < 1.0e15f && fabsf(mA.z) < 1.0e15f && fabsf(mB.x) < 1.0e15f && fabsf(mB.y) < 1.0e15f &&
          fabsf(mB.z) < 1.0e15f && e < 1.0e15f)) { bad = true; continue; }
    lo[0] = fminf(lo[0], fminf(mA.x, mB.x)); lo[1] = fminf(lo[1], fminf(mA.y, mB.y)); lo[2] = fminf(lo[2], fminf(mA.z, mB.z));
    hi[0] = fmaxf(hi[0], fmaxf(mA.x, mB.x)); hi[1] = fmaxf(hi[1], fmaxf(mA.y, mB.y)); hi[2] = fmaxf(hi[2], fmaxf(mA.z, mB.z));
    emax = fmaxf(emax, e);
    const float dx = fabsf(mA.x - rb.centre[0]), dy = fabsf(mA.y - rb.centre[1]), dz = fabsf(mA.z - rb.centre[2]);
    if (dx <= rb.half[0] && dy <= rb.half[1] && dz <= rb.half[2]) {
      rs[0] += mA.x; rs[1] += mA.y; rs[2] += mA.z; rs[3] += dx; rs[4] += dy; rs[5] += dz; rs[6] += rhoA; rs[7] += 1.0f;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    for (int k = 0; k < 3; k++) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(kFullMask, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFullMask, hi[k], o));
    }
    emax = fmaxf(emax, __shfl_xor_sync(kFullMask, emax, o));
    for (int k = 0; k < 8; k++) rs[k] += __shfl_xor_sync(kFullMask, rs[k], o);
  }
  // one set of global atomics per BLOCK (the 15 addresses are shared by every warp of the grid)
  const int wib = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && wib < 8) {
    for (int k = 0; k < 3; k++) { s_red[wib][k] = lo[k]; s_red[wib][3 + k] = hi[k]; }
    s_red[wib][6] = emax;
    for (int k = 0; k < 8; k++) s_red[wib][7 + k] = rs[k];
  }
  __syncthreads();
  if (threadIdx.x < 15) {
    const int k = threadIdx.x, nw = min((int)(blockDim.x >> 5), 8);
    float x = s_red[0][k];
    for (int q = 1; q < nw; q++) x = k < 3 ? fminf(x, s_red[q][k]) : (k < 7 ? fmaxf(x, s_red[q][k]) : x + s_red[q][k]);
    if (k < 3) atomicMin(v.bbox + k, f2o(x));
    else if (k < 7) atomicMax(v.bbox + k, f2o(x));
    else if (x != 0.0f) atomicAdd(v.rstat + (k - 7), x);
  }
  if (bad) atomicExch(&v.flags->unsafe, 1);
  event_stage(v, kEventFlush, 0, 0, 0, 0, 0, 0);
  touched_stage_impl(v.touched, &v.flags->n_pass3, kTouchStage, 0, kTouchFlush);
}
__global__ void k_pass3(LargeView v, float dt, int step, int do_planes, int full_reset) { pass3_impl<true>(v, dt, step, do_planes, full_reset); }
__global__ void k_pass3_exact(LargeView v, float dt, int step, int do_planes, int full_reset) { pass3_impl<false>(v, dt, step, do_planes, full_reset); }
__global__ void k_pass3_all(LargeView v, float dt, int step, int do_planes, int full_reset) { pass3_impl<false, true>(v, dt, step, do_planes, full_reset); }

__global__ void k_grid_params(LargeView v) {
  GridParams g;
  float lo[3], hi[3];
  for (int k = 0; k < 3; k++) { lo[k] = o2f(v.bbox[k]); hi[k] = o2f(v.bbox[3 + k]); }
  const float emax = o2f(v.bbox[6]);
  g.ok = 1;
  // The bulk: mean position, mean absolute deviation and mean envelope of the bodies that were inside
  // last step's accepted box.  The grid covers mean +- (6 deviations + room) -- a uniform box spans
  // +-2 deviations -- and its cells are sized for 4 mean envelopes; bodies outside are clamped into
  // the border cells, bodies with a larger envelope search more cells (k_pairs).
  float e_cap = emax;
  Robust rb = *v.robust;
  const float cnt = v.rstat[7];
  if (cnt >= 0.5f * (float)v.n && cnt >= 1.0f) {
    for (int k = 0; k < 3; k++) {
      const float mean = v.rstat[k] / cnt, dev = v.rstat[3 + k] / cnt;
      const float half = 6.0f * dev + 32.0f * (v.rstat[6] / cnt) + 4.0f;   // dev is about last step's centre: never too small
      if (rb.valid) { lo[k] = fmaxf(lo[k], mean - half); hi[k] = fminf(hi[k], mean + half); }
      rb.centre[k] = mean; rb.half[k] = half;
    }
    if (rb.valid) e_cap = fminf(emax, v.ecap_factor * (v.rstat[6] / cnt) + 1.0e-3f);
    rb.valid = 1;
  } else {
    // lost the bulk (or first step, or a frozen-state query): accept everything, start over
    for (int k = 0; k < 3; k++) { rb.centre[k] = 0.0f; rb.half[k] = 3.0e38f; }
    rb.valid = 0;
  }
  for (int k = 0; k < 3; k++) if (!(lo[k] <= hi[k])) { const float m = 0.5f * (o2f(v.bbox[k]) + o2f(v.bbox[3 + k])); lo[k] = hi[k] = m; }
  *v.robust = rb;
  // cell edge: two envelopes (one per body of a pair) with 25 % head room for envelopes that grow
  float h = 2.0f * e_cap * 1.25f + 1.0e-6f;
  if (!(h > 0.0f) || !(h < 1.0e15f) || !(lo[0] <= hi[0])) { g.ok = 0; v.flags->unsafe = 1; h = 1.0f; lo[0] = lo[1] = lo[2] = 0.0f; hi[0] = hi[1] = hi[2] = 0.0f; }
  for (;;) {
    double cells = 1.0;
    for (int k = 0; k < 3; k++) {
      double d = floor(((double)hi[k] - (double)lo[k]) / (double)h) + 2.0;
      g.dim[k] = (int)fmin(d, 2.0e9);
      cells *= d;
    }
    if (cells <= (double)v.cells_cap) { g.n_cells = g.dim[0] * g.dim[1] * g.dim[2]; break; }
    h *= 1.26f;                                // too many cells for the table: coarser grid (still conservative)
  }
  for (int k = 0; k < 3; k++) g.org[k] = lo[k] - 0.5f * h * 0.001f;
  g.inv_h = 1.0f / h;
  g.e_cap = 0.5f * h;                          // the largest ball the 27-cell search serves: 2 rho <= h (25 % above the sizing radius)
  g.e_cap2 = v.ecap2_factor * g.e_cap;
  *v.grid = g;
}

__device__ __forceinline__ int3 cell_coords(const GridParams &g, float x, float y, float z) {
  int3 c;
  c.x = min(max((int)((x - g.org[0]) * g.inv_h), 0), g.dim[0] - 1);
  c.y = min(max((int)((y - g.org[1]) * g.inv_h), 0), g.dim[1] - 1);
  c.z = min(max((int)((z - g.org[2]) * g.inv_h), 0), g.dim[2] - 1);
  return c;
}

constexpr unsigned kRestBit = 0x80000000u;      // SortedBody::idx, ball_e.x: the entry's body is at rest (it has no phase B)
constexpr unsigned kNoBall = 0xffffffffu;
constexpr unsigned long long kSlotEmpty = 0xffffffffffffffffull, kSlotTomb = 0xfffffffffffffffeull;

// Warp-aggregated increment of a hot counter: the lanes that are here together take one atomic between them.
__device__ __forceinline__ unsigned agg_inc(unsigned *ctr) {
  const unsigned m = __activemask();
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs((int)m) - 1;
  unsigned base = 0;
  if ((int)lane == leader) base = atomicAdd(ctr, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (unsigned)__popc(m & ((1u << lane) - 1u));
}
__device__ __forceinline__ unsigned long long agg_inc64(unsigned long long *ctr) {
  const unsigned m = __activemask();
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs((int)m) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}

// Link a ball into the second cell structure (same cells as the sorted copy; a list per cell, pushed at the
// head).  Insertions and queries never run in the same kernel.
__device__ __forceinline__ void ball_insert(const LargeView &v, const GridParams &g, float3 m, float rho, unsigned ent_flags) {
  if (!(fabsf(m.x) < 1.0e15f && fabsf(m.y) < 1.0e15f && fabsf(m.z) < 1.0e15f && rho < 1.0e15f)) { atomicExch(&v.flags->unsafe, 1); return; }
  const unsigned idx = agg_inc(&v.flags->n_balls);
  if (idx >= (unsigned)v.balls_cap) { atomicExch(&v.flags->overflow_balls, 1); return; }
  v.balls[idx].c = make_float4(m.x, m.y, m.z, rho);
  if (rho > g.e_cap2) {                          // beyond what a query's cell search covers: the list every query walks
    const unsigned hslot = atomicAdd(&v.flags->n_huge, 1u);
    if (hslot < (unsigned)v.huge_cap) v.huge_list[hslot] = (int)idx; else atomicExch(&v.flags->overflow_huge, 1);
    *reinterpret_cast<uint2 *>(&v.balls[idx].ent) = make_uint2(ent_flags, kNoBall);
    return;
  }
  // two levels over the same cells: the bulk (radius <= e_cap: a query looks into 27 cells), and the tail up to
  // e_cap2 (few balls, searched wider)
  const int3 c = cell_coords(g, m.x, m.y, m.z);
  const unsigned cell = (unsigned)((c.z * g.dim[1] + c.y) * g.dim[0] + c.x);
  unsigned *head = v.head2;
  if (rho > g.e_cap) { head = v.head3; atomicMax(&v.flags->tail_max, f2o(rho)); }
  *reinterpret_cast<uint2 *>(&v.balls[idx].ent) = make_uint2(ent_flags, atomicExch(&head[cell], idx));
}

// Visit hash set (open addressing, linear probing): returns the slot if `key` was inserted, -1 if it was there.
__device__ __forceinline__ long long vset_insert(const LargeView &v, unsigned long long key) {
  unsigned long long h = key * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  unsigned long long slot = h & v.vset_mask;
  for (int probe = 0; probe < 256; probe++) {
    const unsigned long long old = atomicCAS(&v.vset[slot], kSlotEmpty, key);
    if (old == kSlotEmpty) { agg_inc(&v.flags->n_vset); return (long long)slot; }
    if (old == key) return -1;
    slot = (slot + 1) & v.vset_mask;
  }
  atomicExch(&v.flags->overflow_vset, 1);
  return -1;
}

// ---------------------------------------------------------------- 2. counting sort of the phase entries into cells
__global__ void k_cell_count(LargeView v, float dt) {
  const GridParams g = *v.grid;
  const int n_ent = 2 * v.n;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_ent; e += gridDim.x * blockDim.x) {
    const float rho = v.env[e];
    if (rho < 0.0f) { v.cell_of[e] = 0xffffffffu; continue; }      // a body at rest has no phase B
    const int i = e >> 1;
    float c[3], vel[3];
    entry_state(v.pos_rest[i], v.vel_rad[i], e & 1, dt, c, vel);
    const float3 m = entry_mid(c[0], c[1], c[2], vel[0], vel[1], vel[2], dt);
    const int3 cc = cell_coords(g, m.x, m.y, m.z);
    const unsigned cell = (unsigned)((cc.z * g.dim[1] + cc.y) * g.dim[0] + cc.x);
    v.cell_of[e] = cell;
    v.cell_rank[e] = atomicAdd(v.cell_cnt + cell, 1u);
    // an entry whose ball exceeds what the 27-cell search of a state-ball query allows for is also linked into
    // the second structure, where queries search wider
    if (rho > g.e_cap) ball_insert(v, g, m, rho, (unsigned)e | (v.pos_rest[i].w != 0.0f ? kRestBit : 0u));
  }
}

// exclusive scan of cell_cnt[0..n_cells] in three launches (block sums, scan of sums, add)
constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;
__global__ void k_scan_local(LargeView v) {
  __shared__ unsigned s_warp[32];
  const int n = v.grid->n_cells + 1;
  const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
  unsigned x[kScanItems], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { x[k] = (base + k < n) ? v.cell_cnt[base + k] : 0u; sum += x[k]; }
  unsigned incl = sum;
  for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(kFullMask, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned w = s_warp[threadIdx.x], wi = w;
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(kFullMask, wi, o); if (threadIdx.x >= o) wi += t; }
    s_warp[threadIdx.x] = wi - w;
    if (threadIdx.x == 31) v.block_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  unsigned excl = incl - sum + s_warp[threadIdx.x >> 5];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) { if (base + k < n) v.cell_cnt[base + k] = excl; excl += x[k]; }
}
__global__ void k_scan_sums(LargeView v, int n_blocks) {
  // single block: serial over chunks of 1024 block sums (n_blocks <= a few thousand)
  __shared__ unsigned s_warp[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + threadIdx.x;
    unsigned x = i < n_blocks ? v.block_sums[i] : 0u, incl = x;
    for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(kFullMask, incl, o); if ((threadIdx.x & 31) >= o) incl += t; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned w = s_warp[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(kFullMask, wi, o); if (threadIdx.x >= o) wi += t; }
      s_warp[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const unsigned excl = incl - x + s_warp[threadIdx.x >> 5] + s_carry;
    if (i < n_blocks) v.block_sums[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + x;
    __syncthreads();
  }
}
__global__ void k_scan_add(LargeView v) {
  const int n = v.grid->n_cells + 1;
  const unsigned add = v.block_sums[blockIdx.x];
  const int base = blockIdx.x * kScanBlock * kScanItems + threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) if (base + k < n) v.cell_cnt[base + k] += add;
}

__global__ void k_scatter(LargeView v, float dt) {
  const int n_ent = 2 * v.n;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_ent; e += gridDim.x * blockDim.x) {
    const unsigned cell = v.cell_of[e];
    if (cell == 0xffffffffu) continue;
    const unsigned slot = v.cell_cnt[cell] + v.cell_rank[e];
    const int i = e >> 1;
    float c[3], vel[3];
    entry_state(v.pos_rest[i], v.vel_rad[i], e & 1, dt, c, vel);
    const float rho = v.env[e];
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(v.sorted + slot), "r"(__float_as_uint(c[0])), "r"(__float_as_uint(c[1])),
                 "r"(__float_as_uint(c[2])), "r"(__float_as_uint(rho)), "r"(__float_as_uint(vel[0])), "r"(__float_as_uint(vel[1])),
                 "r"(__float_as_uint(vel[2])), "r"((unsigned)e | (v.pos_rest[i].w != 0.0f ? 0x80000000u : 0u)) : "memory");
  }
}

// ---------------------------------------------------------------- candidate visits
// A pair of entries is listed by the one with the LARGER ball (ties: lower entry id), which therefore only has
// to find partners within 2 rho of itself: R = ceil(2 rho / h) cells each way -- 1 for everybody the
// grid was sized for (27 cells = 9 contiguous x-runs), more for the few fast entries, and the whole
// sorted array for one so fast that its search would cover most of the grid anyway.
constexpr int kMaxSearch = 12;
constexpr int kPairsBlock = 128;
constexpr int kPairsPrefetch = 2;      // partners whose two 16-byte halves are in flight together
constexpr int kPairsStage = 1024;       // visits a block collects in shared memory between two flushes
// Which visits does the touching entry pair (ea, eb) stand for?  (header comment, step 2)  Returns their number
// and (row, col) of each; `swap[k]`: the row body of visit k is eb's.
__device__ __forceinline__ int entry_pair_visits(unsigned ea, bool rest_a, unsigned eb, bool rest_b, int2 *out, bool *row_is_b) {
  const int ia = (int)(ea >> 1), ib = (int)(eb >> 1);
  const int pa = (int)(ea & 1u), pb = (int)(eb & 1u);
  if (ia == ib || (pa & pb)) return 0;
  if (!pa && !pb) {
    const bool a_lo = ia < ib;
    const int lo = a_lo ? ia : ib, hi = a_lo ? ib : ia;
    out[0] = make_int2(lo, hi); row_is_b[0] = !a_lo;
    if (a_lo ? rest_a : rest_b) { out[1] = make_int2(hi, lo); row_is_b[1] = a_lo; return 2; }   // lo never integrates: row hi sees it as it is
    return 1;
  }
  // one phase B, one phase A: the B body must be the lower index; the visit belongs to the A body's row
  if (pa) { if (!(ia < ib)) return 0; out[0] = make_int2(ib, ia); row_is_b[0] = true; return 1; }
  if (!(ib < ia)) return 0;
  out[0] = make_int2(ia, ib); row_is_b[0] = false;
  return 1;
}

// (out of line: the gate's registers stay out of the search loop, which is bound by the latency of its loads)
__device__ __noinline__ bool listing_gate_rejects(float4 cr, float4 wr, float4 cc, float4 wc, float rs, float dt) {
  const float na2 = __fmaf_rn(wr.x, wr.x, __fmaf_rn(wr.y, wr.y, wr.z * wr.z));
  const float nb2 = __fmaf_rn(wc.x, wc.x, __fmaf_rn(wc.y, wc.y, wc.z * wc.z));
  const float vsum = sqrtf(2.0f * (na2 + nb2)) * 1.0001f;          // >= |vA| + |vB|
  return pgs::pair_gate_rejects(v3(cr.x, cr.y, cr.z), v3(wr.x, wr.y, wr.z), v3(cc.x, cc.y, cc.z), v3(wc.x, wc.y, wc.z), rs, vsum, dt);
}

// One thread per entry (in cell order, so neighbouring threads read neighbouring cells); every thread
// walks its own cell runs without warp votes.  Two entries the grid was sized for (R = 1 both) lie in the
// same or in adjacent cells, so each of them searches only the HALF shell of cells that follow its own
// in cell order (the rest of its own x-run, the next y-row, the three rows of the next z-layer: 5 runs,
// 13.5 cells instead of 27) and lists whatever it finds there; an entry with a larger ball searches
// its whole neighbourhood and lists the partners whose ball is smaller (ties: lower id), and an
// R = 1 entry leaves such partners to them.  Visits are staged in shared memory and appended to the global
// list once per tile of kPairsBlock entries; a tile that finds more than kPairsStage (a dense pile)
// appends the surplus directly.
#ifndef PG_PAIRS_MIN_BLOCKS
#define PG_PAIRS_MIN_BLOCKS 8       // 64 registers: measured 1.31 -> 0.85 ms at 4 M bodies (96 registers, 5 blocks per SM before); 40 registers spill and lose 3x
#endif
__global__ void __launch_bounds__(kPairsBlock, PG_PAIRS_MIN_BLOCKS) k_pairs(LargeView v, float dt) {
  __shared__ int2 s_pairs[kPairsStage];
  __shared__ unsigned s_cnt;
  __shared__ unsigned long long s_base;
  const GridParams g = *v.grid;
  const int n_sorted = (int)v.cell_cnt[g.n_cells];
  const int n_tiles = (n_sorted + kPairsBlock - 1) / kPairsBlock;
  const bool gate_ok = dt > 1.0e-9f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int a = tile * kPairsBlock + threadIdx.x;
    if (a < n_sorted) {
      const float4 ca = v.sorted[a].crho;
      const float4 wa = *reinterpret_cast<const float4 *>(&v.sorted[a].vx);
      const unsigned ea = __float_as_uint(wa.w) & ~kRestBit;
      const bool rest_a = (__float_as_uint(wa.w) & kRestBit) != 0u;
      const float3 ma = entry_mid(ca.x, ca.y, ca.z, wa.x, wa.y, wa.z, dt);
      const int3 c = cell_coords(g, ma.x, ma.y, ma.z);
      const float need = 2.0f * ca.w * g.inv_h * 1.001f;
      const int R = need <= 1.0f ? 1 : (need <= (float)kMaxSearch ? (int)ceilf(need) : 0);
      const bool everything = !(need <= (float)kMaxSearch);
      const bool half = need <= 1.0f;
      auto run = [&](unsigned b0, unsigned b1) {
        // (the loads of up to four partners are issued before the first is looked at: the kernel is bound by
        // the latency of its dependent loads, not by instructions or bytes)
        float4 nb[kPairsPrefetch], nw[kPairsPrefetch];
        for (unsigned bb = b0; bb < b1; bb += kPairsPrefetch) {
#pragma unroll
        for (int u = 0; u < kPairsPrefetch; u++) {
          const SortedBody *sb = v.sorted + min(bb + u, b1 - 1);
          nb[u] = sb->crho; nw[u] = *reinterpret_cast<const float4 *>(&sb->vx);
        }
#pragma unroll
        for (int u = 0; u < kPairsPrefetch; u++) {
          const unsigned b = bb + u;
          if (b >= b1) break;
          const float4 cb = nb[u], wb = nw[u];
          const float3 mb = entry_mid(cb.x, cb.y, cb.z, wb.x, wb.y, wb.z, dt);
          const float dx = ma.x - mb.x, dyy = ma.y - mb.y, dzz = ma.z - mb.z;
          const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dyy, dyy, dzz * dzz));
          const float T = ca.w + cb.w;
          if (!(d2 <= T * T)) continue;
          const unsigned eb = __float_as_uint(wb.w) & ~kRestBit;
          const bool rest_b = (__float_as_uint(wb.w) & kRestBit) != 0u;
          if (half) {
            if (!(2.0f * cb.w * g.inv_h * 1.001f <= 1.0f)) continue;          // a wider-searching partner lists the pair itself
          } else {
            if (ca.w < cb.w || ea == eb || (ca.w == cb.w && !(ea < eb))) continue;
          }
          int2 vis[2]; bool row_b[2];
          const int nv = entry_pair_visits(ea, rest_a, eb, rest_b, vis, row_b);
          for (int q = 0; q < nv; q++) {
            // the closest-approach gate the exact predicate opens with, on the untouched states (row body first)
            bool rej = false;
            if (gate_ok) {
              const float rs = v.r_uniform >= 0.0f ? v.r_uniform + v.r_uniform : v.vel_rad[vis[q].x].w + v.vel_rad[vis[q].y].w;
              rej = row_b[q] ? listing_gate_rejects(cb, wb, ca, wa, rs, dt) : listing_gate_rejects(ca, wa, cb, wb, rs, dt);
            }
            const int2 pr = make_int2(vis[q].x, rej ? (int)((unsigned)vis[q].y | 0x80000000u) : vis[q].y);
            const unsigned slot = atomicAdd(&s_cnt, 1u);
            if (slot < (unsigned)min(v.stage_limit, kPairsStage)) { s_pairs[slot] = pr; continue; }
            const unsigned long long gs = atomicAdd(&v.flags->n_pairs, 1ull);
            if ((long long)gs < v.pairs_cap) v.pairs[gs] = pr;
            else atomicExch(&v.flags->overflow_pairs, 1);
          }
        }
        }
      };
      if (everything) {
        run(0u, (unsigned)n_sorted);
      } else if (half) {
        const int x0 = max(c.x - 1, 0), x1 = min(c.x + 1, g.dim[0] - 1);
        const unsigned row = (unsigned)((c.z * g.dim[1] + c.y) * g.dim[0]);
        // the bounds of all five runs first (independent loads), then the runs
        unsigned lo[5], hi[5];
        lo[0] = (unsigned)a + 1u; hi[0] = v.cell_cnt[row + x1 + 1];            // rest of the own cell, then cell x+1
        const bool up = c.y + 1 < g.dim[1];
        lo[1] = up ? v.cell_cnt[row + g.dim[0] + x0] : 0u; hi[1] = up ? v.cell_cnt[row + g.dim[0] + x1 + 1] : 0u;
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const int y = c.y - 1 + q;
          const bool ok = c.z + 1 < g.dim[2] && y >= 0 && y < g.dim[1];
          const unsigned r2 = (unsigned)(((c.z + 1) * g.dim[1] + y) * g.dim[0]);
          lo[2 + q] = ok ? v.cell_cnt[r2 + x0] : 0u; hi[2 + q] = ok ? v.cell_cnt[r2 + x1 + 1] : 0u;
        }
#pragma unroll
        for (int q = 0; q < 5; q++) run(lo[q], hi[q]);
      } else {
        const int x0 = max(c.x - R, 0), x1 = min(c.x + R, g.dim[0] - 1);
        for (int z = max(c.z - R, 0); z <= min(c.z + R, g.dim[2] - 1); z++)
          for (int y = max(c.y - R, 0); y <= min(c.y + R, g.dim[1] - 1); y++) {
            const unsigned row = (unsigned)((z * g.dim[1] + y) * g.dim[0]);
            run(v.cell_cnt[row + x0], v.cell_cnt[row + x1 + 1]);
          }
      }
    }
    __syncthreads();
    const unsigned staged = min(s_cnt, (unsigned)min(v.stage_limit, kPairsStage));
    if (threadIdx.x == 0 && staged) s_base = atomicAdd(&v.flags->n_pairs, (unsigned long long)staged);
    __syncthreads();
    for (unsigned e = threadIdx.x; e < staged; e += kPairsBlock) {
      const unsigned long long gs = s_base + e;
      if ((long long)gs < v.pairs_cap) v.pairs[gs] = s_pairs[e];
      else atomicExch(&v.flags->overflow_pairs, 1);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- 3. fixed point over timelines
struct BodyState { float px, py, pz, vx, vy, vz; };

// State of body b just before the visit with key (row, col): post-pass-3 state, then the latest
// timeline entry with a smaller key, then the body's own integration if its row lies between.
__device__ __forceinline__ BodyState state_at(const LargeView &v, int src, int b, unsigned long long key, float dt, float *radius) {
  const float4 p = v.pos_rest[b], q = v.vel_rad[b];
  BodyState s;
  s.px = p.x; s.py = p.y; s.pz = p.z; s.vx = q.x; s.vy = q.y; s.vz = q.z;
  *radius = q.w;
  long long last_row = -1;
  const int cnt = min((int)v.tl_cnt[src][b], kTimeline);
  unsigned long long best = 0;
  int best_k = -1;
  for (int k = 0; k < cnt; k++) {
    const unsigned long long kk = v.tl_key[src][(size_t)b * kTimeline + k];
    if (kk < key && (best_k < 0 || kk > best)) { best = kk; best_k = k; }
  }
  if (best_k >= 0) {
    const float4 tp = v.tl_pos[src][(size_t)b * kTimeline + best_k], tv = v.tl_vel[src][(size_t)b * kTimeline + best_k];
    s.px = tp.x; s.py = tp.y; s.pz = tp.z; s.vx = tv.x; s.vy = tv.y; s.vz = tv.z;
    last_row = (long long)(best >> 32);
  }
  const long long row = (long long)(key >> 32);
  if (row > b && last_row <= b && p.w == 0.0f) {            // own row finished in between: integrate (world.c:549-553)
    V3 pp = v3(s.px, s.py, s.pz), vv = v3(s.vx, s.vy, s.vz);
    pgm::integrate(pp, vv, dt);
    s.px = pp.x; s.py = pp.y; s.pz = pp.z; s.vx = vv.x; s.vy = vv.y; s.vz = vv.z;
  }
  return s;
}

__device__ __forceinline__ void timeline_push(const LargeView &v, int dst, int b, unsigned long long key, const Sphere &s) {
  // tl_cnt is a byte array; claim a slot with a 32-bit CAS on the containing word
  unsigned *word = (unsigned *)(v.tl_cnt[dst] + (b & ~3));
  const int sh = (b & 3) * 8;
  unsigned old = *word, assumed;
  int slot;
  do {
    assumed = old;
    slot = (int)((assumed >> sh) & 0xffu);
    if (slot >= 250) break;
    old = atomicCAS(word, assumed, assumed + (1u << sh));
  } while (old != assumed);
  if (slot >= kTimeline) { atomicExch(&v.flags->overflow_timeline, 1); return; }
  atomicOr(&v.pushed_bits[b >> 5], 1u << (b & 31));
  if (slot == 0 && atomicExch(&v.seen[b], 1) == 0) { touched_stage(v, b, kTouchAdd); atomicOr(&v.touched_bits[b >> 5], 1u << (b & 31)); }
  v.tl_key[dst][(size_t)b * kTimeline + slot] = key;
  v.tl_pos[dst][(size_t)b * kTimeline + slot] = make_float4(s.px, s.py, s.pz, 0.0f);
  v.tl_vel[dst][(size_t)b * kTimeline + slot] = make_float4(s.vx, s.vy, s.vz, 0.0f);
}

// Evaluate one visit on the hypothesis `src` WITHOUT touching the timelines: 0 = no event, 1 = event
// (A = row body, B = column body after it), 2 = deferred (one thread per visit only: the interval search
// is a long one, the visit goes to a block).  Warp: all threads of the block, identical arguments.
template <bool Warp>
__device__ __forceinline__ int visit_compute(const LargeView &v, int src, int row, int col, float dt, pgm::LeafLen leaf, Sphere &A, Sphere &B) {
  const unsigned long long key = ((unsigned long long)(unsigned)row << 32) | (unsigned)col;
  float ra, rb;
  const BodyState a = state_at(v, src, row, key, dt, &ra), b = state_at(v, src, col, key, dt, &rb);
  A.px = a.px; A.py = a.py; A.pz = a.pz; A.vx = a.vx; A.vy = a.vy; A.vz = a.vz; A.r = ra; A.rest = 0;
  B.px = b.px; B.py = b.py; B.pz = b.pz; B.vx = b.vx; B.vy = b.vy; B.vz = b.vz; B.r = rb; B.rest = 0;
  if (Warp) return pgs::sphere_pair_exact_block(A, B, dt, leaf) ? 1 : 0;
  return pgs::sphere_pair_exact_try(A, B, dt, leaf);
}

// Copy the entry with `key` (if any) of body b from the src to the dst timeline.
__device__ __forceinline__ bool carry_entry(const LargeView &v, int src, int dst, int b, unsigned long long key) {
  const int cnt = min((int)v.tl_cnt[src][b], kTimeline);
  for (int k = 0; k < cnt; k++) {
    if (v.tl_key[src][(size_t)b * kTimeline + k] != key) continue;
    const float4 tp = v.tl_pos[src][(size_t)b * kTimeline + k], tv = v.tl_vel[src][(size_t)b * kTimeline + k];
    Sphere S; S.px = tp.x; S.py = tp.y; S.pz = tp.z; S.vx = tv.x; S.vy = tv.y; S.vz = tv.z; S.r = 0.0f; S.rest = 0;
    timeline_push(v, dst, b, key, S);
    return true;
  }
  return false;
}

// Did anything body b contributes to a visit with key K change since the previous iteration?
// (entries with key < K: new or different state -> flagged in .w; gone -> tl_chg)
__device__ __forceinline__ bool affected(const LargeView &v, int src, int b, unsigned long long K) {
  const unsigned long long rem = v.tl_chg[src][b];
  if (rem < K) return true;
  const int cnt = min((int)v.tl_cnt[src][b], kTimeline);
  for (int k = 0; k < cnt; k++) {
    const unsigned long long kk = v.tl_key[src][(size_t)b * kTimeline + k];
    if (kk < K && v.tl_pos[src][(size_t)b * kTimeline + k].w != 0.0f) return true;
  }
  return false;
}

// Which candidate visits does this iteration have to look at?  A light, fully occupied pass over the
// whole list so that k_eval (long divergent evaluations) runs on dense warps only.
//   iteration 0   the timelines are empty: every visit sees the untouched states its two entries were listed
//                 with, and k_pairs has already put the closest-approach gate -- the very gate
//                 sphere_pair_exact opens with -- to those; a visit it rejected cannot fire.
//   later         visits with a DIRTY body -- one whose timeline changed in the previous iteration (k_check)
//                 -- or appended by k_requery.  Every visit of two clean bodies would come out as
//                 recorded; d_clear_carry has copied those entries already.
// The order of the active list is irrelevant (timeline slots are claimed atomically anyway).
constexpr int kSelItems = 4;
__device__ __forceinline__ void d_select(const LargeView &v, int src, int iter, float dt) {
  __shared__ unsigned s_cnt, s_base;
  const int kSelBlock = (int)blockDim.x;
  const long long np = (long long)min((unsigned long long)v.pairs_cap, v.flags->n_pairs);
  const long long tile = (long long)kSelBlock * kSelItems;
  for (long long k0 = blockIdx.x * tile; k0 < np; k0 += (long long)gridDim.x * tile) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    unsigned need_bits = 0;
#pragma unroll
    for (int q = 0; q < kSelItems; q++) {
      const long long k = k0 + q * kSelBlock + threadIdx.x;
      if (k >= np) continue;
      const int2 pr = v.pairs[k];
      const bool fresh = pr.x < 0;
      const int i = fresh ? ~pr.x : pr.x, j = pr.y & 0x7fffffff;
      bool need = fresh;
      if (!need && iter > 0) need = ((v.dirty_bits[src][i >> 5] >> (i & 31)) & 1u) || ((v.dirty_bits[src][j >> 5] >> (j & 31)) & 1u);
      else if (!need) need = pr.y >= 0;            // iteration 0: the gate's verdict from k_pairs
      if (need) need_bits |= 1u << q;
    }
    unsigned slot = 0;
    const int mine = __popc(need_bits);
    if (mine) slot = atomicAdd(&s_cnt, (unsigned)mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) s_base = atomicAdd(&v.flags->n_active, s_cnt);
    __syncthreads();
    if (mine) {
      slot += s_base;
#pragma unroll
      for (int q = 0; q < kSelItems; q++)
        if ((need_bits >> q) & 1u) v.active[slot++] = (unsigned)(k0 + q * kSelBlock + threadIdx.x);
    }
    __syncthreads();                               // s_base is rewritten by the next tile
  }
}

// Candidate visit k on the hypothesis `src`; a hit goes to the `dst` timelines of both bodies.  Returns the
// number of hits (0 / 1), or -1 when a one-thread evaluation met a long interval search: nothing has been
// written then and the visit is to be repeated by a whole block (Warp = true: all its threads run this
// with the same k, `writer` = thread 0 does the writes).
template <bool Warp>
__device__ __forceinline__ int eval_pair(const LargeView &v, int src, int dst, int iter, float dt, pgm::LeafLen leaf, long long k, bool writer) {
  const int2 pr = v.pairs[k];
  const bool fresh = pr.x < 0;                     // appended by k_requery: never evaluated yet
  const int row = fresh ? ~pr.x : pr.x, col = pr.y & 0x7fffffff;
  const unsigned long long key = ((unsigned long long)(unsigned)row << 32) | (unsigned)col;
  Sphere A, B;
  if (iter > 0 && !fresh) {
    // neither body has a timeline: the visit sees the untouched states, like in iteration 0
    if (!((v.touched_bits[row >> 5] >> (row & 31)) & 1u) && !((v.touched_bits[col >> 5] >> (col & 31)) & 1u)) return 0;
    // A visit depends only on the entries with SMALLER keys of its two bodies.  If none of those
    // appeared, changed or disappeared since the previous iteration, its outcome is the recorded
    // one: carry it over instead of re-evaluating.
    if (!(affected(v, src, row, key) || affected(v, src, col, key))) {
      int hits = 0;
      if (writer && v.tl_cnt[src][row] != 0 && v.tl_cnt[src][col] != 0 && carry_entry(v, src, dst, row, key)) { carry_entry(v, src, dst, col, key); hits = 1; }
      return hits;
    }
  }
  const int h = visit_compute<Warp>(v, src, row, col, dt, leaf, A, B);
  if (h == 2) return -1;
  if (writer) {
    if (fresh) v.pairs[k].x = row;
    if (h) { timeline_push(v, dst, row, key, A); timeline_push(v, dst, col, key, B); }
  }
  return h;
}

__device__ __forceinline__ void d_eval(const LargeView &v, int src, int dst, int iter, float dt) {
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;
  const long long na = (long long)v.flags->n_active;
  unsigned my_hits = 0;
  touched_stage(v, 0, kTouchInit);
  for (long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x; a < na; a += (long long)gridDim.x * blockDim.x) {
    const unsigned k = v.active[a];
    const int h = eval_pair<false>(v, src, dst, iter, dt, leaf, (long long)k, true);
    if (h >= 0) my_hits += (unsigned)h;
    else v.active[v.pairs_cap + atomicAdd(&v.flags->n_hard, 1u)] = k;      // the hard list: second half of active[]
  }
  // one atomic per warp for the hit statistic (a single hot address otherwise)
  my_hits = __reduce_add_sync(kFullMask, my_hits);
  if ((threadIdx.x & 31) == 0 && my_hits) atomicAdd(&v.flags->n_hits, (unsigned long long)my_hits);
  touched_stage(v, 0, kTouchFlush);
}
// The pairs k_eval deferred (grazing pairs: an interval search of hundreds to thousands of nodes that a
// single thread would walk for up to a millisecond while the rest of the GPU waits; a handful per step):
// one BLOCK per pair, all threads evaluate it redundantly and share the leaf scan of pgm::ss_interval.
__device__ __forceinline__ void d_eval_hard(const LargeView &v, int src, int dst, int iter, float dt) {
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;
  const int nh = (int)v.flags->n_hard;
  touched_stage(v, 0, kTouchInit);
  for (int t = blockIdx.x; t < nh; t += gridDim.x) {
    const long long k = (long long)v.active[v.pairs_cap + t];
    const int h = eval_pair<true>(v, src, dst, iter, dt, leaf, k, threadIdx.x == 0);
    if (threadIdx.x == 0 && h > 0) atomicAdd(&v.flags->n_hits, (unsigned long long)h);
    __syncthreads();
  }
  touched_stage(v, 0, kTouchFlush);
}

// Compare the timelines of two iterations (as sets keyed by visit) and validate envelopes, for the
// bodies that have ever been touched this step.
__device__ __forceinline__ bool same_state(float4 p0, float4 q0, float4 p1, float4 q1) {
  return __float_as_uint(p0.x) == __float_as_uint(p1.x) && __float_as_uint(p0.y) == __float_as_uint(p1.y) &&
         __float_as_uint(p0.z) == __float_as_uint(p1.z) && __float_as_uint(q0.x) == __float_as_uint(q1.x) &&
         __float_as_uint(q0.y) == __float_as_uint(q1.y) && __float_as_uint(q0.z) == __float_as_uint(q1.z);
}
__device__ __forceinline__ void d_check(const LargeView &v, int src, int dst, int iter, float dt) {
  const int nt = (int)v.flags->n_touched;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    v.flags->n_hits += v.flags->n_carried; v.flags->n_carried = 0ull;
    v.flags->n_active = 0u; v.flags->n_hard = 0u;      // for the next iteration's k_begin (zero at step start)
    v.flags->balls_begin = v.flags->n_balls;           // this iteration's state balls start here (k_states_insert)
  }
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int b = v.touched[t];
    // a clean body that k_eval gave nothing holds exactly the entries d_clear_carry copied: same timeline,
    // flags already in their rest state, envelope validated when the entries were new
    if (iter > 0 && !((v.dirty_bits[src][b >> 5] >> (b & 31)) & 1u) && !((v.pushed_bits[b >> 5] >> (b & 31)) & 1u)) continue;
    const int c0 = min((int)v.tl_cnt[src][b], kTimeline), c1 = min((int)v.tl_cnt[dst][b], kTimeline);
    unsigned long long rem = 0xffffffffffffffffull;
    bool any = false;
    for (int k = 0; k < c1; k++) {
      const unsigned long long key = v.tl_key[dst][(size_t)b * kTimeline + k];
      bool found = false;
      for (int m = 0; m < c0; m++) {
        if (v.tl_key[src][(size_t)b * kTimeline + m] != key) continue;
        found = same_state(v.tl_pos[src][(size_t)b * kTimeline + m], v.tl_vel[src][(size_t)b * kTimeline + m],
                           v.tl_pos[dst][(size_t)b * kTimeline + k], v.tl_vel[dst][(size_t)b * kTimeline + k]);
        break;
      }
      v.tl_pos[dst][(size_t)b * kTimeline + k].w = found ? 0.0f : 1.0f;      // new or changed entry
      any = any || !found;
    }
    for (int m = 0; m < c0; m++) {
      const unsigned long long key = v.tl_key[src][(size_t)b * kTimeline + m];
      bool found = false;
      for (int k = 0; k < c1; k++) if (v.tl_key[dst][(size_t)b * kTimeline + k] == key) { found = true; break; }
      if (!found && key < rem) rem = key;
    }
    if (v.tl_cnt[src][b] != v.tl_cnt[dst][b] && rem == 0xffffffffffffffffull && !any) rem = 0ull;      // overflowed counts: be safe
    v.tl_chg[dst][b] = rem;
    if (any || rem != 0xffffffffffffffffull) { v.flags->changed = 1; atomicOr(&v.dirty_bits[dst][b >> 5], 1u << (b & 31)); }
  }
}
// Start of an iteration: the destination timelines of the touched bodies.  A DIRTY body (its timeline
// changed in the previous iteration) starts empty: every pair it belongs to is on the selection's list and
// k_eval rebuilds its entries.  A clean body keeps, by plain copy, the entries whose partner is clean as
// well -- visits nothing has changed for; its entries with a dirty partner come back through k_eval.
// From the third iteration on almost everybody is clean, and an iteration costs what its few changes cost.
__device__ __forceinline__ void d_clear_carry(const LargeView &v, int src, int dst, int iter) {
  const int nt = (int)v.flags->n_touched;
  if (blockIdx.x == 0 && threadIdx.x == 0) { v.flags->changed = 0; v.flags->n_hits = 0ull; }      // (n_active, n_hard: k_check)
  unsigned carried = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int b = v.touched[t];
    const unsigned bit = 1u << (b & 31);
    atomicAnd(&v.dirty_bits[dst][b >> 5], ~bit);
    atomicAnd(&v.pushed_bits[b >> 5], ~bit);
    int m = 0;
    if (iter > 0 && !(v.dirty_bits[src][b >> 5] & bit)) {
      const int cnt = min((int)v.tl_cnt[src][b], kTimeline);
      for (int k = 0; k < cnt; k++) {
        const unsigned long long key = v.tl_key[src][(size_t)b * kTimeline + k];
        const int row = (int)(key >> 32), col = (int)(key & 0xffffffffu);
        const int partner = row == b ? col : row;
        if ((v.dirty_bits[src][partner >> 5] >> (partner & 31)) & 1u) {
          // k_eval decides about this entry (it may not come back): k_check has to look at this body
          atomicOr(&v.pushed_bits[b >> 5], bit);
          continue;
        }
        const float4 tp = v.tl_pos[src][(size_t)b * kTimeline + k];
        v.tl_key[dst][(size_t)b * kTimeline + m] = key;
        v.tl_pos[dst][(size_t)b * kTimeline + m] = make_float4(tp.x, tp.y, tp.z, 0.0f);
        v.tl_vel[dst][(size_t)b * kTimeline + m] = v.tl_vel[src][(size_t)b * kTimeline + k];
        m++;
        if (row == b) carried++;                   // one hit per visit: counted by its row body
      }
      v.tl_chg[dst][b] = 0xffffffffffffffffull;
    }
    v.tl_cnt[dst][b] = (unsigned char)m;
  }
  carried = __reduce_add_sync(kFullMask, carried);
  if ((threadIdx.x & 31) == 0 && carried) atomicAdd(&v.flags->n_carried, (unsigned long long)carried);   // (n_hits is being reset by this very kernel)
}

// ---------------------------------------------------------------- 4. state balls
// Every state that appeared (or changed) in a timeline this iteration gets its ball; so does, for a body whose
// timeline changed at all, what its own row's integration makes of its last phase-A state.  One thread per
// touched body.  Balls are never removed: a state that disappears again only leaves the list a superset.
__global__ void k_states_insert(LargeView v, int src, int dst, float dt) {
  if (v.flags->stop) return;
  const GridParams g = *v.grid;
  const int nt = (int)v.flags->n_touched;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int b = v.touched[t];
    if (!((v.dirty_bits[dst][b >> 5] >> (b & 31)) & 1u)) continue;
    const float4 p = v.pos_rest[b];
    const float r = v.vel_rad[b].w;
    const bool at_rest = p.w != 0.0f;
    const unsigned rest_flag = at_rest ? kRestBit : 0u;
    const int cnt = min((int)v.tl_cnt[dst][b], kTimeline);
    unsigned long long lastA_key = 0ull;
    int lastA = -1;
    for (int k = 0; k < cnt; k++) {
      const unsigned long long key = v.tl_key[dst][(size_t)b * kTimeline + k];
      const bool phase_a = (long long)(key >> 32) <= (long long)b;      // rows up to the body's own see it before its integration
      if (phase_a && (lastA < 0 || key > lastA_key)) { lastA = k; lastA_key = key; }
      const float4 tp = v.tl_pos[dst][(size_t)b * kTimeline + k];
      if (tp.w == 0.0f) continue;                                        // unchanged since the previous iteration: has its ball
      const float4 tv = v.tl_vel[dst][(size_t)b * kTimeline + k];
      const int ph = (phase_a || at_rest) ? 0 : 1;
      ball_insert(v, g, entry_mid(tp.x, tp.y, tp.z, tv.x, tv.y, tv.z, dt), entry_rho(r, tp.x, tp.y, tp.z, tv.x, tv.y, tv.z, dt),
                  (unsigned)(2 * b + ph) | rest_flag);
    }
    if (lastA >= 0 && !at_rest) {
      const float4 tp = v.tl_pos[dst][(size_t)b * kTimeline + lastA], tv = v.tl_vel[dst][(size_t)b * kTimeline + lastA];
      // ... unless the previous iteration's timeline ended its phase A in the very same state (that ball exists)
      const int cnt0 = min((int)v.tl_cnt[src][b], kTimeline);
      unsigned long long prevA_key = 0ull;
      int prevA = -1;
      for (int k = 0; k < cnt0; k++) {
        const unsigned long long key = v.tl_key[src][(size_t)b * kTimeline + k];
        if ((long long)(key >> 32) <= (long long)b && (prevA < 0 || key > prevA_key)) { prevA = k; prevA_key = key; }
      }
      if (prevA >= 0 && same_state(v.tl_pos[src][(size_t)b * kTimeline + prevA], v.tl_vel[src][(size_t)b * kTimeline + prevA], tp, tv)) continue;
      V3 pp = v3(tp.x, tp.y, tp.z), vv = v3(tv.x, tv.y, tv.z);
      pgm::integrate(pp, vv, dt);
      ball_insert(v, g, entry_mid(pp.x, pp.y, pp.z, vv.x, vv.y, vv.z, dt), entry_rho(r, pp.x, pp.y, pp.z, vv.x, vv.y, vv.z, dt),
                  (unsigned)(2 * b + 1));
    }
  }
}

// One group of PG_QUERY_LANES lanes per ball of the current iteration: which entries of other bodies does it touch, and which visits does
// that add?  Three sources of partners: the cell-sorted untouched entries (27 cells: the lanes take the entries of
// each x-run), the state balls and large untouched entries linked in the second structure (a lane per cell, walking
// its list), and the few balls beyond e_cap2 (list).  A visit the untouched entries of the two bodies already listed
// (k_pairs: the same ball test, repeated here bit for bit) is skipped; the rest go through the hash set.
// G lanes per ball.  The query is a chain of dependent loads: with many balls to look up (the first iterations of a
// hot step: a million) more balls in flight beat more lanes per ball (G = 8); with few (the long tail of iterations,
// or a small world) the kernel lasts as long as ONE query, and a whole warp per ball shortens that chain (G = 32).
template <unsigned G>
__device__ __forceinline__ void d_states_query(const LargeView &v, float dt) {
  const GridParams g = *v.grid;
  const unsigned lane = threadIdx.x & (G - 1u);                          // lane within the ball's group
  const unsigned gmask = G == 32u ? 0xffffffffu : (((1u << G) - 1u) << (threadIdx.x & 31u & ~(G - 1u)));
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) / G, n_warps = (gridDim.x * blockDim.x) / G;   // group index / count
  const unsigned j0 = v.flags->balls_begin, j1 = min(v.flags->n_balls, (unsigned)v.balls_cap);
  for (unsigned j = j0 + warp; j < j1; j += n_warps) {
    const float4 bc = v.balls[j].c;
    const unsigned bflags = v.balls[j].ent;
    const int e = (int)(bflags & ~kRestBit), b = e >> 1;
    const bool rest_b = (bflags & kRestBit) != 0u;
    // the body's own untouched entry of the same phase (for the "already listed" test)
    float c0[3], v0[3];
    entry_state(v.pos_rest[b], v.vel_rad[b], e & 1, dt, c0, v0);
    const float3 m0 = entry_mid(c0[0], c0[1], c0[2], v0[0], v0[1], v0[2], dt);
    const float rho0 = v.env[e];
    auto consider = [&](int ex, bool rest_x, float3 mx, float rhox, bool partner_is_untouched_entry) {
      if ((ex >> 1) == b) return;
      const float dx = bc.x - mx.x, dyy = bc.y - mx.y, dzz = bc.z - mx.z;
      const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dyy, dyy, dzz * dzz));
      const float T = bc.w + rhox;
      if (!(d2 <= T * T)) return;
      int2 vis[2]; bool row_b[2];
      const int nv = entry_pair_visits((unsigned)e, rest_b, (unsigned)ex, rest_x, vis, row_b);
      if (nv == 0) return;
      // did k_pairs list the visits of this entry pair?  (untouched balls of both entries touch)
      float3 ux = mx; float urho = rhox;
      if (!partner_is_untouched_entry) {
        const int x = ex >> 1;
        urho = v.env[ex];
        if (urho >= 0.0f) {
          float cx[3], vx[3];
          entry_state(v.pos_rest[x], v.vel_rad[x], ex & 1, dt, cx, vx);
          ux = entry_mid(cx[0], cx[1], cx[2], vx[0], vx[1], vx[2], dt);
        }
      }
      if (rho0 >= 0.0f && urho >= 0.0f) {
        const float ex_ = m0.x - ux.x, ey_ = m0.y - ux.y, ez_ = m0.z - ux.z;
        const float u2 = __fmaf_rn(ex_, ex_, __fmaf_rn(ey_, ey_, ez_ * ez_));
        const float U = rho0 + urho;
        if (u2 <= U * U) return;
      }
      for (int q = 0; q < nv; q++) {
        const unsigned long long key = ((unsigned long long)(unsigned)vis[q].x << 32) | (unsigned)vis[q].y;
        const long long hs = vset_insert(v, key);
        if (hs < 0) continue;
        const unsigned long long slot = agg_inc64(&v.flags->n_pairs);
        if ((long long)slot < v.pairs_cap) v.pairs[slot] = make_int2(~vis[q].x, vis[q].y);   // ~row marks it fresh
        else { atomicExch(&v.flags->overflow_pairs, 1); v.vset[hs] = kSlotTomb; }            // (the host makes room and repeats the query)
      }
    };
    // (1) untouched entries within e_cap of the grid's sizing: rho + e_cap away at most
    {
      const float need = (bc.w + g.e_cap) * g.inv_h * 1.001f;
      if (!(need <= (float)kMaxSearch)) {
        const unsigned n_sorted = v.cell_cnt[g.n_cells];
        for (unsigned a = lane; a < n_sorted; a += G) {
          const float4 cx = v.sorted[a].crho;
          const float4 wx = *reinterpret_cast<const float4 *>(&v.sorted[a].vx);
          const unsigned idx = __float_as_uint(wx.w);
          if (cx.w > g.e_cap) continue;
          consider((int)(idx & ~kRestBit), (idx & kRestBit) != 0u, entry_mid(cx.x, cx.y, cx.z, wx.x, wx.y, wx.z, dt), cx.w, true);
        }
      } else {
        const int R = max((int)ceilf(need), 1);
        const int3 cc = cell_coords(g, bc.x, bc.y, bc.z);
        const int side = 2 * R + 1;
        const int x0 = max(cc.x - R, 0), x1 = min(cc.x + R, g.dim[0] - 1);
        // the bounds of up to 32 x-runs at a time, a lane each (independent loads), then the runs one after the
        // other with the lanes on their entries
        for (int r0 = 0; r0 < side * side; r0 += (int)G) {
          unsigned my_a0 = 0u, my_a1 = 0u;
          {
            const int r = r0 + (int)lane;
            const int y = cc.y + r % side - R, z = cc.z + r / side - R;
            if (r < side * side && y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
              const unsigned row = (unsigned)((z * g.dim[1] + y) * g.dim[0]);
              my_a0 = v.cell_cnt[row + x0]; my_a1 = v.cell_cnt[row + x1 + 1];
            }
          }
          const int n_here = min((int)G, side * side - r0);
          for (int q = 0; q < n_here; q++) {
            const unsigned a0 = __shfl_sync(gmask, my_a0, q, G), a1 = __shfl_sync(gmask, my_a1, q, G);
            for (unsigned a = a0 + lane; a < a1; a += G) {
              const float4 cx = v.sorted[a].crho;
              if (cx.w > g.e_cap) continue;                  // (linked in the second structure as well: found there)
              const float4 wx = *reinterpret_cast<const float4 *>(&v.sorted[a].vx);
              const unsigned idx = __float_as_uint(wx.w);
              consider((int)(idx & ~kRestBit), (idx & kRestBit) != 0u, entry_mid(cx.x, cx.y, cx.z, wx.x, wx.y, wx.z, dt), cx.w, true);
            }
          }
        }
      }
    }
    // (2) balls of the second structure: the bulk level (radius <= e_cap) rho + e_cap away at most, the tail level
    // (radius <= e_cap2) rho + e_cap2; a lane per cell, walking its list
    const unsigned tail_o = v.flags->tail_max;      // (insertions and queries never share a kernel)
    for (int level = 0; level < (tail_o ? 2 : 1); level++) {
      const unsigned *head = level ? v.head3 : v.head2;
      const float need = (bc.w + (level ? fminf(g.e_cap2, o2f(tail_o)) : g.e_cap)) * g.inv_h * 1.001f;
      const int R = min(max((int)ceilf(need), 1), 64);
      const int3 cc = cell_coords(g, bc.x, bc.y, bc.z);
      const int side = 2 * R + 1;
      const int n_cells_q = side * side * side;
      for (int q = (int)lane; q < n_cells_q; q += (int)G) {
        const int x = cc.x + q % side - R, y = cc.y + (q / side) % side - R, z = cc.z + q / (side * side) - R;
        if (x < 0 || x >= g.dim[0] || y < 0 || y >= g.dim[1] || z < 0 || z >= g.dim[2]) continue;
        unsigned k = head[(unsigned)((z * g.dim[1] + y) * g.dim[0] + x)];
        while (k != kNoBall) {
          // both halves of the node in flight together (the walk is bound by the latency of its dependent loads)
          const float4 kc = v.balls[k].c;
          const uint2 ke = *reinterpret_cast<const uint2 *>(&v.balls[k].ent);
          if (k != j) consider((int)(ke.x & ~kRestBit), (ke.x & kRestBit) != 0u, make_float3(kc.x, kc.y, kc.z), kc.w, false);
          k = ke.y;
        }
      }
    }
    // (3) the few balls beyond e_cap2
    {
      const unsigned nh = min(v.flags->n_huge, (unsigned)v.huge_cap);
      for (unsigned h = lane; h < nh; h += G) {
        const unsigned k = (unsigned)v.huge_list[h];
        if (k == j) continue;
        const float4 kc = v.balls[k].c;
        const unsigned kx = v.balls[k].ent;
        consider((int)(kx & ~kRestBit), (kx & kRestBit) != 0u, make_float3(kc.x, kc.y, kc.z), kc.w, false);
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_states_query(LargeView v, float dt) {
  if (v.flags->stop) return;
  const unsigned n_new = min(v.flags->n_balls, (unsigned)v.balls_cap) - v.flags->balls_begin;
  if (n_new * 8u >= gridDim.x * blockDim.x) d_states_query<8u>(v, dt);      // enough balls to fill the grid eight lanes each
  else d_states_query<32u>(v, dt);
}

// Rehash the visit set into a larger table (host: when the load passes 1/4).
__global__ void k_vset_rehash(LargeView v, const unsigned long long *old_tab, unsigned long long old_slots) {
  for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < old_slots; s += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long key = old_tab[s];
    if (key == kSlotEmpty || key == kSlotTomb) continue;
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    unsigned long long slot = h & v.vset_mask;
    while (atomicCAS(&v.vset[slot], kSlotEmpty, key) != kSlotEmpty) slot = (slot + 1) & v.vset_mask;
  }
}


// Several iterations are queued per host read-back; this one-thread kernel opens each of them: once the fixed
// point has converged -- or something needs the host (a buffer to grow, an error) -- it raises `stop`, and
// every kernel queued behind it returns at once.
__global__ void k_iter_gate(LargeView v, int first) {
  Flags *f = v.flags;
  if (f->stop) return;
  if (!first && !f->changed) { f->stop = 1; f->converged = 1; return; }
  if (f->overflow_pairs || f->overflow_timeline || f->overflow_balls || f->overflow_vset || f->overflow_huge || f->unsafe ||
      (unsigned long long)f->n_vset * 4ull > v.vset_mask + 1ull || (long long)f->n_balls * 2 > (long long)v.balls_cap) { f->stop = 1; return; }
  f->iters_done++;
}
#ifndef PG_EVAL_MIN_BLOCKS
#define PG_EVAL_MIN_BLOCKS 6
#endif
__global__ void __launch_bounds__(128, PG_EVAL_MIN_BLOCKS) k_eval(LargeView v, int src, int dst, int iter, float dt) { if (v.flags->stop) return; d_eval(v, src, dst, iter, dt); }
__global__ void __launch_bounds__(512) k_eval_hard(LargeView v, int src, int dst, int iter, float dt) { if (v.flags->stop) return; d_eval_hard(v, src, dst, iter, dt); }
__global__ void k_check(LargeView v, int src, int dst, int iter, float dt) { if (v.flags->stop) return; d_check(v, src, dst, iter, dt); }
// One launch opens an iteration: the carry (per touched body) and the selection (per candidate pair) read the
// same source hypothesis and write disjoint things.
__global__ void __launch_bounds__(256) k_begin(LargeView v, int src, int dst, int iter, float dt) {
  if (v.flags->stop) return;
  d_clear_carry(v, src, dst, iter);
  d_select(v, src, iter, dt);
}

// ---------------------------------------------------------------- 5. final state + pass-4 events
__global__ void k_finalize(LargeView v, int src, float dt, int step) {
  const int stride = gridDim.x * blockDim.x;
  event_stage(v, kEventInit, 0, 0, 0, 0, 0, 0);
  // two bodies per trip: the second one's loads are issued before the first one's stores
  float4 p_next = make_float4(0, 0, 0, 0), q_next = p_next;
  int cnt_next = 0;
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < v.n) { p_next = v.pos_rest[b]; q_next = v.vel_rad[b]; cnt_next = v.tl_cnt[src][b]; }
  for (; b < v.n; b += stride) {
    float4 p = p_next, q = q_next;
    const int cnt = min(cnt_next, kTimeline);
    if (b + stride < v.n) { p_next = v.pos_rest[b + stride]; q_next = v.vel_rad[b + stride]; cnt_next = v.tl_cnt[src][b + stride]; }
    long long last_row = -1;
    if (cnt > 0) {
      unsigned long long best = 0; int best_k = -1;
      for (int k = 0; k < cnt; k++) {
        const unsigned long long kk = v.tl_key[src][(size_t)b * kTimeline + k];
        if (best_k < 0 || kk > best) { best = kk; best_k = k; }
        if ((int)(kk >> 32) == b) emit_event(v, step, 4, PG_CLASS_DYNAMIC, b, PG_CLASS_DYNAMIC, (int)(kk & 0xffffffffu));
      }
      const float4 tp = v.tl_pos[src][(size_t)b * kTimeline + best_k], tv = v.tl_vel[src][(size_t)b * kTimeline + best_k];
      p.x = tp.x; p.y = tp.y; p.z = tp.z; q.x = tv.x; q.y = tv.y; q.z = tv.z;
      last_row = (long long)(best >> 32);
    }
    if (last_row <= b && p.w == 0.0f) {
      V3 pp = v3(p.x, p.y, p.z), vv = v3(q.x, q.y, q.z);
      pgm::integrate(pp, vv, dt);
      p.x = pp.x; p.y = pp.y; p.z = pp.z; q.x = vv.x; q.y = vv.y; q.z = vv.z;
    }
    v.pos_rest[b] = p;
    v.vel_rad[b] = q;
  }
  event_stage(v, kEventFlush, 0, 0, 0, 0, 0, 0);
}

// End of step: the bodies on the touched list are the only ones whose fixed-point bookkeeping left its
// rest state; put it back so that the next k_pass3 does not have to rewrite it for every body.
__global__ void k_reset_touched(LargeView v) {
  const int nt = (int)v.flags->n_touched;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int b = v.touched[t];
    v.tl_cnt[0][b] = 0;
    v.tl_cnt[1][b] = 0;
    v.tl_chg[0][b] = 0xffffffffffffffffull;
    v.tl_chg[1][b] = 0xffffffffffffffffull;
    v.seen[b] = 0;
    const unsigned bit = 1u << (b & 31);
    atomicAnd(&v.dirty_bits[0][b >> 5], ~bit);
    atomicAnd(&v.dirty_bits[1][b >> 5], ~bit);
    atomicAnd(&v.pushed_bits[b >> 5], ~bit);
  }
}

// ---------------------------------------------------------------- frozen-state broadphase pairs
__global__ void k_env_frozen(LargeView v, float dt) {
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f}, emax = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
    const float4 p = v.pos_rest[i], q = v.vel_rad[i];
    // nothing moves between the visits of a frozen-state query: phase A entries only
    const float e = entry_rho(q.w, p.x, p.y, p.z, q.x, q.y, q.z, dt);
    const float3 m = entry_mid(p.x, p.y, p.z, q.x, q.y, q.z, dt);
    reinterpret_cast<float2 *>(v.env)[i] = make_float2(e, -1.0f);
    lo[0] = fminf(lo[0], m.x); lo[1] = fminf(lo[1], m.y); lo[2] = fminf(lo[2], m.z);
    hi[0] = fmaxf(hi[0], m.x); hi[1] = fmaxf(hi[1], m.y); hi[2] = fmaxf(hi[2], m.z);
    emax = fmaxf(emax, e);
  }
  for (int o = 16; o > 0; o >>= 1) {
    for (int k = 0; k < 3; k++) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(kFullMask, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFullMask, hi[k], o));
    }
    emax = fmaxf(emax, __shfl_xor_sync(kFullMask, emax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    for (int k = 0; k < 3; k++) { atomicMin(v.bbox + k, f2o(lo[k])); atomicMax(v.bbox + 3 + k, f2o(hi[k])); }
    atomicMax(v.bbox + 6, f2o(emax));
  }
}
__global__ void k_pairs_exact(LargeView v, float dt, unsigned char *keep) {
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;
  const long long np = (long long)min((unsigned long long)v.pairs_cap, v.flags->n_pairs);
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < np; k += (long long)gridDim.x * blockDim.x) {
    int2 pr = v.pairs[k];
    if (pr.y < 0 || !(pr.x < pr.y)) { keep[k] = 0; continue; }      // the gate rejected it / the second visit of a pair whose lower body rests
    const float4 pa = v.pos_rest[pr.x], qa = v.vel_rad[pr.x], pb = v.pos_rest[pr.y], qb = v.vel_rad[pr.y];
    const V3 cA = v3(0.0f + pa.x, 0.0f + pa.y, 0.0f + pa.z), cB = v3(0.0f + pb.x, 0.0f + pb.y, 0.0f + pb.z);
    const V3 vA = v3(qa.x, qa.y, qa.z), vB = v3(qb.x, qb.y, qb.z);
    keep[k] = pgm::ss_interval(cA, vA, pgm::norm(vA), cB, vB, pgm::norm(vB), qa.w + qb.w, 0.0f, dt, leaf) ? 1 : 0;
  }
}

}  // namespace

// ==================================================================== host side

struct PgLarge {
  int device;
  int cap_n, cap_static;
  LargeView v;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  cudaEvent_t kev[K_COUNT][2];
  float kms[K_COUNT];
  long long launches;
  int sm_count;
  double *d_stats;
  float *d_stage_pos, *d_stage_vel;
  unsigned char *d_stage_rest;
  int last_iterations, last_proven;
  PgWorlds *player_world;        // players x level geometry (passes 1-2) run through the any-collider kernel
  int n_players;
  int cap_static_total;
  long long last_candidates, last_hits;
  int step_in_call;
  int bookkeeping_clean;         // the last step ended with k_reset_touched: k_pass3 may skip the per-body resets
  int iter_batch;                // fixed-point iterations queued per host read-back
  int timing;                    // record a CUDA event pair around every kernel (pg_large_kernel_ms); off by default:
                                 // the pipeline of a small world is launch-bound and the events double its API calls
};

#define LW_LAUNCH(w, kid, kern, grid, block, ...)                                 \
  do {                                                                            \
    if ((w)->timing) cudaEventRecord((w)->kev[kid][0], (w)->stream);              \
    kern<<<(grid), (block), 0, (w)->stream>>>(__VA_ARGS__);                       \
    if ((w)->timing) cudaEventRecord((w)->kev[kid][1], (w)->stream);              \
    (w)->launches++;                                                              \
  } while (0)

static int lw_grid(const PgLarge *w, int n, int block) { return std::max(1, std::min((n + block - 1) / block, w->sm_count * 8)); }

static void lw_accumulate(PgLarge *w, int kid) {
  if (!w->timing) return;
  float ms = 0.0f;
  if (cudaEventElapsedTime(&ms, w->kev[kid][0], w->kev[kid][1]) == cudaSuccess) w->kms[kid] += ms;
}

// build grid + candidate pairs from v.env / current positions
static int lw_build_pairs(PgLarge *w, float dt) {
  LargeView &v = w->v;
  k_grid_params<<<1, 1, 0, w->stream>>>(v);
  w->launches++;
  PG_CUDA(cudaMemsetAsync(v.cell_cnt, 0, sizeof(unsigned) * ((size_t)v.cells_cap + 1), w->stream));
  PG_CUDA(cudaMemsetAsync(&v.flags->n_pairs, 0, sizeof(unsigned long long), w->stream));
  LW_LAUNCH(w, K_CELLS, k_cell_count, lw_grid(w, 2 * v.n, 256), 256, v, dt);
  // no read-back of the grid: the scan is launched for the largest table the buffers hold (blocks past
  // n_cells see zeros), and a grid that could not be built raises flags->unsafe for the next read-back
  const int scan_blocks = (v.cells_cap + 1 + kScanBlock * kScanItems - 1) / (kScanBlock * kScanItems);
  if (w->timing) cudaEventRecord(w->kev[K_SCAN][0], w->stream);
  k_scan_local<<<scan_blocks, kScanBlock, 0, w->stream>>>(v);
  k_scan_sums<<<1, 1024, 0, w->stream>>>(v, scan_blocks);
  k_scan_add<<<scan_blocks, kScanBlock, 0, w->stream>>>(v);
  if (w->timing) cudaEventRecord(w->kev[K_SCAN][1], w->stream);
  w->launches += 3;
  LW_LAUNCH(w, K_SCATTER, k_scatter, lw_grid(w, 2 * v.n, 256), 256, v, dt);
  LW_LAUNCH(w, K_PAIRS, k_pairs, lw_grid(w, 2 * v.n, kPairsBlock), kPairsBlock, v, dt);
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

static int lw_read_flags(PgLarge *w, Flags *f) {
  PG_CUDA(cudaMemcpyAsync(f, w->v.flags, sizeof(Flags), cudaMemcpyDeviceToHost, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  return PG_OK;
}

// The candidate list outgrew its buffer (dense piles: every sphere touches a dozen neighbours).
// Reallocate for `needed` pairs, keeping the first `keep` entries, and clear the overflow mark.
static int lw_grow_pairs(PgLarge *w, unsigned long long needed, unsigned long long keep) {
  LargeView &v = w->v;
  const long long cap = (long long)(needed + needed / 4 + 65536);
  int2 *bigger = nullptr;
  if (cudaMalloc(&bigger, sizeof(int2) * (size_t)cap) != cudaSuccess) {
    cudaGetLastError();
    pg_set_error("large world: cannot grow the candidate pair buffer to %lld pairs", cap);
    return PG_ERR_CAPACITY;
  }
  if (keep) PG_CUDA(cudaMemcpyAsync(bigger, v.pairs, sizeof(int2) * (size_t)keep, cudaMemcpyDeviceToDevice, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  cudaFree(v.pairs);
  v.pairs = bigger;
  v.pairs_cap = cap;
  cudaFree(v.active);
  v.active = nullptr;
  if (cudaMalloc(&v.active, sizeof(unsigned) * 2 * (size_t)cap) != cudaSuccess) {
    cudaGetLastError();
    pg_set_error("large world: cannot grow the active pair list to %lld entries", cap);
    return PG_ERR_CAPACITY;
  }
  PG_CUDA(cudaMemsetAsync(&v.flags->overflow_pairs, 0, sizeof(int), w->stream));
  PG_CUDA(cudaMemcpyAsync(&v.flags->n_pairs, &keep, sizeof(keep), cudaMemcpyHostToDevice, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  return PG_OK;
}

// Four times the slots for the visit hash set, keys carried over.
static int lw_grow_vset(PgLarge *w) {
  LargeView &v = w->v;
  const unsigned long long old_slots = v.vset_mask + 1ull, new_slots = old_slots * 4ull;
  unsigned long long *bigger = nullptr, *old = v.vset;
  if (cudaMalloc(&bigger, sizeof(unsigned long long) * new_slots) != cudaSuccess) {
    cudaGetLastError();
    pg_set_error("large world: cannot grow the visit hash set to %llu slots", new_slots);
    return PG_ERR_CAPACITY;
  }
  PG_CUDA(cudaMemsetAsync(bigger, 0xff, sizeof(unsigned long long) * new_slots, w->stream));
  v.vset = bigger;
  v.vset_mask = new_slots - 1ull;
  k_vset_rehash<<<w->sm_count * 8, 256, 0, w->stream>>>(v, old, old_slots);
  w->launches++;
  PG_CUDA(cudaStreamSynchronize(w->stream));
  cudaFree(old);
  return PG_OK;
}

// Room for more state balls (contents kept).
static int lw_grow_balls(PgLarge *w, unsigned used) {
  LargeView &v = w->v;
  if (v.balls_cap > (1 << 29)) { pg_set_error("large world: state-ball buffer at its limit (%d)", v.balls_cap); return PG_ERR_CAPACITY; }
  const int cap = std::max(v.balls_cap * 2, (int)std::min<unsigned>(used + used / 2, 1u << 30));
  BallRec *bigger = nullptr;
  if (cudaMalloc(&bigger, sizeof(BallRec) * (size_t)cap) != cudaSuccess) {
    cudaGetLastError();
    pg_set_error("large world: cannot grow the state-ball buffer to %d", cap);
    return PG_ERR_CAPACITY;
  }
  const size_t keep = std::min<size_t>(used, (size_t)v.balls_cap);
  PG_CUDA(cudaMemcpyAsync(bigger, v.balls, sizeof(BallRec) * keep, cudaMemcpyDeviceToDevice, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  cudaFree(v.balls);
  v.balls = bigger; v.balls_cap = cap;
  return PG_OK;
}

static int lw_step_body(PgLarge *w, float dt_in);

// One step, all or nothing: pass 3 and the final kernel write the state in place, so the state the step began with is
// kept aside (two device copies, 32 B per body) and put back -- events and the touched-body bookkeeping with it -- if
// the step ends with an error (a body with more contacts than a timeline holds, a scratch buffer that cannot grow,
// non-finite state).  A step that is merely not PROVEN (PG_ERR_UNPROVEN) has run to its end and stays.
static int lw_step_once(PgLarge *w, float dt_in) {
  LargeView &v = w->v;
  const size_t bytes = sizeof(float4) * (size_t)v.n;
  PG_CUDA(cudaMemcpyAsync(v.snap_pos, v.pos_rest, bytes, cudaMemcpyDeviceToDevice, w->stream));
  PG_CUDA(cudaMemcpyAsync(v.snap_vel, v.vel_rad, bytes, cudaMemcpyDeviceToDevice, w->stream));
  PG_CUDA(cudaMemcpyAsync(v.snap_n_events, v.n_events, sizeof(int), cudaMemcpyDeviceToDevice, w->stream));
  const int rc = lw_step_body(w, dt_in);
  if (rc == PG_OK) return PG_OK;
  // roll back (best effort: if the device itself is gone there is nothing to restore)
  cudaStreamSynchronize(w->stream);
  cudaMemcpyAsync(v.pos_rest, v.snap_pos, bytes, cudaMemcpyDeviceToDevice, w->stream);
  cudaMemcpyAsync(v.vel_rad, v.snap_vel, bytes, cudaMemcpyDeviceToDevice, w->stream);
  cudaMemcpyAsync(v.n_events, v.snap_n_events, sizeof(int), cudaMemcpyDeviceToDevice, w->stream);
  cudaStreamSynchronize(w->stream);
  w->bookkeeping_clean = 0;                     // the next step resets the fixed point's per-body state for every body
  w->last_proven = 0;
  return rc;
}

static int lw_step_body(PgLarge *w, float dt_in) {
  LargeView &v = w->v;
  float dt = dt_in;
  if ((double)dt > 0.01666) dt = (float)0.01666;
  pg_leaf_lengths(0.0f, dt, &v.leaf_lo, &v.leaf_hi);
  const unsigned bbox_init[7] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u};
  PG_CUDA(cudaMemcpyAsync(v.bbox, bbox_init, sizeof(bbox_init), cudaMemcpyHostToDevice, w->stream));
  PG_CUDA(cudaMemsetAsync(v.flags, 0, sizeof(Flags), w->stream));
  PG_CUDA(cudaMemsetAsync(v.rstat, 0, sizeof(float) * 8, w->stream));
  PG_CUDA(cudaMemsetAsync(v.touched_bits, 0, sizeof(unsigned) * ((size_t)v.n / 32 + 1), w->stream));
  PG_CUDA(cudaMemsetAsync(v.head2, 0xff, sizeof(unsigned) * ((size_t)v.cells_cap + 1), w->stream));
  PG_CUDA(cudaMemsetAsync(v.head3, 0xff, sizeof(unsigned) * ((size_t)v.cells_cap + 1), w->stream));
  PG_CUDA(cudaMemsetAsync(v.vset, 0xff, sizeof(unsigned long long) * (v.vset_mask + 1ull), w->stream));
  const int full_reset = w->bookkeeping_clean ? 0 : 1;
  if (full_reset) {
    const size_t words = (size_t)w->cap_n / 32 + 1;
    PG_CUDA(cudaMemsetAsync(v.dirty_bits[0], 0, sizeof(unsigned) * words, w->stream));
    PG_CUDA(cudaMemsetAsync(v.dirty_bits[1], 0, sizeof(unsigned) * words, w->stream));
    PG_CUDA(cudaMemsetAsync(v.pushed_bits, 0, sizeof(unsigned) * words, w->stream));
  }
  w->bookkeeping_clean = 0;                     // until this step's k_reset_touched is in the stream
  if (w->timing) cudaEventRecord(w->kev[K_PASS3][0], w->stream);
  if (v.n >= v.pass3_split_min) {
    k_pass3<<<lw_grid(w, v.n, 256), 256, 0, w->stream>>>(v, dt, w->step_in_call, 1, full_reset);
    k_pass3_exact<<<w->sm_count * 8, 128, 0, w->stream>>>(v, dt, w->step_in_call, 1, full_reset);
    w->launches += 2;
  } else {
    k_pass3_all<<<lw_grid(w, v.n, 256), 256, 0, w->stream>>>(v, dt, w->step_in_call, 1, full_reset);
    w->launches++;
  }
  if (w->timing) cudaEventRecord(w->kev[K_PASS3][1], w->stream);
  w->last_proven = 1;
  w->last_iterations = 0;
  int rc = lw_build_pairs(w, dt);
  if (rc) return rc;
  Flags f;
  rc = lw_read_flags(w, &f);            // pair count: sizes the eval grid, detects overflow
  if (rc) return rc;
  lw_accumulate(w, K_PASS3); lw_accumulate(w, K_CELLS); lw_accumulate(w, K_SCAN); lw_accumulate(w, K_SCATTER); lw_accumulate(w, K_PAIRS);
  if (f.overflow_pairs) {
    // k_pairs is a pure function of the grid: grow the buffer and run it again
    rc = lw_grow_pairs(w, f.n_pairs, 0);
    if (rc) return rc;
    LW_LAUNCH(w, K_PAIRS, k_pairs, lw_grid(w, 2 * v.n, kPairsBlock), kPairsBlock, v, dt);
    rc = lw_read_flags(w, &f);
    if (rc) return rc;
    lw_accumulate(w, K_PAIRS);
    if (f.overflow_pairs) { pg_set_error("large world: candidate pair buffer too small (%lld)", v.pairs_cap); return PG_ERR_CAPACITY; }
  }
  if (f.unsafe) { pg_set_error("large world: non-finite body state"); return PG_ERR_ARG; }
  w->last_candidates = (long long)f.n_pairs;
  const bool debug = getenv("PG_LARGE_DEBUG") != nullptr && w->timing;
  bool converged = false;
  // Iterations are queued in batches between host read-backs (k_iter_gate stops a batch on the device when the
  // fixed point has converged or the host has to act); with per-kernel timing on, one at a time.
  const int wide_grid = w->sm_count * 8;
  const int batch = w->timing ? 1 : w->iter_batch;
  int queued = 0;                       // iterations queued so far this step (== the iteration index of the next one)
  while (queued < kMaxIter) {
    const unsigned long long pairs_before = std::min<unsigned long long>(f.n_pairs, (unsigned long long)v.pairs_cap);
    const int first_of_batch = queued;
    for (int q = 0; q < batch && queued < kMaxIter; q++, queued++) {
      const int it = queued, src = it & 1, dst = src ^ 1;
      k_iter_gate<<<1, 1, 0, w->stream>>>(v, it == 0 ? 1 : 0);
      if (w->timing) cudaEventRecord(w->kev[K_EVAL][0], w->stream);
      k_begin<<<wide_grid, 256, 0, w->stream>>>(v, src, dst, it, dt);      // carry + selection; also resets the per-iteration flags
      k_eval<<<w->sm_count * 4, 128, 0, w->stream>>>(v, src, dst, it, dt);
      k_eval_hard<<<w->sm_count, 512, 0, w->stream>>>(v, src, dst, it, dt);
      if (w->timing) cudaEventRecord(w->kev[K_EVAL][1], w->stream);
      LW_LAUNCH(w, K_CHECK, k_check, wide_grid, 256, v, src, dst, it, dt);
      // balls of the states that appeared in the NEW timelines, and the visits they add (marked fresh: the next
      // iteration evaluates them)
      if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][0], w->stream);
      k_states_insert<<<wide_grid, 128, 0, w->stream>>>(v, src, dst, dt);
      k_states_query<<<wide_grid, 256, 0, w->stream>>>(v, dt);
      if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][1], w->stream);
      w->launches += 6;
    }
    rc = lw_read_flags(w, &f);
    if (rc) return rc;
    lw_accumulate(w, K_EVAL); lw_accumulate(w, K_CHECK); lw_accumulate(w, K_BOUNDS);
    if (debug) {
      float e = 0, c = 0, rq = 0;
      cudaEventElapsedTime(&e, w->kev[K_EVAL][0], w->kev[K_EVAL][1]);
      cudaEventElapsedTime(&c, w->kev[K_CHECK][0], w->kev[K_CHECK][1]);
      cudaEventElapsedTime(&rq, w->kev[K_BOUNDS][0], w->kev[K_BOUNDS][1]);
      fprintf(stderr, "  [pg_large] iter %d: eval %.3f ms check %.3f ms states %.3f ms hits %llu changed %d visits %llu -> %llu touched %u balls %u (+%u) huge %u set %u\n",
              first_of_batch, e, c, rq, (unsigned long long)f.n_hits, f.changed, pairs_before, (unsigned long long)f.n_pairs, f.n_touched, f.n_balls,
              f.n_balls - f.balls_begin, f.n_huge, f.n_vset);
    }
    w->last_iterations = (int)f.iters_done;
    w->last_hits = (long long)f.n_hits;
    const bool was_stopped = f.stop != 0;
    if (f.overflow_timeline) { pg_set_error("large world: more than %d contacts on one body in one step", kTimeline); return PG_ERR_CAPACITY; }
    if (f.overflow_vset) { pg_set_error("large world: visit hash set too small (%llu slots)", (unsigned long long)v.vset_mask + 1ull); return PG_ERR_CAPACITY; }
    if (f.unsafe) { pg_set_error("large world: non-finite body state"); return PG_ERR_ARG; }
    if (f.overflow_huge) { pg_set_error("large world: more than %d balls beyond the tail bound of the grid", v.huge_cap); return PG_ERR_CAPACITY; }
    // Whatever stopped the batch, the timelines the LAST executed iteration wrote are intact; its state balls and
    // the visits they add may be incomplete if a buffer ran full.  Make room and repeat those two kernels: balls
    // that were stored already are linked a second time (harmless duplicates), visits that were stored already
    // are in the hash set (the query took the ones it could not store out again).
    int guard = 0;
    while (f.overflow_pairs || f.overflow_balls) {
      if (++guard > 8) { pg_set_error("large world: candidate visit / state-ball buffers too small (%lld, %d)", v.pairs_cap, v.balls_cap); return PG_ERR_CAPACITY; }
      const int last_it = (int)f.iters_done - 1, lsrc = last_it & 1, ldst = lsrc ^ 1;
      if (f.overflow_pairs) { rc = lw_grow_pairs(w, f.n_pairs, (unsigned long long)v.pairs_cap); if (rc) return rc; }
      const bool redo_insert = f.overflow_balls != 0;
      if (redo_insert) {
        const unsigned kept = std::min<unsigned>(f.n_balls, (unsigned)v.balls_cap);
        rc = lw_grow_balls(w, std::max<unsigned>(f.n_balls, kept));
        if (rc) return rc;
        PG_CUDA(cudaMemcpyAsync(&v.flags->n_balls, &kept, sizeof(kept), cudaMemcpyHostToDevice, w->stream));
        PG_CUDA(cudaMemsetAsync(&v.flags->overflow_balls, 0, sizeof(int), w->stream));
      }
      PG_CUDA(cudaMemsetAsync(&v.flags->stop, 0, sizeof(int), w->stream));
      if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][0], w->stream);
      if (redo_insert) { k_states_insert<<<wide_grid, 128, 0, w->stream>>>(v, lsrc, ldst, dt); w->launches++; }
      k_states_query<<<wide_grid, 256, 0, w->stream>>>(v, dt);
      if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][1], w->stream);
      w->launches++;
      rc = lw_read_flags(w, &f);
      if (rc) return rc;
      lw_accumulate(w, K_BOUNDS);
    }
    w->last_candidates = (long long)f.n_pairs;
    if (f.converged) { converged = true; break; }
    // keep the hash set and the ball buffer roomy (load <= 1/4, fill <= 1/2) for the iterations to come
    if ((unsigned long long)f.n_vset * 4ull > v.vset_mask + 1ull) { rc = lw_grow_vset(w); if (rc) return rc; }
    if ((long long)f.n_balls * 2 > (long long)v.balls_cap) { rc = lw_grow_balls(w, f.n_balls); if (rc) return rc; }
    if (was_stopped) {
      // the batch stopped early for the host's sake: the iterations that did not run are queued again
      PG_CUDA(cudaMemsetAsync(&v.flags->stop, 0, sizeof(int), w->stream));
      queued = (int)f.iters_done;
    }
  }
  const int src = (int)(f.iters_done & 1u);       // the timelines the last executed iteration wrote
  if (!converged) w->last_proven = 0;
  LW_LAUNCH(w, K_FINAL, k_finalize, lw_grid(w, v.n, 256), 256, v, src, dt, w->step_in_call);
  k_reset_touched<<<w->sm_count * 4, 256, 0, w->stream>>>(v);
  w->launches++;
  w->bookkeeping_clean = 1;
  PG_CUDA(cudaGetLastError());
  PG_CUDA(cudaStreamSynchronize(w->stream));
  lw_accumulate(w, K_FINAL);
  return PG_OK;
}

// World-space box of a static AABB body: node_decompose + AABB_update (distance.c:27-36, aabb.c:59-78)
// in host arithmetic -- IEEE float/double like the device, so the bits equal what the device
// functions of pg_math.cuh would produce for the same record.
static void host_world_box(const PgBody &b, float *c, float *e) {
  for (int k = 0; k < 3; k++) { c[k] = 0.0f; e[k] = 0.0f; }
  if (!b.has_node) return;
  const float *m = b.node_xform;
  float pos[3], scl[3], rot[9];
  for (int k = 0; k < 3; k++) pos[k] = m[0 + k] * 0.0f + m[4 + k] * 0.0f + m[8 + k] * 0.0f + m[12 + k] * 1.0f;
  for (int col = 0; col < 3; col++) {
    const float *q = m + col * 4;
    scl[col] = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  }
  for (int col = 0; col < 3; col++) for (int r = 0; r < 3; r++) rot[col * 3 + r] = m[col * 4 + r];
  if (scl[0] != 0.0f) { const float inv = 1.0f / scl[0]; for (int k = 0; k < 9; k++) rot[k] *= inv; }
  for (int k = 0; k < 3; k++) pos[k] += b.velocity[k] * 0.0f;            // muladds(velocity, time, pos) with a static body
  for (int i = 0; i < 3; i++) {
    float cc = pos[i], ee = 0.0f;
    for (int j = 0; j < 3; j++) {
      cc += rot[j * 3 + i] * b.shape[j];
      ee = (float)((double)ee + fabs((double)rot[j * 3 + i]) * (double)b.shape[3 + j] * fabs((double)scl[i]));
    }
    c[i] = cc; e[i] = ee;
  }
}

extern "C" {

PgLarge *pg_large_create(int max_spheres, int max_static, int device) {
  if (max_spheres <= 0 || max_static < 0 || max_static > (1 << 22)) { pg_set_error("pg_large_create: bad argument"); return nullptr; }
  PG_CUDA_NULL(cudaSetDevice(device));
  PgLarge *w = new PgLarge();
  memset(w, 0, sizeof(*w));
  w->timing = getenv("PG_LARGE_TIMING") != nullptr || getenv("PG_LARGE_DEBUG") != nullptr;
  w->iter_batch = getenv("PG_LARGE_ITER_BATCH") ? std::max(1, atoi(getenv("PG_LARGE_ITER_BATCH"))) : 4;
  w->device = device;
  w->cap_n = max_spheres;
  w->cap_static = std::max(max_static, 1);
  cudaDeviceProp prop;
  PG_CUDA_NULL(cudaGetDeviceProperties(&prop, device));
  w->sm_count = prop.multiProcessorCount;
  PG_CUDA_NULL(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
  PG_CUDA_NULL(cudaEventCreate(&w->ev0));
  PG_CUDA_NULL(cudaEventCreate(&w->ev1));
  for (int k = 0; k < K_COUNT; k++) { PG_CUDA_NULL(cudaEventCreate(&w->kev[k][0])); PG_CUDA_NULL(cudaEventCreate(&w->kev[k][1])); }
  LargeView &v = w->v;
  const size_t n = (size_t)max_spheres;
  // test hooks (read once per handle): tiny shared-memory stages reach their overflow paths, and a zero
  // threshold keeps k_requery on its eight-lane path in worlds where few bodies grow
  v.stage_limit = getenv("PG_LARGE_SMALL_STAGES") ? 3 : (1 << 20);
  v.pass3_split_min = getenv("PG_LARGE_PASS3_SPLIT_MIN") ? atoi(getenv("PG_LARGE_PASS3_SPLIT_MIN")) : 262144;
  v.ecap_factor = getenv("PG_LARGE_ECAP_FACTOR") ? (float)atof(getenv("PG_LARGE_ECAP_FACTOR")) : 1.5f;
  v.ecap2_factor = getenv("PG_LARGE_ECAP2_FACTOR") ? (float)atof(getenv("PG_LARGE_ECAP2_FACTOR")) : 4.0f;
  v.r_uniform = -1.0f;
  v.cells_cap = (int)std::min<size_t>(4 * n + 4096, (size_t)1 << 28);
  v.pairs_cap = (long long)(8 * n + 65536);
  v.max_events = (int)std::max<size_t>(n / 4, 65536);
  PG_CUDA_NULL(cudaMalloc(&v.pos_rest, sizeof(float4) * n));
  PG_CUDA_NULL(cudaMalloc(&v.vel_rad, sizeof(float4) * n));
  PG_CUDA_NULL(cudaMalloc(&v.snap_pos, sizeof(float4) * n));
  PG_CUDA_NULL(cudaMalloc(&v.snap_vel, sizeof(float4) * n));
  PG_CUDA_NULL(cudaMalloc(&v.snap_n_events, sizeof(int)));
  PG_CUDA_NULL(cudaMalloc(&v.planes, sizeof(float4) * kMaxPlanesL));
  PG_CUDA_NULL(cudaMalloc(&v.plane_sidx, sizeof(int) * kMaxPlanesL));
  w->cap_static_total = std::max(max_static, kMaxPlanesL);
  PG_CUDA_NULL(cudaMalloc(&v.box_c, sizeof(float4) * (size_t)w->cap_static_total));
  PG_CUDA_NULL(cudaMalloc(&v.box_e, sizeof(float4) * (size_t)w->cap_static_total));
  PG_CUDA_NULL(cudaMalloc(&v.static_type, (size_t)w->cap_static_total));
  PG_CUDA_NULL(cudaMalloc(&v.env, sizeof(float) * 2 * n));
  PG_CUDA_NULL(cudaMalloc(&v.bbox, sizeof(unsigned) * 8));
  PG_CUDA_NULL(cudaMalloc(&v.grid, sizeof(GridParams)));
  PG_CUDA_NULL(cudaMalloc(&v.rstat, sizeof(float) * 8));
  PG_CUDA_NULL(cudaMalloc(&v.robust, sizeof(Robust)));
  {
    Robust r0;
    for (int k = 0; k < 3; k++) { r0.centre[k] = 0.0f; r0.half[k] = 3.0e38f; }
    r0.valid = 0;
    PG_CUDA_NULL(cudaMemcpy(v.robust, &r0, sizeof(r0), cudaMemcpyHostToDevice));
  }
  v.balls_cap = (int)std::min<size_t>(4 * n + 65536, (size_t)1 << 30);
  const bool tiny = getenv("PG_LARGE_TINY_BUFFERS") != nullptr;      // test hook: every growth / repeat path of the fixed point runs
  if (tiny) { v.balls_cap = 256; v.pairs_cap = 4096; }
  PG_CUDA_NULL(cudaMalloc(&v.balls, sizeof(BallRec) * (size_t)v.balls_cap));
  PG_CUDA_NULL(cudaMalloc(&v.head2, sizeof(unsigned) * ((size_t)v.cells_cap + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.head3, sizeof(unsigned) * ((size_t)v.cells_cap + 1)));
  v.huge_cap = 65536;
  PG_CUDA_NULL(cudaMalloc(&v.huge_list, sizeof(int) * (size_t)v.huge_cap));
  {
    unsigned long long slots = 1ull << 16;
    while (slots < 8ull * n) slots <<= 1;
    if (tiny) slots = 1ull << 12;
    v.vset_mask = slots - 1ull;
    PG_CUDA_NULL(cudaMalloc(&v.vset, sizeof(unsigned long long) * slots));
  }
  PG_CUDA_NULL(cudaMalloc(&v.cell_of, sizeof(unsigned) * 2 * n));
  PG_CUDA_NULL(cudaMalloc(&v.cell_cnt, sizeof(unsigned) * ((size_t)v.cells_cap + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.block_sums, sizeof(unsigned) * ((size_t)v.cells_cap / (kScanBlock * kScanItems) + 2)));
  PG_CUDA_NULL(cudaMalloc(&v.sorted, sizeof(SortedBody) * 2 * n));
  PG_CUDA_NULL(cudaMalloc(&v.pairs, sizeof(int2) * (size_t)v.pairs_cap));
  PG_CUDA_NULL(cudaMalloc(&v.active, sizeof(unsigned) * 2 * (size_t)v.pairs_cap));
  PG_CUDA_NULL(cudaMalloc(&v.seen, sizeof(int) * n));
  PG_CUDA_NULL(cudaMalloc(&v.touched_bits, sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.dirty_bits[0], sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.dirty_bits[1], sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.pushed_bits, sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.touched, sizeof(int) * n));
  PG_CUDA_NULL(cudaMalloc(&v.cell_rank, sizeof(unsigned) * 2 * n));
  for (int k = 0; k < 2; k++) {
    PG_CUDA_NULL(cudaMalloc(&v.tl_chg[k], sizeof(unsigned long long) * n));
    PG_CUDA_NULL(cudaMalloc(&v.tl_cnt[k], n + 8));
    PG_CUDA_NULL(cudaMemset(v.tl_cnt[k], 0, n + 8));
    PG_CUDA_NULL(cudaMalloc(&v.tl_key[k], sizeof(unsigned long long) * n * kTimeline));
    PG_CUDA_NULL(cudaMalloc(&v.tl_pos[k], sizeof(float4) * n * kTimeline));
    PG_CUDA_NULL(cudaMalloc(&v.tl_vel[k], sizeof(float4) * n * kTimeline));
  }
  PG_CUDA_NULL(cudaMalloc(&v.flags, sizeof(Flags)));
  PG_CUDA_NULL(cudaMalloc(&v.events, sizeof(PgEvent) * (size_t)v.max_events));
  PG_CUDA_NULL(cudaMalloc(&v.n_events, sizeof(int)));
  PG_CUDA_NULL(cudaMemset(v.n_events, 0, sizeof(int)));
  PG_CUDA_NULL(cudaMalloc(&w->d_stats, 8 * sizeof(double)));
  PG_CUDA_NULL(cudaMalloc(&w->d_stage_pos, sizeof(float) * 3 * n));
  PG_CUDA_NULL(cudaMalloc(&w->d_stage_vel, sizeof(float) * 3 * n));
  PG_CUDA_NULL(cudaMalloc(&w->d_stage_rest, n));
  return w;
}

void pg_large_destroy(PgLarge *w) {
  if (!w) return;
  cudaSetDevice(w->device);
  cudaStreamSynchronize(w->stream);
  LargeView &v = w->v;
  if (w->player_world) pg_worlds_destroy(w->player_world);
  void *ptrs[] = {v.plane_sidx, v.box_c, v.box_e, v.static_type, v.sg_start, v.sg_items, v.pos_rest, v.vel_rad, v.snap_pos, v.snap_vel, v.snap_n_events, v.planes, v.env, v.bbox, v.grid, v.cell_of, v.cell_cnt, v.rstat, v.robust, v.balls, v.head2, v.head3, v.huge_list, v.vset, v.touched_bits, v.dirty_bits[0], v.dirty_bits[1], v.pushed_bits, v.block_sums, v.sorted,
                  v.pairs, v.active, v.tl_chg[0], v.tl_chg[1], v.seen, v.touched, v.cell_rank, v.tl_cnt[0], v.tl_cnt[1], v.tl_key[0], v.tl_key[1], v.tl_pos[0], v.tl_pos[1], v.tl_vel[0], v.tl_vel[1],
                  v.flags, v.events, v.n_events, w->d_stats, w->d_stage_pos, w->d_stage_vel, w->d_stage_rest};
  for (void *p : ptrs) if (p) cudaFree(p);
  cudaEventDestroy(w->ev0); cudaEventDestroy(w->ev1);
  for (int k = 0; k < K_COUNT; k++) { cudaEventDestroy(w->kev[k][0]); cudaEventDestroy(w->kev[k][1]); }
  cudaStreamDestroy(w->stream);
  delete w;
}

}  // extern "C"

namespace {
__global__ void k_lw_pack(float4 *pos_rest, float4 *vel_rad, int n, const float *pos, const float *vel, const unsigned char *rest,
                          const float *radius, float r_uniform, int set_radius) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    pos_rest[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], (rest && rest[i]) ? 1.0f : 0.0f);
    const float r = set_radius ? (radius ? radius[i] : r_uniform) : vel_rad[i].w;
    vel_rad[i] = make_float4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], r);
  }
}
__global__ void k_lw_unpack(const float4 *pos_rest, const float4 *vel_rad, int n, float *pos, float *vel, unsigned char *rest) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 a = pos_rest[i], b = vel_rad[i];
    pos[3 * i] = a.x; pos[3 * i + 1] = a.y; pos[3 * i + 2] = a.z;
    vel[3 * i] = b.x; vel[3 * i + 1] = b.y; vel[3 * i + 2] = b.z;
    if (rest) rest[i] = a.w != 0.0f ? 1 : 0;
  }
}
__global__ void k_lw_stats(LargeView v, double *out, double *out2) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
    const float4 a = v.pos_rest[i], b = v.vel_rad[i];
    acc[0] += 1.0;
    acc[1] += 0.5 * ((double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z);
    acc[2] += b.x; acc[3] += b.y; acc[4] += b.z;
    acc[6] += a.w != 0.0f ? 1.0 : 0.0;
    const bool fin = isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(b.x) && isfinite(b.y) && isfinite(b.z);
    acc[7] += fin ? 0.0 : 1.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) acc[5] += (double)*v.n_events;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    double x = acc[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFullMask, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) { atomicAdd(out + q, x); if (out2) atomicAdd(out2 + q, x); }
  }
}
}  // namespace

extern "C" {

int pg_large_set(PgLarge *w, const PgBody *statics, int n_static, int n_spheres, const float *pos, const float *vel,
                 const float *radius_or_null, float radius, float restitution) {
  if (!w || n_spheres < 0 || n_spheres > w->cap_n || n_static < 0 || n_static > w->cap_static_total || !pos || !vel) { pg_set_error("pg_large_set: bad argument"); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  w->bookkeeping_clean = 0;                     // the body count may grow: the next step resets every body
  LargeView &v = w->v;
  // statics: planes (visited against body->position) and AABBs (visited through scene nodes), array order kept
  std::vector<float4> pl, bc((size_t)std::max(n_static, 1)), be((size_t)std::max(n_static, 1));
  std::vector<int> pidx;
  std::vector<unsigned char> types((size_t)std::max(n_static, 1));
  int n_boxes = 0;
  float blo[3] = {3e38f, 3e38f, 3e38f}, bhi[3] = {-3e38f, -3e38f, -3e38f};
  for (int i = 0; i < n_static; i++) {
    types[i] = (unsigned char)statics[i].collider_type;
    bc[i] = make_float4(0, 0, 0, 0); be[i] = make_float4(0, 0, 0, 0);
    if (statics[i].collider_type == PG_PLANE) {
      if ((int)pl.size() >= kMaxPlanesL) { pg_set_error("pg_large_set: at most %d static planes", kMaxPlanesL); return PG_ERR_CAPACITY; }
      pl.push_back(make_float4(statics[i].shape[0], statics[i].shape[1], statics[i].shape[2], statics[i].shape[3]));
      pidx.push_back(i);
    } else if (statics[i].collider_type == PG_AABB) {
      if (!statics[i].has_node) { pg_set_error("pg_large_set: static AABB %d has no scene node (the reference reads boxes through scene_node->world_transform)", i); return PG_ERR_UNSUPPORTED; }
      float c[3], e[3];
      host_world_box(statics[i], c, e);
      bc[i] = make_float4(c[0], c[1], c[2], 0.0f); be[i] = make_float4(e[0], e[1], e[2], 0.0f);
      for (int k = 0; k < 3; k++) { blo[k] = std::min(blo[k], c[k] - e[k]); bhi[k] = std::max(bhi[k], c[k] + e[k]); }
      n_boxes++;
    } else {
      pg_set_error("pg_large_set: static bodies must be planes or AABBs");
      return PG_ERR_UNSUPPORTED;
    }
  }
  v.n = n_spheres; v.ns = (int)pl.size(); v.restitution = restitution;
  v.r_uniform = radius_or_null ? -1.0f : radius * 1.0f;
  v.n_static_total = n_static; v.n_boxes = n_boxes;
  if (!pl.empty()) {
    PG_CUDA(cudaMemcpy(v.planes, pl.data(), sizeof(float4) * pl.size(), cudaMemcpyHostToDevice));
    PG_CUDA(cudaMemcpy(v.plane_sidx, pidx.data(), sizeof(int) * pidx.size(), cudaMemcpyHostToDevice));
  }
  if (n_static) {
    PG_CUDA(cudaMemcpy(v.box_c, bc.data(), sizeof(float4) * (size_t)n_static, cudaMemcpyHostToDevice));
    PG_CUDA(cudaMemcpy(v.box_e, be.data(), sizeof(float4) * (size_t)n_static, cudaMemcpyHostToDevice));
    PG_CUDA(cudaMemcpy(v.static_type, types.data(), (size_t)n_static, cudaMemcpyHostToDevice));
  }
  if (v.sg_start) { cudaFree(v.sg_start); v.sg_start = nullptr; }
  if (v.sg_items) { cudaFree(v.sg_items); v.sg_items = nullptr; }
  if (n_boxes) {
    // level grid: every box is registered in the cells its bounds, inflated by sg_reach, overlap.  A
    // sphere whose reach r + |v| dt is below sg_reach only needs the list of the cell its centre is in.
    const float reach = 1.0f;
    double vol = 1.0;
    for (int k = 0; k < 3; k++) vol *= std::max(1.0, (double)bhi[k] - (double)blo[k] + 2.0 * reach);
    float h = std::max(2.0f, (float)cbrt(vol / (8.0 * (double)n_boxes)));
    int dim[3];
    for (;;) {
      double cells = 1.0;
      for (int k = 0; k < 3; k++) { dim[k] = (int)floor(((double)bhi[k] - (double)blo[k] + 2.0 * reach) / h) + 1; cells *= dim[k]; }
      if (cells <= 4.0e6) break;
      h *= 1.26f;
    }
    v.sg_reach = reach; v.sg_inv_h = 1.0f / h;
    for (int k = 0; k < 3; k++) { v.sg_org[k] = blo[k] - reach; v.sg_dim[k] = dim[k]; }
    const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    std::vector<unsigned> start(ncell + 1, 0u);
    auto cell_range = [&](int i, int *lo3, int *hi3) {
      const float c[3] = {bc[i].x, bc[i].y, bc[i].z}, e[3] = {be[i].x, be[i].y, be[i].z};
      for (int k = 0; k < 3; k++) {
        // one extra cell of slack on each side: the device computes floor((p - org) * inv_h) in float
        lo3[k] = std::max(0, (int)floorf((c[k] - e[k] - reach - v.sg_org[k]) * v.sg_inv_h) - 1);
        hi3[k] = std::min(dim[k] - 1, (int)floorf((c[k] + e[k] + reach - v.sg_org[k]) * v.sg_inv_h) + 1);
      }
    };
    for (int i = 0; i < n_static; i++) {
      if (types[i] != PG_AABB) continue;
      int lo3[3], hi3[3]; cell_range(i, lo3, hi3);
      for (int z = lo3[2]; z <= hi3[2]; z++) for (int y = lo3[1]; y <= hi3[1]; y++) for (int x = lo3[0]; x <= hi3[0]; x++)
        start[((size_t)z * dim[1] + y) * dim[0] + x + 1]++;
    }
    for (size_t c = 0; c < ncell; c++) start[c + 1] += start[c];
    std::vector<int> items(start[ncell]);
    std::vector<unsigned> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < n_static; i++) {          // ascending i: every cell list ends up sorted by static index
      if (types[i] != PG_AABB) continue;
      int lo3[3], hi3[3]; cell_range(i, lo3, hi3);
      for (int z = lo3[2]; z <= hi3[2]; z++) for (int y = lo3[1]; y <= hi3[1]; y++) for (int x = lo3[0]; x <= hi3[0]; x++)
        items[fill[((size_t)z * dim[1] + y) * dim[0] + x]++] = i;
    }
    PG_CUDA(cudaMalloc(&v.sg_start, sizeof(unsigned) * (ncell + 1)));
    PG_CUDA(cudaMalloc(&v.sg_items, sizeof(int) * std::max<size_t>(items.size(), 1)));
    PG_CUDA(cudaMemcpy(v.sg_start, start.data(), sizeof(unsigned) * (ncell + 1), cudaMemcpyHostToDevice));
    if (!items.empty()) PG_CUDA(cudaMemcpy(v.sg_items, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice));
  }
  // players x level geometry run through the any-collider kernel on a private one-world handle
  if (w->player_world) { pg_worlds_destroy(w->player_world); w->player_world = nullptr; }
  w->n_players = 0;
  if (n_static) {
    w->player_world = pg_worlds_create(1, n_static, 1, 8, 65536, w->device);
    if (!w->player_world) return PG_ERR_CUDA;
    int rc = pg_worlds_set_bodies(w->player_world, 0, PG_CLASS_STATIC, statics, n_static);
    if (rc) return rc;
  }
  const size_t n = (size_t)n_spheres;
  PG_CUDA(cudaMemcpyAsync(w->d_stage_pos, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, w->stream));
  PG_CUDA(cudaMemcpyAsync(w->d_stage_vel, vel, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, w->stream));
  float *d_rad = nullptr;
  if (radius_or_null) {
    PG_CUDA(cudaMalloc(&d_rad, sizeof(float) * n));
    PG_CUDA(cudaMemcpyAsync(d_rad, radius_or_null, sizeof(float) * n, cudaMemcpyHostToDevice, w->stream));
  }
  k_lw_pack<<<lw_grid(w, n_spheres, 256), 256, 0, w->stream>>>(v.pos_rest, v.vel_rad, n_spheres, w->d_stage_pos, w->d_stage_vel, nullptr, d_rad, radius * 1.0f, 1);
  w->launches++;
  PG_CUDA(cudaStreamSynchronize(w->stream));
  if (d_rad) cudaFree(d_rad);
  return PG_OK;
}

int pg_large_upload_state(PgLarge *w, const float *pos, const float *vel, const uint8_t *at_rest) {
  if (!w || !pos || !vel) return PG_ERR_ARG;
  PG_CUDA(cudaSetDevice(w->device));
  const size_t n = (size_t)w->v.n;
  PG_CUDA(cudaMemcpyAsync(w->d_stage_pos, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, w->stream));
  PG_CUDA(cudaMemcpyAsync(w->d_stage_vel, vel, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, w->stream));
  if (at_rest) PG_CUDA(cudaMemcpyAsync(w->d_stage_rest, at_rest, n, cudaMemcpyHostToDevice, w->stream));
  k_lw_pack<<<lw_grid(w, w->v.n, 256), 256, 0, w->stream>>>(w->v.pos_rest, w->v.vel_rad, w->v.n, w->d_stage_pos, w->d_stage_vel,
                                                              at_rest ? w->d_stage_rest : nullptr, nullptr, 0.0f, 0);
  w->launches++;
  PG_CUDA(cudaGetLastError());
  return PG_OK;
}

int pg_large_download_state(PgLarge *w, float *pos, float *vel, uint8_t *at_rest) {
  if (!w || !pos || !vel) return PG_ERR_ARG;
  PG_CUDA(cudaSetDevice(w->device));
  const size_t n = (size_t)w->v.n;
  k_lw_unpack<<<lw_grid(w, w->v.n, 256), 256, 0, w->stream>>>(w->v.pos_rest, w->v.vel_rad, w->v.n, w->d_stage_pos, w->d_stage_vel,
                                                                at_rest ? w->d_stage_rest : nullptr);
  w->launches++;
  PG_CUDA(cudaMemcpyAsync(pos, w->d_stage_pos, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, w->stream));
  PG_CUDA(cudaMemcpyAsync(vel, w->d_stage_vel, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, w->stream));
  if (at_rest) PG_CUDA(cudaMemcpyAsync(at_rest, w->d_stage_rest, n, cudaMemcpyDeviceToHost, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  return PG_OK;
}

int pg_large_step(PgLarge *w, float dt, int n_steps) {
  if (!w || n_steps < 0) return PG_ERR_ARG;
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaMemsetAsync(w->v.n_events, 0, sizeof(int), w->stream));
  for (int k = 0; k < K_COUNT; k++) w->kms[k] = 0.0f;
  PG_CUDA(cudaEventRecord(w->ev0, w->stream));
  int rc = PG_OK;
  // passes 1-2 (players x statics, players x dynamics): independent of the spheres, because the
  // sphere-capsule table entries never fire (undefined upstream) -- all steps at once
  if (w->player_world && w->n_players > 0 && n_steps > 0) {
    rc = pg_worlds_step(w->player_world, dt, n_steps, 1);
    if (rc) return rc;
    w->launches++;
  }
  int unproven = 0, first_unproven = -1;
  for (int s = 0; s < n_steps && rc == PG_OK; s++) {
    w->step_in_call = s;
    rc = lw_step_once(w, dt);
    if (rc == PG_OK && !w->last_proven) { if (!unproven) first_unproven = s; unproven++; }
  }
  PG_CUDA(cudaEventRecord(w->ev1, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  if (rc == PG_OK && unproven) {
    // loud, not silent: the state HAS been advanced (all n_steps ran), but at least one step's fixed point did
    // not converge or an envelope outgrew the grid, so equality with the sequential sweep is not established
    w->last_proven = 0;
    pg_set_error("pg_large_step: %d of %d steps not proven identical to the sequential reference order (first: step %d)", unproven, n_steps, first_unproven);
    return PG_ERR_UNPROVEN;
  }
  return rc;
}

int pg_large_sync(PgLarge *w) {
  if (!w) return PG_ERR_ARG;
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  return PG_OK;
}

long long pg_large_pairs(PgLarge *w, float dt_in, int32_t *out_ij, long long cap) {
  if (!w) return PG_ERR_ARG;
  if (cudaSetDevice(w->device) != cudaSuccess) return PG_ERR_CUDA;
  LargeView &v = w->v;
  float dt = dt_in;
  if ((double)dt > 0.01666) dt = (float)0.01666;
  pg_leaf_lengths(0.0f, dt, &v.leaf_lo, &v.leaf_hi);
  const unsigned bbox_init[7] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u};
  PG_CUDA(cudaMemcpyAsync(v.bbox, bbox_init, sizeof(bbox_init), cudaMemcpyHostToDevice, w->stream));
  PG_CUDA(cudaMemsetAsync(v.flags, 0, sizeof(Flags), w->stream));
  PG_CUDA(cudaMemsetAsync(v.rstat, 0, sizeof(float) * 8, w->stream));       // no running statistics here: plain bounding box
  PG_CUDA(cudaMemsetAsync(v.head2, 0xff, sizeof(unsigned) * ((size_t)v.cells_cap + 1), w->stream));
  PG_CUDA(cudaMemsetAsync(v.head3, 0xff, sizeof(unsigned) * ((size_t)v.cells_cap + 1), w->stream));
  k_env_frozen<<<lw_grid(w, v.n, 256), 256, 0, w->stream>>>(v, dt);
  w->launches++;
  int rc = lw_build_pairs(w, dt);
  if (rc) return rc;
  Flags f;
  rc = lw_read_flags(w, &f);
  if (rc) return rc;
  if (f.unsafe) { pg_set_error("large world: non-finite state, grid cannot be built"); return PG_ERR_ARG; }
  if (f.overflow_pairs) { pg_set_error("pg_large_pairs: candidate buffer too small"); return PG_ERR_CAPACITY; }
  const long long np = (long long)f.n_pairs;
  if (np == 0) return 0;
  unsigned char *d_keep = nullptr;
  PG_CUDA(cudaMalloc(&d_keep, (size_t)np));
  k_pairs_exact<<<lw_grid(w, (int)np, 128), 128, 0, w->stream>>>(v, dt, d_keep);
  w->launches++;
  std::vector<unsigned char> keep((size_t)np);
  std::vector<int2> pairs((size_t)np);
  PG_CUDA(cudaMemcpyAsync(keep.data(), d_keep, (size_t)np, cudaMemcpyDeviceToHost, w->stream));
  PG_CUDA(cudaMemcpyAsync(pairs.data(), v.pairs, sizeof(int2) * (size_t)np, cudaMemcpyDeviceToHost, w->stream));
  PG_CUDA(cudaStreamSynchronize(w->stream));
  cudaFree(d_keep);
  std::vector<std::pair<int, int>> out;
  for (long long k = 0; k < np; k++) if (keep[k]) out.emplace_back(pairs[k].x, pairs[k].y);
  std::sort(out.begin(), out.end());
  for (size_t k = 0; k < out.size() && (long long)k < cap; k++) { out_ij[2 * k] = out[k].first; out_ij[2 * k + 1] = out[k].second; }
  return (long long)out.size();
}

long long pg_large_events(PgLarge *w, PgEvent *out, long long cap, int *overflowed) {
  if (!w) return PG_ERR_ARG;
  if (cudaSetDevice(w->device) != cudaSuccess || cudaStreamSynchronize(w->stream) != cudaSuccess) return PG_ERR_CUDA;
  int n = 0;
  if (cudaMemcpy(&n, w->v.n_events, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return PG_ERR_CUDA;
  int over = 0;
  if (n > w->v.max_events) { over = 1; n = w->v.max_events; }
  std::vector<PgEvent> buf((size_t)n);
  if (n && cudaMemcpy(buf.data(), w->v.events, sizeof(PgEvent) * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return PG_ERR_CUDA;
  for (auto &e : buf) e.pad_ >>= 8;                       // device tag -> pass number
  if (w->player_world && w->n_players > 0) {
    int pover = 0;
    long long np = pg_worlds_events(w->player_world, nullptr, 0, &pover);
    if (np > 0) {
      std::vector<PgEvent> pe((size_t)np);
      np = pg_worlds_events(w->player_world, pe.data(), np, &pover);
      buf.insert(buf.end(), pe.begin(), pe.begin() + np);   // pad_ already holds the pass (1 or 2)
    }
    over |= pover;
  }
  if (overflowed) *overflowed = over;
  // solve order: step, pass, row body, column body.  Rows: player (1-2), dynamic sphere (3-4).
  auto row_col = [](const PgEvent &e, int *row, int *col) {
    const int row_class = e.pad_ <= 2 ? PG_CLASS_PLAYER : PG_CLASS_DYNAMIC;
    if (e.pad_ == 4) { *row = e.a_index; *col = e.b_index; return; }
    if (e.a_class == row_class) { *row = e.a_index; *col = e.b_index; } else { *row = e.b_index; *col = e.a_index; }
  };
  std::stable_sort(buf.begin(), buf.end(), [&](const PgEvent &a, const PgEvent &b) {
    if (a.step != b.step) return a.step < b.step;
    if (a.pad_ != b.pad_) return a.pad_ < b.pad_;
    int ra, ca, rb, cb;
    row_col(a, &ra, &ca); row_col(b, &rb, &cb);
    if (ra != rb) return ra < rb;
    return ca < cb;
  });
  const long long total = (long long)buf.size();
  if (out) for (long long i = 0; i < total && i < cap; i++) out[i] = buf[(size_t)i];
  return total;
}

int pg_large_set_players(PgLarge *w, const PgBody *players, int n) {
  if (!w || n < 0 || n > 8 || (n > 0 && !players)) { pg_set_error("pg_large_set_players: bad argument (at most 8 players)"); return PG_ERR_ARG; }
  if (!w->player_world) { pg_set_error("pg_large_set_players: set the level (pg_large_set with static bodies) first"); return PG_ERR_ARG; }
  // Passes 1-2 run the players against the LEVEL only: that is the whole of the reference's result when every
  // player is a capsule (scene.c:311-318), because the sphere-capsule table entries are empty functions
  // upstream and never fire.  A sphere or AABB player has live entries against the dynamic spheres
  // (world.c:277-367) which this path does not evaluate -- refuse it instead of silently dropping contacts.
  for (int i = 0; i < n; i++)
    if (players[i].collider_type != PG_CAPSULE) {
      pg_set_error("pg_large_set_players: player %d is not a capsule (collider type %d); only capsule players are supported in a large world", i, players[i].collider_type);
      return PG_ERR_UNSUPPORTED;
    }
  int rc = pg_worlds_set_bodies(w->player_world, 0, PG_CLASS_PLAYER, players, n);
  if (rc) return rc;
  w->n_players = n;
  return PG_OK;
}

int pg_large_get_players(PgLarge *w, PgBody *out, int cap) {
  if (!w || !out) return PG_ERR_ARG;
  if (!w->player_world || w->n_players == 0) return 0;
  return pg_worlds_get_bodies(w->player_world, 0, PG_CLASS_PLAYER, out, cap);
}

int pg_large_stats(PgLarge *w, PgStats *host_out, void *dev_out8) {
  if (!w) return PG_ERR_ARG;
  PG_CUDA(cudaSetDevice(w->device));
  PG_CUDA(cudaMemsetAsync(w->d_stats, 0, 8 * sizeof(double), w->stream));
  if (dev_out8) PG_CUDA(cudaMemsetAsync(dev_out8, 0, 8 * sizeof(double), w->stream));
  k_lw_stats<<<w->sm_count * 4, 256, 0, w->stream>>>(w->v, w->d_stats, (double *)dev_out8);
  w->launches++;
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

long long pg_large_launches(const PgLarge *w) { return w ? w->launches : 0; }

float pg_large_last_step_ms(PgLarge *w) {
  if (!w) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(w->ev1) != cudaSuccess) return -1.0f;
  if (cudaEventElapsedTime(&ms, w->ev0, w->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

int pg_large_kernel_count(void) { return K_COUNT; }
const char *pg_large_kernel_name(int k) { return (k >= 0 && k < K_COUNT) ? kKernelNames[k] : ""; }
float pg_large_kernel_ms(PgLarge *w, int k) { return (w && k >= 0 && k < K_COUNT) ? w->kms[k] : -1.0f; }
int pg_large_set_timing(PgLarge *w, int on) {
  if (!w) return PG_ERR_ARG;
  w->timing = on ? 1 : 0;
  return PG_OK;
}

int pg_large_exactness(PgLarge *w, int *iterations, int *proven_exact, long long *candidates, long long *hits) {
  if (!w) return PG_ERR_ARG;
  if (iterations) *iterations = w->last_iterations;
  if (proven_exact) *proven_exact = w->last_proven;
  if (candidates) *candidates = w->last_candidates;
  if (hits) *hits = w->last_hits;
  return PG_OK;
}

}  // extern "C"
