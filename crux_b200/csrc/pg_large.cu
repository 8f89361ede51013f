// pg_large.cu -- physics_step for ONE large world of dynamic spheres + static planes (configs C2, C4).
//
// The reference visits all N(N-1) ordered pairs in array order, mutating bodies in place
// (world.c:473-554).  Here the same RESULT is produced without the quadratic sweep and without
// giving up the solve order:
//
//   1. pass 3 (spheres x statics) touches only the sphere -> one thread per sphere: a streaming probe
//      kernel with the cheap gates, then the exact visits of the few bodies that pass one
//                                                                        [k_pass3, k_pass3_exact]
//   2. uniform grid over the post-pass-3 centres; every body carries an envelope
//        env = r + 2 (|v| + g dt) dt (1+eps)
//      that bounds where it can be, plus how far interval_collision can reach from there, at ANY
//      moment of pass 4 as long as no contact changes it.  Two bodies whose envelopes do not
//      touch can never pass the root test of interval_collision (world.c:559-565), so only pairs
//      from the 27-cell neighbourhood with touching envelopes are candidates.
//                                           [k_cell_count, scan, k_scatter (counting sort), k_pairs]
//   3. ORDER-PRESERVING FIXED POINT.  Each body keeps a timeline: the contacts it has suffered so
//      far, keyed by (row, column) = the position of the visit in the reference's sweep, with the
//      state they left behind.  A candidate pair is evaluated as its two visits (i,j) and (j,i)
//      on the states the timelines imply AT THAT KEY (including whether the body's own row, hence
//      its integration, lies before the key).  The hits found form the next timelines; iterate
//      until nothing changes.  The earliest event of the true sequential sweep is found in the
//      first iteration and never changes again; by induction iteration k fixes the k-th link of
//      every dependency chain, so the fixed point IS the sequential result -- bit for bit, since
//      every visit runs the exact arithmetic of pg_math.cuh.  An iteration only looks at what can
//      have changed: k_begin copies the entries of bodies whose timeline did not change (with
//      partners that did not either) and lists the pairs that need a look; k_eval evaluates them,
//      one thread per pair, and hands the rare long interval searches to k_eval_hard (a block per
//      pair); k_check compares the timelines of the bodies involved.
//                                                          [k_begin, k_eval, k_eval_hard, k_check]
//   4. envelopes are re-validated against the new timelines (a contact may have sped a body
//      up); if one grew, the pairs it adds are appended and the iteration continues warm.
//                                                       [k_check, k_requery, k_requery_wide]
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
const char *kKernelNames[K_COUNT] = {"k_pass3", "k_requery", "k_cell_count", "k_scan", "k_scatter", "k_pairs", "k_eval", "k_check", "k_finalize"};

struct GridParams {
  float org[3];
  float inv_h;
  int dim[3];
  int n_cells;
  int ok;
  float e_cap;            // envelope the cell edge is sized for; bodies beyond it are "fast" (few) and search wider
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
  int grew;               // an envelope had to grow
  int overflow_pairs;
  int overflow_timeline;
  int unsafe;             // non-finite state or an envelope the grid cannot hold
  unsigned long long n_pairs;
  unsigned long long n_hits;
  unsigned long long n_carried;   // hits d_clear_carry copied this iteration; folded into n_hits by k_check
  unsigned n_touched;
  unsigned n_fast;        // bodies on the fast list
  unsigned n_active;      // candidate pairs the selection (k_begin) handed to k_eval this iteration
  unsigned n_hard;        // of those, pairs k_eval passed on to k_eval_hard (long interval searches)
  unsigned n_wide;        // grown bodies k_requery passed on to k_requery_wide
  unsigned n_pass3;       // bodies k_pass3 passed on to k_pass3_exact
};

// One body of the cell-ordered copy.  A whole 32-byte sector per body: the counting sort's scattered
// store then never leaves a partial sector behind (16 B + 4 B stores into shared sectors made the L2
// fetch the rest from DRAM: 2.6x the algorithmic traffic), and a reader gets the index with the centre.
struct __align__(32) SortedBody { float4 cenv; unsigned idx, pad0, pad1, pad2; };

struct LargeView {
  int n, ns;
  int stage_limit;                   // entries the per-block shared-memory stages may hold (tests shrink it to reach the overflow paths)
  int requery_few;                   // k_requery: below this many grown bodies each gets a block
  int pass3_split_min;               // worlds of at least this many bodies run pass 3 as probe + dense exact kernel
  float4 *pos_rest, *vel_rad;        // state (in/out), body order = reference array order
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
  float *env;                        // envelope radius per body
  // reductions
  unsigned *bbox;                    // 6 ordered-uint encodings: min xyz, max xyz ; [6] = max env
  float *rstat;                      // [8] sums over the bodies inside the previous robust box: x y z |dx| |dy| |dz| env count
  Robust *robust;
  int *fast_list;                    // [n] bodies whose envelope exceeds grid->e_cap
  int *wide_list;                    // [n] grown bodies whose search is wider than 27 cells (k_requery_wide)
  GridParams *grid;
  // grid
  unsigned *cell_of;                 // [n]
  unsigned *cell_cnt;                // [cells_cap + 1]  counts -> exclusive starts
  unsigned *block_sums;              // scan scratch
  int cells_cap;
  SortedBody *sorted;                // [n]  bodies in cell order: centre xyz, env, body index -- one 32-byte sector each
  // candidates
  int2 *pairs;
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
  unsigned *cell_rank;                // rank of a body inside its cell (returned by the counting atomic)
  unsigned char *grown;               // envelope grew during the current fixed point
  float *env_prev;                   // envelope the current candidate list was built with
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

__device__ __forceinline__ float envelope(float r, float vx, float vy, float vz, float dt) {
  // r + 2 (|v| + g dt) dt, rounded up: displacement of one integration + the predicate's reach
  const float sp = sqrtf(__fmaf_rn(vx, vx, __fmaf_rn(vy, vy, vz * vz)));
  return (r + 2.0f * (sp + pgm::kGravity * dt) * dt) * 1.0002f + 2.0e-6f;
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
    const float e = envelope(q.w, q.x, q.y, q.z, dt);
    v.env[i] = e;
    v.env_prev[i] = e;
    if (full_reset) {
      // per-step bookkeeping of the fixed point: only bodies that get a timeline entry ever change it, and
      // k_reset_touched puts those back at the end of the step (23 B per body and step not written here)
      v.grown[i] = 0;
      v.tl_cnt[0][i] = 0;
      v.tl_cnt[1][i] = 0;
      v.tl_chg[0][i] = 0xffffffffffffffffull;
      v.tl_chg[1][i] = 0xffffffffffffffffull;
      v.seen[i] = 0;
    }
    if (!(fabsf(p.x) < 1.0e15f && fabsf(p.y) < 1.0e15f && fabsf(p.z) < 1.0e15f && e < 1.0e15f)) { bad = true; continue; }
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    emax = fmaxf(emax, e);
    const float dx = fabsf(p.x - rb.centre[0]), dy = fabsf(p.y - rb.centre[1]), dz = fabsf(p.z - rb.centre[2]);
    if (dx <= rb.half[0] && dy <= rb.half[1] && dz <= rb.half[2]) {
      rs[0] += p.x; rs[1] += p.y; rs[2] += p.z; rs[3] += dx; rs[4] += dy; rs[5] += dz; rs[6] += e; rs[7] += 1.0f;
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
    if (rb.valid) e_cap = fminf(emax, 4.0f * (v.rstat[6] / cnt) + 1.0e-3f);
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
  g.e_cap = 0.5f * h;                          // the largest envelope the 27-cell search serves: 2 env <= h (25 % above the sizing envelope)
  *v.grid = g;
}

__device__ __forceinline__ int3 cell_coords(const GridParams &g, float x, float y, float z) {
  int3 c;
  c.x = min(max((int)((x - g.org[0]) * g.inv_h), 0), g.dim[0] - 1);
  c.y = min(max((int)((y - g.org[1]) * g.inv_h), 0), g.dim[1] - 1);
  c.z = min(max((int)((z - g.org[2]) * g.inv_h), 0), g.dim[2] - 1);
  return c;
}

// ---------------------------------------------------------------- 2. counting sort into cells
__global__ void k_cell_count(LargeView v) {
  const GridParams g = *v.grid;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
    const float4 p = v.pos_rest[i];
    const int3 c = cell_coords(g, p.x, p.y, p.z);
    const unsigned cell = (unsigned)((c.z * g.dim[1] + c.y) * g.dim[0] + c.x);
    v.cell_of[i] = cell;
    v.cell_rank[i] = atomicAdd(v.cell_cnt + cell, 1u);
    if (v.env[i] > g.e_cap) v.fast_list[atomicAdd(&v.flags->n_fast, 1u)] = i;
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

__global__ void k_scatter(LargeView v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += gridDim.x * blockDim.x) {
    const unsigned cell = v.cell_of[i];
    const unsigned slot = v.cell_cnt[cell] + v.cell_rank[i];
    const float4 p = v.pos_rest[i];
    const float e = v.env[i];
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(v.sorted + slot), "r"(__float_as_uint(p.x)), "r"(__float_as_uint(p.y)),
                 "r"(__float_as_uint(p.z)), "r"(__float_as_uint(e)), "r"((unsigned)i), "r"(0u), "r"(0u), "r"(0u) : "memory");
  }
}

// ---------------------------------------------------------------- candidate pairs
// A pair is listed by its body with the LARGER envelope (ties: lower index), which therefore only has
// to find partners within 2 env of itself: R = ceil(2 env / h) cells each way -- 1 for everybody the
// grid was sized for (27 cells = 9 contiguous x-runs), more for the few fast bodies, and the whole
// sorted array for a body so fast that its search would cover most of the grid anyway.
constexpr int kMaxSearch = 12;
constexpr int kPairsBlock = 128;
constexpr int kPairsStage = 1024;       // pairs a block collects in shared memory between two flushes
// One thread per body (in cell order, so neighbouring threads read neighbouring cells); every thread
// walks its own cell runs without warp votes.  Two bodies the grid was sized for (R = 1 both) lie in the
// same or in adjacent cells, so each of them searches only the HALF shell of cells that follow its own
// in cell order (the rest of its own x-run, the next y-row, the three rows of the next z-layer: 5 runs,
// 13.5 cells instead of 27) and lists whatever it finds there; a body with a larger envelope searches
// its whole neighbourhood and lists the partners whose envelope is smaller (ties: lower index), and an
// R = 1 body leaves such partners to them.  Pairs are staged in shared memory and appended to the global
// list once per tile of kPairsBlock bodies; a tile that finds more than kPairsStage pairs (a dense pile)
// appends the surplus directly.
__global__ void __launch_bounds__(kPairsBlock) k_pairs(LargeView v) {
  __shared__ int2 s_pairs[kPairsStage];
  __shared__ unsigned s_cnt;
  __shared__ unsigned long long s_base;
  const GridParams g = *v.grid;
  const int n_tiles = (v.n + kPairsBlock - 1) / kPairsBlock;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int a = tile * kPairsBlock + threadIdx.x;
    if (a < v.n) {
      const float4 ca = v.sorted[a].cenv;
      const unsigned ia = v.sorted[a].idx;
      const int3 c = cell_coords(g, ca.x, ca.y, ca.z);
      const float need = 2.0f * ca.w * g.inv_h * 1.001f;
      const int R = need <= 1.0f ? 1 : (need <= (float)kMaxSearch ? (int)ceilf(need) : 0);
      const bool everything = !(need <= (float)kMaxSearch);
      const bool half = need <= 1.0f;
      auto run = [&](unsigned b0, unsigned b1) {
        // (the loads of up to four partners are issued before the first is looked at: the kernel is bound by
        // the latency of its dependent loads, not by instructions or bytes)
        float4 nb[4];
        for (unsigned bb = b0; bb < b1; bb += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) nb[u] = v.sorted[min(bb + u, b1 - 1)].cenv;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const unsigned b = bb + u;
          if (b >= b1) break;
          const float4 cb = nb[u];
          const float dx = ca.x - cb.x, dyy = ca.y - cb.y, dzz = ca.z - cb.z;
          const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dyy, dyy, dzz * dzz));
          const float T = ca.w + cb.w;
          if (!(d2 <= T * T)) continue;
          const unsigned ib = v.sorted[b].idx;
          if (half) {
            if (!(2.0f * cb.w * g.inv_h * 1.001f <= 1.0f)) continue;          // a wider-searching partner lists the pair itself
          } else {
            if (ca.w < cb.w || ia == ib || (ca.w == cb.w && !(ia < ib))) continue;
          }
          const int2 pr = make_int2((int)min(ia, ib), (int)max(ia, ib));
          const unsigned slot = atomicAdd(&s_cnt, 1u);
          if (slot < (unsigned)min(v.stage_limit, kPairsStage)) { s_pairs[slot] = pr; continue; }
          const unsigned long long gs = atomicAdd(&v.flags->n_pairs, 1ull);
          if ((long long)gs < v.pairs_cap) v.pairs[gs] = pr;
          else atomicExch(&v.flags->overflow_pairs, 1);
        }
        }
      };
      if (everything) {
        run(0u, (unsigned)v.n);
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
// `skip_key` (if non-zero) names an entry to ignore and `ovr` an entry to use in its place: the
// fresh outcome of the pair's first visit when its second visit is evaluated (Gauss-Seidel inside
// a pair; the fixed points are the same as without it).
struct Override { unsigned long long key; bool present; float px, py, pz, vx, vy, vz; };

__device__ __forceinline__ BodyState state_at(const LargeView &v, int src, int b, unsigned long long key, float dt, float *radius,
                                             unsigned long long skip_key, const Override *ovr) {
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
    if (kk == skip_key) continue;
    if (kk < key && (best_k < 0 || kk > best)) { best = kk; best_k = k; }
  }
  if (best_k >= 0) {
    const float4 tp = v.tl_pos[src][(size_t)b * kTimeline + best_k], tv = v.tl_vel[src][(size_t)b * kTimeline + best_k];
    s.px = tp.x; s.py = tp.y; s.pz = tp.z; s.vx = tv.x; s.vy = tv.y; s.vz = tv.z;
    last_row = (long long)(best >> 32);
  }
  if (ovr && ovr->present && ovr->key < key && (best_k < 0 || ovr->key > best)) {
    s.px = ovr->px; s.py = ovr->py; s.pz = ovr->pz; s.vx = ovr->vx; s.vy = ovr->vy; s.vz = ovr->vz;
    last_row = (long long)(ovr->key >> 32);
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
// (A = row body, B = column body after it), 2 = deferred (one thread per pair only: the interval search
// is a long one, the pair goes to a block).  Warp: all threads of the block, identical arguments.
template <bool Warp>
__device__ __forceinline__ int visit_compute(const LargeView &v, int src, int row, int col, float dt, pgm::LeafLen leaf,
                                             unsigned long long skip_key, const Override *in_row, const Override *in_col,
                                             Sphere &A, Sphere &B) {
  const unsigned long long key = ((unsigned long long)(unsigned)row << 32) | (unsigned)col;
  float ra, rb;
  const BodyState a = state_at(v, src, row, key, dt, &ra, skip_key, in_row), b = state_at(v, src, col, key, dt, &rb, skip_key, in_col);
  A.px = a.px; A.py = a.py; A.pz = a.pz; A.vx = a.vx; A.vy = a.vy; A.vz = a.vz; A.r = ra; A.rest = 0;
  B.px = b.px; B.py = b.py; B.pz = b.pz; B.vx = b.vx; B.vy = b.vy; B.vz = b.vz; B.r = rb; B.rest = 0;
  if (Warp) return pgs::sphere_pair_exact_block(A, B, dt, leaf) ? 1 : 0;
  return pgs::sphere_pair_exact_try(A, B, dt, leaf);
}
__device__ __forceinline__ void set_override(Override *o, unsigned long long key, const Sphere &S) {
  o->key = key; o->present = true; o->px = S.px; o->py = S.py; o->pz = S.pz; o->vx = S.vx; o->vy = S.vy; o->vz = S.vz;
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
// (entries with key < K other than `except`: new or different state -> flagged in .w; gone -> tl_chg)
__device__ __forceinline__ bool affected(const LargeView &v, int src, int b, unsigned long long K, unsigned long long except) {
  const unsigned long long rem = v.tl_chg[src][b];
  if (rem < K) return true;                    // (conservative when rem == except: other removals may hide behind it)
  const int cnt = min((int)v.tl_cnt[src][b], kTimeline);
  for (int k = 0; k < cnt; k++) {
    const unsigned long long kk = v.tl_key[src][(size_t)b * kTimeline + k];
    if (kk < K && kk != except && v.tl_pos[src][(size_t)b * kTimeline + k].w != 0.0f) return true;
  }
  return false;
}

// Which candidate pairs does this iteration have to look at?  A light, fully occupied pass over the
// whole list so that k_eval (122 registers, long divergent evaluations) runs on dense warps only.
//   iteration 0   the timelines are empty, so both visits of a pair see the post-pass-3 states (the
//                 second one its column body integrated): a pair both of whose closest-approach gates
//                 -- the very gate sphere_pair_exact opens with -- reject cannot fire and is dropped.
//                 If the first gate passes the pair is kept whatever the second says (a hit would
//                 change what the second visit sees).
//   later         pairs with a DIRTY body -- one whose timeline changed in the previous iteration (k_check)
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
      const int i = fresh ? ~pr.x : pr.x, j = pr.y;
      bool need = fresh;
      if (!need && iter > 0) {
        need = ((v.dirty_bits[src][i >> 5] >> (i & 31)) & 1u) || ((v.dirty_bits[src][j >> 5] >> (j & 31)) & 1u);
      } else if (!need) {
        const float4 pi = v.pos_rest[i], qi = v.vel_rad[i], pj = v.pos_rest[j], qj = v.vel_rad[j];
        // visit (i, j) of row i (i < j): neither body integrated yet
        const V3 cI = v3(0.0f + pi.x, 0.0f + pi.y, 0.0f + pi.z), cJ = v3(0.0f + pj.x, 0.0f + pj.y, 0.0f + pj.z);
        const V3 vI = v3(qi.x, qi.y, qi.z), vJ = v3(qj.x, qj.y, qj.z);
        const float rs = qi.w + qj.w, rs2 = qj.w + qi.w;
        const float nI = pgm::norm(vI), nJ = pgm::norm(vJ);
        need = !pgs::pair_gate_rejects(cI, vI, cJ, vJ, rs, nI + nJ, dt);
        if (!need) {
          // visit (j, i) of row j: body i has been integrated by its own row unless it rests (state_at)
          V3 p2 = v3(pi.x, pi.y, pi.z), v2 = vI;
          if (pi.w == 0.0f) pgm::integrate(p2, v2, dt);
          const V3 c2 = v3(0.0f + p2.x, 0.0f + p2.y, 0.0f + p2.z);
          need = !pgs::pair_gate_rejects(cJ, vJ, c2, v2, rs2, nJ + pgm::norm(v2), dt);
        }
      }
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

// Both visits of candidate pair k on the hypothesis `src`; hits go to the `dst` timelines.  Returns the
// number of hits, or -1 when a one-thread evaluation met a long interval search: nothing has been
// written then and the pair is to be repeated by a whole block (Warp = true: all its threads run this
// with the same k, `writer` = thread 0 does the writes).
template <bool Warp>
__device__ __forceinline__ int eval_pair(const LargeView &v, int src, int dst, int iter, float dt, pgm::LeafLen leaf, long long k, bool writer) {
  const int2 pr = v.pairs[k];
  const bool fresh = pr.x < 0;                     // appended by k_requery: never evaluated yet
  const int i = fresh ? ~pr.x : pr.x, j = pr.y;
  const unsigned long long key1 = ((unsigned long long)(unsigned)i << 32) | (unsigned)j;
  const unsigned long long key2 = ((unsigned long long)(unsigned)j << 32) | (unsigned)i;
  Sphere A1, B1, A2, B2;
  if (iter > 0 && !fresh) {
    // neither body has a timeline: nothing to carry, nothing that could have changed
    if (!((v.touched_bits[i >> 5] >> (i & 31)) & 1u) && !((v.touched_bits[j >> 5] >> (j & 31)) & 1u)) return 0;
    // A visit depends only on the entries with SMALLER keys of its two bodies.  If none of those
    // appeared, changed or disappeared since the previous iteration, its outcome is the recorded
    // one: carry it over instead of re-evaluating.  The second visit saw the first one's outcome
    // directly (Gauss-Seidel), so that entry is excluded from its test.
    const bool need1 = affected(v, src, i, key1, 0ull) || affected(v, src, j, key1, 0ull);
    if (!need1) {
      const bool need2 = affected(v, src, i, key2, key1) || affected(v, src, j, key2, key1);
      int h2 = 0;
      if (need2) {
        h2 = visit_compute<Warp>(v, src, j, i, dt, leaf, 0ull, nullptr, nullptr, A2, B2);
        if (h2 == 2) return -1;
      }
      int hits = 0;
      if (writer) {
        const bool have_i = v.tl_cnt[src][i] != 0, have_j = v.tl_cnt[src][j] != 0;
        if (have_i && have_j && carry_entry(v, src, dst, i, key1)) { carry_entry(v, src, dst, j, key1); hits++; }
        if (!need2) {
          if (have_i && have_j && carry_entry(v, src, dst, i, key2)) { carry_entry(v, src, dst, j, key2); hits++; }
        } else if (h2) {
          timeline_push(v, dst, j, key2, A2); timeline_push(v, dst, i, key2, B2); hits++;
        }
      }
      return hits;
    }
  }
  // visit (i, j) of row i, then visit (j, i) of row j on top of the first one's fresh outcome
  Override oi, oj;
  oi.present = false; oj.present = false; oi.key = key1; oj.key = key1;
  const int h1 = visit_compute<Warp>(v, src, i, j, dt, leaf, 0ull, nullptr, nullptr, A1, B1);
  if (h1 == 2) return -1;
  if (h1) { set_override(&oi, key1, A1); set_override(&oj, key1, B1); }
  const int h2 = visit_compute<Warp>(v, src, j, i, dt, leaf, key1, &oj, &oi, A2, B2);
  if (h2 == 2) return -1;
  if (writer) {
    if (fresh) v.pairs[k].x = i;
    if (h1) { timeline_push(v, dst, i, key1, A1); timeline_push(v, dst, j, key1, B1); }
    if (h2) { timeline_push(v, dst, j, key2, A2); timeline_push(v, dst, i, key2, B2); }
  }
  return h1 + h2;
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
    v.flags->n_wide = 0u;                               // for this iteration's k_requery
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
    // envelope validation on the new timeline: every state the body takes must stay inside env[b]
    const float4 p = v.pos_rest[b], q = v.vel_rad[b];
    float need = 0.0f;
    for (int k = 0; k < c1; k++) {
      const unsigned long long key = v.tl_key[dst][(size_t)b * kTimeline + k];
      const float4 tp = v.tl_pos[dst][(size_t)b * kTimeline + k], tv = v.tl_vel[dst][(size_t)b * kTimeline + k];
      const float dx = tp.x - p.x, dy = tp.y - p.y, dz = tp.z - p.z;
      const float dev = sqrtf(dx * dx + dy * dy + dz * dz);
      const float sp = sqrtf(tv.x * tv.x + tv.y * tv.y + tv.z * tv.z);
      const bool before_own_row = (long long)(key >> 32) <= (long long)b;
      // before its own row the body will still be integrated once: one more displacement and speed bump
      const float reach = before_own_row ? 2.0f * (sp + pgm::kGravity * dt) * dt : sp * dt;
      need = fmaxf(need, (q.w + dev + reach) * 1.0002f + 2.0e-6f);
    }
    if (need > v.env[b]) {
      const float e_cap = v.grid->e_cap;
      if (!(v.env[b] > e_cap) && need * 1.05f > e_cap) v.fast_list[atomicAdd(&v.flags->n_fast, 1u)] = b;   // joins the fast bodies
      v.env[b] = need * 1.05f;
      v.grown[b] = 1;
      atomicAdd(&v.flags->grew, 1);
      atomicMax(v.bbox + 6, f2o(need * 1.05f));      // largest envelope in the world, for k_requery's search radius
      if (!(need < 1.0e15f)) v.flags->unsafe = 1;
    }
  }
}
// Start of an iteration: the destination timelines of the touched bodies.  A DIRTY body (its timeline
// changed in the previous iteration) starts empty: every pair it belongs to is on the selection's list and
// k_eval rebuilds its entries.  A clean body keeps, by plain copy, the entries whose partner is clean as
// well -- visits nothing has changed for; its entries with a dirty partner come back through k_eval.
// From the third iteration on almost everybody is clean, and an iteration costs what its few changes cost.
__device__ __forceinline__ void d_clear_carry(const LargeView &v, int src, int dst, int iter) {
  const int nt = (int)v.flags->n_touched;
  if (blockIdx.x == 0 && threadIdx.x == 0) { v.flags->changed = 0; v.flags->grew = 0; v.flags->n_hits = 0ull; }      // (n_active, n_hard: k_check)
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

// Envelopes of a few bodies grew (a contact sped them up).  Find the candidate pairs they gain --
// touching under the new envelopes but not under the ones the list was built with -- through the
// existing grid and append them marked as fresh, so that the next iteration evaluates them (pairs
// that were already listed keep their recorded outcomes).  If both bodies of a pair grew, the lower
// index appends it.
// The search of one grown body b, shared by `stride` threads (this one is number `lane` of them).
// ByEntry = false: every thread takes whole cell runs (8 lanes, sparse cells: the runs' dependent loads are
// in flight together); true (a block): the warps take the runs, the lanes of a warp the bodies inside a run --
// for wide searches and for dense cells (a settled world has dozens of bodies per floor cell).
template <bool ByEntry>
__device__ __forceinline__ void requery_body(const LargeView &v, const GridParams &g, int b, int lane, int stride) {
  const float4 p = v.pos_rest[b];
  const float eb = v.env[b], eb_old = v.env_prev[b];
  auto consider = [&](int x, float4 cx) {
    const float dx = p.x - cx.x, dyy = p.y - cx.y, dzz = p.z - cx.z;
    const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dyy, dyy, dzz * dzz));
    const float ex = v.env[x], ex_old = v.env_prev[x];
    const float Tn = eb + ex, To = eb_old + ex_old;
    if (!(d2 <= Tn * Tn) || d2 <= To * To) return;            // not a candidate, or already listed
    if (v.grown[x] && x < b) return;                          // the other grown body appends it
    const unsigned long long slot = atomicAdd(&v.flags->n_pairs, 1ull);
    if ((long long)slot < v.pairs_cap) v.pairs[slot] = make_int2(~min(b, x), max(b, x));   // ~i marks it fresh
    else atomicExch(&v.flags->overflow_pairs, 1);
  };
  // Partners the grid was sized for (envelope <= e_cap) can be up to eb + e_cap away: search that many
  // cells in every direction.  Fast partners are few and are checked one by one from their list
  // (which also holds bodies that outgrew e_cap this step).
  const float need = (eb + g.e_cap) * g.inv_h * 1.001f;
  if (!(need <= (float)kMaxSearch)) {
    for (int a = lane; a < v.n; a += stride) {
      const int x = (int)v.sorted[a].idx;
      if (x != b && !(v.env[x] > g.e_cap)) consider(x, v.sorted[a].cenv);
    }
  } else {
    // every thread takes whole x-runs of the (2R+1)^2 the search covers (9 for a body the grid was sized
    // for), so their dependent loads -- cell table, cell contents, envelopes -- are in flight together
    const int R = max((int)ceilf(need), 1);
    const int3 c = cell_coords(g, p.x, p.y, p.z);
    const int side = 2 * R + 1;
    const int x0 = max(c.x - R, 0), x1 = min(c.x + R, g.dim[0] - 1);
    const int r_first = ByEntry ? (lane >> 5) : lane, r_step = ByEntry ? max(stride >> 5, 1) : stride;
    const unsigned a_first = ByEntry ? (unsigned)(lane & 31) : 0u, a_step = ByEntry ? 32u : 1u;
    for (int r = r_first; r < side * side; r += r_step) {
      const int y = c.y + r % side - R, z = c.z + r / side - R;
      if (y < 0 || y >= g.dim[1] || z < 0 || z >= g.dim[2]) continue;
      const unsigned row = (unsigned)((z * g.dim[1] + y) * g.dim[0]);
      const unsigned a1 = v.cell_cnt[row + x1 + 1];
      for (unsigned a = v.cell_cnt[row + x0] + a_first; a < a1; a += a_step) {
        const int x = (int)v.sorted[a].idx;
        if (x != b && !(v.env[x] > g.e_cap)) consider(x, v.sorted[a].cenv);
      }
    }
  }
  const int nf = (int)v.flags->n_fast;
  for (int f = lane; f < nf; f += stride) {
    const int x = v.fast_list[f];
    if (x == b || !(v.env[x] > g.e_cap)) continue;
    const float4 px = v.pos_rest[x];
    consider(x, make_float4(px.x, px.y, px.z, 0.0f));
  }
}
__device__ __forceinline__ void d_requery(const LargeView &v, int src) {
  const GridParams g = *v.grid;
  const int nt = (int)v.flags->n_touched;
  // Eight lanes per touched body whose search is the usual 9 short cell runs (each a chain of dependent
  // loads: more bodies in flight beat more lanes per body).  A body that searches wider -- a fast one, or a
  // straggler whose search covers most of the grid and falls back to the whole sorted array -- goes on a
  // list and gets a whole block (k_requery_wide): eight lanes walking 65 k bodies held the step up for 0.7 ms.
  constexpr int kGroup = 8;
  const int lane = threadIdx.x & (kGroup - 1);
  const int groups = (gridDim.x * blockDim.x) / kGroup;
  // when only a few bodies grew (every iteration but the first) there is a block to spare for each of them
  const bool few = v.flags->grew < v.requery_few;
  const unsigned group_mask = 0xffu << (threadIdx.x & 24);
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) / kGroup; t < nt; t += groups) {
    const int b = v.touched[t];
    if (!v.grown[b]) continue;
    const float need = (v.env[b] + g.e_cap) * g.inv_h * 1.001f;
    bool wide = few || !(need <= 1.0f);
    if (!wide) {
      // ... and so does a body whose 27 cells hold many bodies (a settled world has dozens per floor cell):
      // a lane would walk its runs for too long.  The lanes add up the lengths of their runs first.
      const float4 p = v.pos_rest[b];
      const int3 c = cell_coords(g, p.x, p.y, p.z);
      const int x0 = max(c.x - 1, 0), x1 = min(c.x + 1, g.dim[0] - 1);
      unsigned work = 0;
      for (int r = lane; r < 9; r += kGroup) {
        const int y = c.y + r % 3 - 1, z = c.z + r / 3 - 1;
        if (y < 0 || y >= g.dim[1] || z < 0 || z >= g.dim[2]) continue;
        const unsigned row = (unsigned)((z * g.dim[1] + y) * g.dim[0]);
        work += v.cell_cnt[row + x1 + 1] - v.cell_cnt[row + x0];
      }
      for (int o = 4; o > 0; o >>= 1) work += __shfl_xor_sync(group_mask, work, o);
      wide = work > 96u;                        // (measured: C4 at step 110 and the settled C2 are both slower with 512)
    }
    if (wide) {
      if (lane == 0) v.wide_list[atomicAdd(&v.flags->n_wide, 1u)] = b;
      continue;
    }
    requery_body<false>(v, g, b, lane, kGroup);
  }
}
__device__ __forceinline__ void d_requery_wide(const LargeView &v) {
  const GridParams g = *v.grid;
  const int nw = (int)v.flags->n_wide;
  for (int t = blockIdx.x; t < nw; t += gridDim.x) requery_body<true>(v, g, v.wide_list[t], (int)threadIdx.x, (int)blockDim.x);
}
__device__ __forceinline__ void d_requery_done(const LargeView &v) {
  const int nt = (int)v.flags->n_touched;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const int b = v.touched[t];
    if (v.grown[b]) { v.grown[b] = 0; v.env_prev[b] = v.env[b]; }
  }
}


__global__ void k_eval(LargeView v, int src, int dst, int iter, float dt) { d_eval(v, src, dst, iter, dt); }
__global__ void __launch_bounds__(512) k_eval_hard(LargeView v, int src, int dst, int iter, float dt) { d_eval_hard(v, src, dst, iter, dt); }
__global__ void k_check(LargeView v, int src, int dst, int iter, float dt) { d_check(v, src, dst, iter, dt); }
// One launch opens an iteration: the carry (per touched body) and the selection (per candidate pair) read the
// same source hypothesis and write disjoint things.
__global__ void __launch_bounds__(256) k_begin(LargeView v, int src, int dst, int iter, float dt) {
  d_clear_carry(v, src, dst, iter);
  d_select(v, src, iter, dt);
}
__global__ void k_requery(LargeView v, int src) { d_requery(v, src); }
__global__ void k_requery_wide(LargeView v) { d_requery_wide(v); }
__global__ void k_requery_done(LargeView v) { d_requery_done(v); }

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
    v.grown[b] = 0;
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
    const float e = pgs::reach(q.w, q.x, q.y, q.z, dt);      // r + |v| dt: the predicate's own reach, nothing moves
    v.env[i] = e;
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
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
    const int2 pr = v.pairs[k];
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
static int lw_build_pairs(PgLarge *w) {
  LargeView &v = w->v;
  k_grid_params<<<1, 1, 0, w->stream>>>(v);
  w->launches++;
  PG_CUDA(cudaMemsetAsync(v.cell_cnt, 0, sizeof(unsigned) * ((size_t)v.cells_cap + 1), w->stream));
  PG_CUDA(cudaMemsetAsync(&v.flags->n_pairs, 0, sizeof(unsigned long long), w->stream));
  LW_LAUNCH(w, K_CELLS, k_cell_count, lw_grid(w, v.n, 256), 256, v);
  // no read-back of the grid: the scan is launched for the largest table the buffers hold (blocks past
  // n_cells see zeros), and a grid that could not be built raises flags->unsafe for the next read-back
  const int scan_blocks = (v.cells_cap + 1 + kScanBlock * kScanItems - 1) / (kScanBlock * kScanItems);
  if (w->timing) cudaEventRecord(w->kev[K_SCAN][0], w->stream);
  k_scan_local<<<scan_blocks, kScanBlock, 0, w->stream>>>(v);
  k_scan_sums<<<1, 1024, 0, w->stream>>>(v, scan_blocks);
  k_scan_add<<<scan_blocks, kScanBlock, 0, w->stream>>>(v);
  if (w->timing) cudaEventRecord(w->kev[K_SCAN][1], w->stream);
  w->launches += 3;
  LW_LAUNCH(w, K_SCATTER, k_scatter, lw_grid(w, v.n, 256), 256, v);
  LW_LAUNCH(w, K_PAIRS, k_pairs, lw_grid(w, v.n, kPairsBlock), kPairsBlock, v);
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

static int lw_step_once(PgLarge *w, float dt_in) {
  LargeView &v = w->v;
  float dt = dt_in;
  if ((double)dt > 0.01666) dt = (float)0.01666;
  pg_leaf_lengths(0.0f, dt, &v.leaf_lo, &v.leaf_hi);
  const unsigned bbox_init[7] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u};
  PG_CUDA(cudaMemcpyAsync(v.bbox, bbox_init, sizeof(bbox_init), cudaMemcpyHostToDevice, w->stream));
  PG_CUDA(cudaMemsetAsync(v.flags, 0, sizeof(Flags), w->stream));
  PG_CUDA(cudaMemsetAsync(v.rstat, 0, sizeof(float) * 8, w->stream));
  PG_CUDA(cudaMemsetAsync(v.touched_bits, 0, sizeof(unsigned) * ((size_t)v.n / 32 + 1), w->stream));
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
  int rc = lw_build_pairs(w);
  if (rc) return rc;
  Flags f;
  rc = lw_read_flags(w, &f);            // pair count: sizes the eval grid, detects overflow
  if (rc) return rc;
  lw_accumulate(w, K_PASS3); lw_accumulate(w, K_CELLS); lw_accumulate(w, K_SCAN); lw_accumulate(w, K_SCATTER); lw_accumulate(w, K_PAIRS);
  if (f.overflow_pairs) {
    // k_pairs is a pure function of the grid: grow the buffer and run it again
    rc = lw_grow_pairs(w, f.n_pairs, 0);
    if (rc) return rc;
    LW_LAUNCH(w, K_PAIRS, k_pairs, lw_grid(w, v.n, kPairsBlock), kPairsBlock, v);
    rc = lw_read_flags(w, &f);
    if (rc) return rc;
    lw_accumulate(w, K_PAIRS);
    if (f.overflow_pairs) { pg_set_error("large world: candidate pair buffer too small (%lld)", v.pairs_cap); return PG_ERR_CAPACITY; }
  }
  if (f.unsafe) { pg_set_error("large world: non-finite body state"); return PG_ERR_ARG; }
  w->last_candidates = (long long)f.n_pairs;
  const bool debug = getenv("PG_LARGE_DEBUG") != nullptr && w->timing;
  bool converged = false;
  int requeries = 0;
  int src = 0, dst = 1;
  bool requery_pending = false;
  // One host read-back per iteration: k_requery always rides along (it returns at once for bodies whose
  // envelope did not grow), its bookkeeping kernel is deferred to the start of the next iteration so
  // that an overflowing candidate list can still be grown and the append repeated.
  const int wide_grid = w->sm_count * 8;
  for (int it = 0; it < kMaxIter; it++) {
    if (requery_pending) {
      k_requery_done<<<wide_grid, 256, 0, w->stream>>>(v);
      w->launches++;
      requery_pending = false;
    }
    const unsigned long long pairs_before = std::min<unsigned long long>(f.n_pairs, (unsigned long long)v.pairs_cap);
    // the active list's length stays on the device: k_eval runs as many blocks as fit at once and strides
    if (w->timing) cudaEventRecord(w->kev[K_EVAL][0], w->stream);
    k_begin<<<wide_grid, 256, 0, w->stream>>>(v, src, dst, it, dt);      // carry + selection; also resets the per-iteration flags
    k_eval<<<w->sm_count * 4, 128, 0, w->stream>>>(v, src, dst, it, dt);
    k_eval_hard<<<w->sm_count, 512, 0, w->stream>>>(v, src, dst, it, dt);
    if (w->timing) cudaEventRecord(w->kev[K_EVAL][1], w->stream);
    w->launches += 3;
    LW_LAUNCH(w, K_CHECK, k_check, wide_grid, 256, v, src, dst, it, dt);
    // candidates the grown envelopes add (on top of the NEW timelines); their bodies are marked dirty so
    // the next iteration re-evaluates them (warm start)
    if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][0], w->stream);
    k_requery<<<wide_grid, 128, 0, w->stream>>>(v, dst);
    k_requery_wide<<<wide_grid, 256, 0, w->stream>>>(v);
    if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][1], w->stream);
    w->launches += 2;
    rc = lw_read_flags(w, &f);
    if (rc) return rc;
    lw_accumulate(w, K_EVAL); lw_accumulate(w, K_CHECK); lw_accumulate(w, K_BOUNDS);
    if (debug) {
      float e = 0, c = 0, rq = 0;
      cudaEventElapsedTime(&e, w->kev[K_EVAL][0], w->kev[K_EVAL][1]);
      cudaEventElapsedTime(&c, w->kev[K_CHECK][0], w->kev[K_CHECK][1]);
      cudaEventElapsedTime(&rq, w->kev[K_BOUNDS][0], w->kev[K_BOUNDS][1]);
      fprintf(stderr, "  [pg_large] iter %d: eval %.3f ms check %.3f ms requery %.3f ms hits %llu changed %d grew %d pairs %llu -> %llu touched %u fast %u wide %u\n",
              it, e, c, rq, (unsigned long long)f.n_hits, f.changed, f.grew, pairs_before, (unsigned long long)f.n_pairs, f.n_touched, f.n_fast, f.n_wide);
    }
    w->last_iterations++;
    w->last_hits = (long long)f.n_hits;
    if (f.overflow_timeline) { pg_set_error("large world: more than %d contacts on one body in one step", kTimeline); return PG_ERR_CAPACITY; }
    std::swap(src, dst);
    if (f.grew) {
      if (f.overflow_pairs) {
        // k_requery only appends: keep the list as it was, make room, and let it append again
        rc = lw_grow_pairs(w, f.n_pairs, pairs_before);
        if (rc) return rc;
        PG_CUDA(cudaMemsetAsync(&v.flags->n_wide, 0, sizeof(unsigned), w->stream));
        if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][0], w->stream);
        k_requery<<<wide_grid, 128, 0, w->stream>>>(v, src);
        k_requery_wide<<<wide_grid, 256, 0, w->stream>>>(v);
        if (w->timing) cudaEventRecord(w->kev[K_BOUNDS][1], w->stream);
        w->launches += 2;
        rc = lw_read_flags(w, &f);
        if (rc) return rc;
        lw_accumulate(w, K_BOUNDS);
        if (f.overflow_pairs) { pg_set_error("large world: candidate pair buffer too small (%lld)", v.pairs_cap); return PG_ERR_CAPACITY; }
      }
      requery_pending = true;
      if (f.unsafe) { w->last_proven = 0; if (debug) fprintf(stderr, "  [pg_large] envelope outgrew the grid: result not proven\n"); }
      w->last_candidates = (long long)f.n_pairs;
      requeries++;
      continue;                           // at least one more iteration with the new pairs
    }
    if (!f.changed) { converged = true; break; }
  }
  if (requery_pending) {
    k_requery_done<<<wide_grid, 256, 0, w->stream>>>(v);
    w->launches++;
  }
  if (!converged) w->last_proven = 0;
  (void)requeries;
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
  v.requery_few = getenv("PG_LARGE_REQUERY_FEW") ? atoi(getenv("PG_LARGE_REQUERY_FEW")) : 4096;
  v.pass3_split_min = getenv("PG_LARGE_PASS3_SPLIT_MIN") ? atoi(getenv("PG_LARGE_PASS3_SPLIT_MIN")) : 262144;
  v.cells_cap = (int)std::min<size_t>(4 * n + 4096, (size_t)1 << 28);
  v.pairs_cap = (long long)(4 * n + 65536);
  v.max_events = (int)std::max<size_t>(n / 4, 65536);
  PG_CUDA_NULL(cudaMalloc(&v.pos_rest, sizeof(float4) * n));
  PG_CUDA_NULL(cudaMalloc(&v.vel_rad, sizeof(float4) * n));
  PG_CUDA_NULL(cudaMalloc(&v.planes, sizeof(float4) * kMaxPlanesL));
  PG_CUDA_NULL(cudaMalloc(&v.plane_sidx, sizeof(int) * kMaxPlanesL));
  w->cap_static_total = std::max(max_static, kMaxPlanesL);
  PG_CUDA_NULL(cudaMalloc(&v.box_c, sizeof(float4) * (size_t)w->cap_static_total));
  PG_CUDA_NULL(cudaMalloc(&v.box_e, sizeof(float4) * (size_t)w->cap_static_total));
  PG_CUDA_NULL(cudaMalloc(&v.static_type, (size_t)w->cap_static_total));
  PG_CUDA_NULL(cudaMalloc(&v.env, sizeof(float) * n));
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
  PG_CUDA_NULL(cudaMalloc(&v.fast_list, sizeof(int) * n));
  PG_CUDA_NULL(cudaMalloc(&v.wide_list, sizeof(int) * n));
  PG_CUDA_NULL(cudaMalloc(&v.cell_of, sizeof(unsigned) * n));
  PG_CUDA_NULL(cudaMalloc(&v.cell_cnt, sizeof(unsigned) * ((size_t)v.cells_cap + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.block_sums, sizeof(unsigned) * ((size_t)v.cells_cap / (kScanBlock * kScanItems) + 2)));
  PG_CUDA_NULL(cudaMalloc(&v.sorted, sizeof(SortedBody) * n));
  PG_CUDA_NULL(cudaMalloc(&v.pairs, sizeof(int2) * (size_t)v.pairs_cap));
  PG_CUDA_NULL(cudaMalloc(&v.active, sizeof(unsigned) * 2 * (size_t)v.pairs_cap));
  PG_CUDA_NULL(cudaMalloc(&v.grown, n + 8));
  PG_CUDA_NULL(cudaMemset(v.grown, 0, n + 8));
  PG_CUDA_NULL(cudaMalloc(&v.env_prev, sizeof(float) * n));
  PG_CUDA_NULL(cudaMalloc(&v.seen, sizeof(int) * n));
  PG_CUDA_NULL(cudaMalloc(&v.touched_bits, sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.dirty_bits[0], sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.dirty_bits[1], sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.pushed_bits, sizeof(unsigned) * (n / 32 + 1)));
  PG_CUDA_NULL(cudaMalloc(&v.touched, sizeof(int) * n));
  PG_CUDA_NULL(cudaMalloc(&v.cell_rank, sizeof(unsigned) * n));
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
  void *ptrs[] = {v.plane_sidx, v.box_c, v.box_e, v.static_type, v.sg_start, v.sg_items, v.pos_rest, v.vel_rad, v.planes, v.env, v.bbox, v.grid, v.cell_of, v.cell_cnt, v.rstat, v.robust, v.fast_list, v.wide_list, v.touched_bits, v.dirty_bits[0], v.dirty_bits[1], v.pushed_bits, v.block_sums, v.sorted,
                  v.pairs, v.active, v.grown, v.env_prev, v.tl_chg[0], v.tl_chg[1], v.seen, v.touched, v.cell_rank, v.tl_cnt[0], v.tl_cnt[1], v.tl_key[0], v.tl_key[1], v.tl_pos[0], v.tl_pos[1], v.tl_vel[0], v.tl_vel[1],
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
  k_env_frozen<<<lw_grid(w, v.n, 256), 256, 0, w->stream>>>(v, dt);
  w->launches++;
  int rc = lw_build_pairs(w);
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
