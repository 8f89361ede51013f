// pg_spheres.cu -- batched sphere worlds (config C3: 4096 x 256-sphere bouncehouse worlds).
//
// ONE WARP PER WORLD, state in shared memory, reference solve order preserved exactly.
//
// Body i of a world belongs to lane (i & 31), slot (i >> 5); a world holds up to 32*CPT spheres.
// The whole n_steps loop runs inside one launch: positions/velocities are read from HBM once and
// written once per launch (32 B + 32 B per body), everything in between is shared-memory, register
// and shuffle traffic.  A world's sequential sweep only ever needs warp-level votes; the one block
// barrier per step exists for the instruction cache (see PG_SPH_SYNC).
//
// Pass 4 of physics_step (world.c:473-554) is N rows; row i visits every column j != i, mutating
// i and j in place, then integrates i.  Rows are processed 32 at a time:
//   filter   every lane tests its CPT columns against the 32 rows with a CONSERVATIVE reach test
//            (|ci-cj|^2 against (rho_i+rho_j)^2, rho = (r + |v| dt)(1+1e-4), expanded form, two
//            columns per FFMA2): a pair that fails it cannot pass the root test of
//            interval_collision (world.c:559-565), so skipping it is exact.  It runs on the state
//            each row WILL see if no contact intervenes (earlier rows of the block integrated).
//   gate     the surviving (row, column) bits of the block are gated one pair per lane by the
//            closest approach of the two linear paths over [0,dt] against rs + delta.
//   visits   rows that still hold a bit are visited in order, their columns in increasing j
//            (warp min-reduce), with the exact predicate + narrowphase + resolver of pg_math.cuh,
//            redundantly on all lanes (warp-uniform control flow); rows in between are integrated
//            in bulk.
//   repair   a contact rewrites the bits of the two bodies it changed: the row re-filters its
//            remaining columns, later rows of the block re-test the changed columns (one lane per
//            row, a ballot is the new word), a changed later row rebuilds itself when its turn comes.
//
// Pass 3 (dynamic x static planes, world.c:369-471) touches only the dynamic body, so each lane
// runs it for its own bodies, planes in array order.
#include "pg_common.cuh"
#include "pg_math.cuh"
#include "pg_sphere_exact.cuh"

#include <algorithm>
#include <cstdlib>

namespace {

using pgm::V3;
using pgm::v3;

#ifndef PG_SPH_WARPS
#define PG_SPH_WARPS 16   // most warps (= worlds) per block, one block per SM (128 registers per thread); the launcher
#endif                    // picks the actual number per launch so that the batch fills whole waves of blocks
#ifndef PG_SPH_SYNC
#define PG_SPH_SYNC 1     // P > 0: a block barrier every P steps keeps the block's warps in phase, so that an instruction
#endif                    //    line fetched once serves all of them (the kernel is instruction-fetch bound); 0: none
constexpr int kMaxPlanes = 16;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kHuge = 1.0e8f;      // padding bodies: far from everything, squares still finite

using pgs::Sphere;
using pgs::reach;

using pgs::sphere_plane_exact;

__device__ __noinline__ void emit_event(const WorldsView &v, int w, int step, int pass, int ac, int ai, int bc, int bi) {
  int slot = atomicAdd(v.n_events + w, 1);
  if (slot >= v.max_events) return;
  PgEvent e;
  e.world = w; e.step = step; e.event_type = 0;       // WORLD x WORLD -> EVENT_COLLISION
  e.a_class = ac; e.a_index = ai; e.b_class = bc; e.b_index = bi;
  e.pad_ = pass << 8;
  v.events[(size_t)w * v.max_events + slot] = e;
}

// Shared-memory image of one world: every body as two float4.
//   P4[i] = (px, py, pz, rho)   rho = conservative reach (r + |v| dt)(1+1e-4), refreshed with v
//   V4[i] = (vx, vy, vz, r)

// Slow half of pass 3 for one body against one plane (per lane, divergent): refined gate, then the
// exact visit.  Returns true if the body changed; *rest is set when it came to rest.
__device__ __noinline__ bool plane_slow_path(float4 *P4, float4 *V4, int i, float4 pl, float restitution, float dt,
                                             bool gate_ok, int *rest, pgm::LeafLen leaf) {
  const float4 p = P4[i], q = V4[i];
  if (gate_ok) {
    // signed distance is linear in t; if the sphere surface stays farther than delta from the
    // plane over [0,dt] no leaf of the interval search can pass (see sphere_pair_exact).
    const float s0 = __fmaf_rn(p.x, pl.x, __fmaf_rn(p.y, pl.y, p.z * pl.z)) - pl.w;
    const float ndv = __fmaf_rn(q.x, pl.x, __fmaf_rn(q.y, pl.y, q.z * pl.z));
    const float s1 = __fmaf_rn(ndv, dt, s0);
    const float gap = (s0 > 0.0f) == (s1 > 0.0f) ? fminf(fabsf(s0), fabsf(s1)) : 0.0f;
    const float spd = sqrtf(__fmaf_rn(q.x, q.x, __fmaf_rn(q.y, q.y, q.z * q.z))) * 1.0001f;
    const float delta = 1.1e-6f * spd + 4.0e-6f * (fabsf(pl.w) + fabsf(p.x) + fabsf(p.y) + fabsf(p.z) + 2.0f * spd * dt) + 1.0e-6f;
    if (gap > (q.w + delta) * 1.00001f) return false;
  }
  Sphere A;
  A.px = p.x; A.py = p.y; A.pz = p.z; A.vx = q.x; A.vy = q.y; A.vz = q.z; A.r = q.w; A.rest = 0;
  if (!sphere_plane_exact(A, pl, restitution, dt, leaf)) return false;
  P4[i] = make_float4(A.px, A.py, A.pz, reach(A.r, A.vx, A.vy, A.vz, dt));
  V4[i] = make_float4(A.vx, A.vy, A.vz, A.r);
  if (A.rest) *rest = 1;
  return true;
}

// ---- pass 4 helpers -------------------------------------------------------------------------------
// packed binary32 pairs (Blackwell FFMA2 / FADD2): two columns of the reach filter per instruction
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(unsigned long long x, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// The expanded reach test |q_r - q_c|^2 - (rho_r + rho_c)^2 < 0 is evaluated with ~10 roundings on terms of
// magnitude <= 2 (|q_r|^2 + |q_c|^2): error <= 1.2e-6 (|q_r|^2 + |q_c|^2).  The slack kFilterSlack (|q_r|^2 + |q_c|^2) + 1e-6
// covers it six times over, per PAIR, so that one far-away body does not blunt the filter for the others.
constexpr float kFilterSlack = 8.0e-6f;
constexpr int kMaxRowChanges = 8;
constexpr int kGateList = 128;        // candidate bits of one 32-row block gated per pass

// All visits of row i, in increasing column order.  `mask` bit s = this lane's column s*32+lane is a
// candidate that already passed the reach filter AND the closest-approach gate for the state the
// row sees; with `fresh` the row body changed since those ran and the candidates are rebuilt here
// from shared memory, which holds the TRUE current state (rows < i integrated).  Columns changed by
// a contact are appended to chg[]; returns their number (> kMaxRowChanges: list overflowed).
__device__ __noinline__ int row_visits(float4 *P4, float4 *V4, int i, unsigned mask, bool fresh, int lane, int ncols_slots,
                                       float dt, const WorldsView &v, int w, int step, bool record, int *chg, int *contacts) {
  __syncwarp();
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;
  int n_chg = 0;
  int j0 = 0;
  for (;;) {
    if (fresh) {
      // reach filter + closest-approach gate of this lane's remaining columns against the row body
      mask = 0;
      const float4 ap = P4[i];
      for (int s = j0 >> 5; s < ncols_slots; s++) {
        const int c = s * 32 + lane;
        const float4 q = P4[c];
        const float dx = ap.x - q.x, dy = ap.y - q.y, dz = ap.z - q.z;
        const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz));
        const float T = ap.w + q.w;
        if (d2 <= T * T && c != i && c >= j0 && !pgs::pair_gate_far(P4, V4, i, c, false, dt)) mask |= 1u << s;
      }
      fresh = false;
    }
    // lowest candidate column >= j0 of this lane: slot bits below the first admissible slot are dropped
    const int smin = (j0 >> 5) + ((lane < (j0 & 31)) ? 1 : 0);
    const unsigned m = smin < 32 ? (mask >> smin) << smin : 0u;
    const int jl = m ? (__ffs((int)m) - 1) * 32 + lane : 0x7fffffff;
    const int j = __reduce_min_sync(kFull, jl);
    if (j == 0x7fffffff) break;
    const float4 ap = P4[i], av = V4[i], bp = P4[j], bv = V4[j];
    Sphere A, B;
    A.px = ap.x; A.py = ap.y; A.pz = ap.z; A.vx = av.x; A.vy = av.y; A.vz = av.z; A.r = av.w; A.rest = 0;
    B.px = bp.x; B.py = bp.y; B.pz = bp.z; B.vx = bv.x; B.vy = bv.y; B.vz = bv.z; B.r = bv.w; B.rest = 0;
    if (pgs::sphere_pair_exact_gated(A, B, dt, leaf)) {
      __syncwarp();
      if (lane == 0) {
        P4[i] = make_float4(A.px, A.py, A.pz, reach(A.r, A.vx, A.vy, A.vz, dt)); V4[i] = make_float4(A.vx, A.vy, A.vz, A.r);
        P4[j] = make_float4(B.px, B.py, B.pz, reach(B.r, B.vx, B.vy, B.vz, dt)); V4[j] = make_float4(B.vx, B.vy, B.vz, B.r);
        if (record) emit_event(v, w, step, 4, PG_CLASS_DYNAMIC, i, PG_CLASS_DYNAMIC, j);
        *contacts += 1;
        if (n_chg < kMaxRowChanges) chg[n_chg] = j;
      }
      n_chg++;
      __syncwarp();
      fresh = true;                                  // the row body changed: rebuild the remaining candidates
    }
    j0 = j + 1;
  }
  return n_chg;
}

// Column c changed (contact) while row row0+rl was being swept: which of the block's LATER rows
// (bit l, l > rl) now have it as a candidate?  A later row r sees column c integrated iff c < r;
// shared memory already holds it integrated iff c < done.
__device__ __noinline__ unsigned recheck_column(const float4 *P4, const float4 *V4, int row0, int rl, int c, bool c_rest, int done,
                                                int n, float dt, int lane) {
  const int r = row0 + lane;
  bool pass = false;
  if (lane > rl && r < n && r != c) {
    const bool integ = c >= done && c < r && !c_rest;
    pass = !pgs::pair_gate_far(P4, V4, r, c, integ, dt);
  }
  return __ballot_sync(kFull, pass);
}

// The filter's column registers are ROTATED: while row block `rb` (rows rb*32 .. rb*32+31) is swept,
// static slot k holds the body of actual slot (rb + k) mod CPT, so the block's own bodies always sit
// in static slot 0.  That keeps every register index a compile-time constant with ONE copy of the
// filter loop in the instruction stream.
//
// Shared-memory image of one world (one warp).
template <int CPT>
struct WarpSmem {
  float4 P4[CPT * 32];
  float4 V4[CPT * 32];
  float4 planes[kMaxPlanes];
  ulonglong2 RA[64];                               // filter coefficients of the 32 rows being swept, as duplicated pairs for FFMA2
  unsigned long long RB[32];
  unsigned W0[CPT * 32];                           // the filter's bits as they came out (enumeration order of the gate)
  unsigned W[CPT * 32];                       // candidate bits: [static slot][lane], bit = row of the block
  unsigned list[kGateList];
  unsigned rest[CPT];                       // at-rest flags by actual slot, bit = lane
  int chg[kMaxRowChanges];
  int n_list;                                      // gate list allocation counter (zero between uses)
  int contacts;                                    // contacts (events) of this world in this launch, for the statistics
};

template <int CPT, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1)
k_step_spheres(WorldsView v, float dt_in, int n_steps, int *work_counter, int first_world, int end_world,
               const int *order, unsigned *cost, double *stats, StageView stage) {
  // per-warp shared memory (dynamic: a block of 16 worlds needs more than the 48 KB static limit)
  extern __shared__ __align__(16) unsigned char s_raw[];
  WarpSmem<CPT> &S = reinterpret_cast<WarpSmem<CPT> *>(s_raw)[threadIdx.x >> 5];

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  // Worlds are handed out through an atomic counter: the grid holds exactly the warps that fit on
  // the GPU, each warp keeps pulling worlds until none are left, so there is no partial last wave.
  __shared__ int s_base;
  // Aggregate statistics (PgStats: the 8 doubles a multi-GPU run all-reduces) are taken in the launch's
  // epilogue, while a world's final state is still in shared memory: lane q < 8 of every warp carries
  // statistic q summed over the worlds the warp has stepped; when the block runs out of worlds its warps
  // add up in shared memory and 8 atomics per block reach the device vector.  No separate pass over HBM.
  __shared__ double s_stats[MAXW][8];
  double stat_acc = 0.0;
  auto flush_stats = [&]() {                        // on the way out (PG_SPH_SYNC > 0: all threads of the block together)
    if (!stats) return;
    if constexpr (PG_SPH_SYNC > 0) {
      if (lane < 8) s_stats[wib][lane] = stat_acc;
      __syncthreads();
      if (threadIdx.x < 8) {
        double x = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) x += s_stats[q][threadIdx.x];
        if (x != 0.0) atomicAdd(stats + threadIdx.x, x);
      }
    } else {
      if (lane < 8 && stat_acc != 0.0) atomicAdd(stats + lane, stat_acc);
    }
  };
  for (;;) {
  int w = 0;
  bool valid = true;
  if constexpr (PG_SPH_SYNC > 0) {
    // the block takes one world per warp at a time and its warps move through the step in phase (block
    // barriers below), so that an instruction line fetched once serves every warp of the block
    __syncthreads();
    if (threadIdx.x == 0) s_base = first_world + atomicAdd(work_counter, (int)(blockDim.x >> 5));
    __syncthreads();
    if (s_base >= end_world) { flush_stats(); return; }
    w = s_base + wib;
    valid = w < end_world;
    if (!valid) w = end_world - 1;
  } else {
    if (lane == 0) w = first_world + atomicAdd(work_counter, 1);
    w = __shfl_sync(kFull, w, 0);
    if (w >= end_world) { flush_stats(); return; }
  }
  if (order) w = order[w];                 // dearest worlds first (see k_order_worlds)
  const long long t_start = clock64();
  __syncwarp();
  const int n = valid ? v.counts[w * 3 + 1] : 0;
  const int ns = min(v.counts[w * 3 + 0], kMaxPlanes);
  const int n_blocks = (n + 31) >> 5;
  const float dt = pgm::clamp_dt(dt_in);
  const float restitution = v.restitution;
  const bool record = v.max_events > 0;
  const bool gate_ok = dt > 1.0e-9f;
  float4 *P4 = S.P4, *V4 = S.V4;
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;

  if (lane < ns) S.planes[lane] = v.planes[(size_t)w * v.cap[0] + lane];
  if (lane == 0) { S.n_list = 0; S.contacts = 0; }

  unsigned rest = 0;                       // bit s: this lane's body of actual slot s is at rest
  // bit s: that body is at rest AND pass 3 has already looked at exactly its present state and found nothing (no
  // contact, no event).  A body at rest is never integrated, so until a contact writes it the planes would be
  // visited with bit-identical inputs and give the bit-identical "nothing" again: the visits are skipped.  In a
  // thermalised world one body in six lies on the floor and used to pay the exact plane predicate every step.
  unsigned quiet = 0;
  const size_t base = (size_t)w * v.cap[1];
#pragma unroll 1
  for (int s = 0; s < CPT; s++) {
    const int i = s * 32 + lane;
    float4 a, b;
    if (i < n) {
      if (stage.load) {
        // the host's packed [body][3] floats straight from the staging buffer the upload landed in
        const float *sp = reinterpret_cast<const float *>(stage.pos + (size_t)w * stage.stride) + 3 * i;
        const float *sv = reinterpret_cast<const float *>(stage.vel + (size_t)w * stage.stride) + 3 * i;
        a = make_float4(sp[0], sp[1], sp[2], (stage.rest && stage.rest[(size_t)w * stage.rest_stride + i]) ? 1.0f : 0.0f);
        b = make_float4(sv[0], sv[1], sv[2], v.vel_rad[base + i].w);
      } else {
        a = v.pos_rest[base + i]; b = v.vel_rad[base + i];
      }
      if (a.w != 0.0f) rest |= 1u << s;
      a.w = reach(b.w, b.x, b.y, b.z, dt);
    } else {
      // padding bodies sit far away from everything (and from each other) with zero reach
      a = make_float4(kHuge + 1000.0f * (float)i, kHuge, kHuge, 0.0f);
      b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    P4[i] = a; V4[i] = b;
  }
  __syncwarp();

  float cx0 = 0.0f, cy0 = 0.0f, cz0 = 0.0f, trim = 3.0e38f;   // filter centre and trimming radius, carried from step to step

  for (int step = 0; step < n_steps; step++) {
    if constexpr (PG_SPH_SYNC >= 1) { if (step % PG_SPH_SYNC == 0) __syncthreads(); }
    // ---------------- pass 3: every sphere against the planes, in plane order (world.c:369-471)
    // cheap gate: |signed distance| beyond the reach -> the root test of the interval search fails
    auto plane_near = [&](const float4 a, const float4 pl) {
      const float s0 = __fmaf_rn(a.x, pl.x, __fmaf_rn(a.y, pl.y, a.z * pl.z)) - pl.w;
      return !(gate_ok && fabsf(s0) > a.w * 1.0001f + 1.0e-5f * (fabsf(s0) + fabsf(pl.w) + 1.0f));
    };
    // Every lane first lists its (body, plane) visits that pass the cheap gate -- 8 slots x 8 planes
    // at a time in one 64-bit word -- then all lanes work through their own lists TOGETHER: the
    // exact evaluations of different bodies run in the same SIMT pass instead of one after the
    // other.  Per body the planes stay in order, and a contact re-gates that body's remaining planes
    // against its new state.
    unsigned fired = 0;                    // bit s: a plane visit of that body fired in this step
#pragma unroll 1
    for (int g = 0; g < ns; g += 8) {
      const int g_end = min(g + 8, ns);
#pragma unroll 1
      for (int sg = 0; sg < n_blocks; sg += 8) {
        unsigned long long pend = 0;
#pragma unroll 1
        for (int s = sg; s < min(sg + 8, n_blocks); s++) {
          const int i = s * 32 + lane;
          if (i < n && !((quiet >> s) & 1u)) {
            const float4 a = P4[i];
#pragma unroll 1
            for (int k = g; k < g_end; k++) if (plane_near(a, S.planes[k])) pend |= 1ull << ((s - sg) * 8 + (k - g));
          }
        }
        while (__any_sync(kFull, pend != 0ull)) {
          if (pend) {
            const int bit = __ffsll((long long)pend) - 1;
            pend &= pend - 1ull;
            const int s = sg + (bit >> 3), k = g + (bit & 7), i = s * 32 + lane;
            int came_to_rest = 0;
            if (plane_slow_path(P4, V4, i, S.planes[k], restitution, dt, gate_ok, &came_to_rest, leaf)) {
              if (came_to_rest) rest |= 1u << s;
              fired |= 1u << s;
              if (record) emit_event(v, w, step, 3, PG_CLASS_DYNAMIC, i, PG_CLASS_STATIC, k);
              atomicAdd(&S.contacts, 1);
              const float4 a = P4[i];
#pragma unroll 1
              for (int k2 = k + 1; k2 < g_end; k2++) {
                const unsigned long long b2 = 1ull << ((bit & ~7) + (k2 - g));
                if (plane_near(a, S.planes[k2])) pend |= b2; else pend &= ~b2;
              }
            }
          }
        }
      }
    }
    quiet = (quiet | rest) & ~fired;
    // Centre the filter's coordinates are taken relative to.  It must sit among the bulk of the
    // bodies whatever a few stragglers do (a sphere that tunnelled through the floor keeps falling
    // for ever): the mean of the bodies within `trim` (twice the mean L1 distance, as of the
    // previous step) of the previous centre.  One pass over the world per step.
    auto warp_sum = [&](float x) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
      return x;
    };
    bool wild = false;
#pragma unroll 1
    for (int pass = (trim > 1.0e38f) ? 0 : 1; pass < 2; pass++) {      // a fresh start takes two passes
    float ax = 0.0f, ay = 0.0f, az = 0.0f, cntf = 0.0f, ad = 0.0f, dmax = 0.0f;
#pragma unroll 1
    for (int s = 0; s < n_blocks; s++) {
      const int i = s * 32 + lane;
      if (i < n) {
        const float4 a = P4[i];
        const float d = fabsf(a.x - cx0) + fabsf(a.y - cy0) + fabsf(a.z - cz0);
        dmax = fmaxf(dmax, d); if (!(d == d)) dmax = 3.0e38f;
        if (d <= trim) { ax += a.x; ay += a.y; az += a.z; cntf += 1.0f; ad += d; }
      }
    }
    wild = !(__uint_as_float(__reduce_max_sync(kFull, __float_as_uint(dmax))) < 1.0e15f);   // non-finite or astronomically spread: the expanded filter is not trusted
    cntf = warp_sum(cntf);
    if (cntf >= 1.0f) {
      cx0 = warp_sum(ax) / cntf; cy0 = warp_sum(ay) / cntf; cz0 = warp_sum(az) / cntf;
      trim = 2.0f * warp_sum(ad) / cntf + 2.0f;              // + room for the shift of the centre itself
    } else {
      trim = 3.0e38f;                                        // lost the bulk: start over from the plain mean
    }
    }
    if (!(fabsf(cx0) < 1.0e30f && fabsf(cy0) < 1.0e30f && fabsf(cz0) < 1.0e30f)) { cx0 = cy0 = cz0 = 0.0f; }
    __syncwarp();

    {
    // ---------------- pass 4, batched filter: rows in order (world.c:473-554)
    // Per block of 32 rows: (1) the reach filter of all 32 rows against all columns in one
    // branch-free loop (bit rl of W[k] = row rl x this lane's column of static slot k), using the
    // state every row WILL see if no contact intervenes (columns of the block's own slot: not yet
    // integrated for earlier rows, speculatively integrated for later ones); (2) only rows with a
    // candidate are visited, in order; (3) a contact rewrites the bits of the bodies it changed.
    int done = 0;                                  // rows [0, done) are integrated in shared memory
#pragma unroll
    for (int sl = 0; sl < CPT; sl++) {
      const unsigned b = __ballot_sync(kFull, (rest >> sl) & 1u);
      if (lane == 0) S.rest[sl] = b;
    }
#pragma unroll 1
    for (int rb = 0; rb < n_blocks; rb++) {
      const int row0 = rb * 32;
      const int rows_here = min(32, n - row0);
      constexpr int NP = (CPT + 2) / 2;            // column pairs: CPT slots + the integrated own column (+ padding)
      unsigned long long cx2[NP], cy2[NP], cz2[NP], cr2[NP], cc2[NP];
      {
        float fx[2 * NP], fy[2 * NP], fz[2 * NP], fr[2 * NP], fc[2 * NP];
        auto put = [&](int m, float px, float py, float pz, float rho) {
          fx[m] = px - cx0; fy[m] = py - cy0; fz[m] = pz - cz0; fr[m] = rho;
          const float n2c = __fmaf_rn(fx[m], fx[m], __fmaf_rn(fy[m], fy[m], fz[m] * fz[m]));
          fc[m] = __fmaf_rn(n2c, -kFilterSlack, n2c) - rho * rho;           // |q|^2 (1 - slack) - rho^2: the column's share of the slack
        };
#pragma unroll
        for (int k = 0; k < CPT; k++) {
          int sa = rb + k; if (sa >= CPT) sa -= CPT;
          const float4 a = P4[sa * 32 + lane];
          put(k, a.x, a.y, a.z, a.w);
        }
        // own body: filter coefficients of its row, and the state later rows of the block will see
        const float4 a = P4[row0 + lane], b = V4[row0 + lane];
        {
          const float rx = a.x - cx0, ry = a.y - cy0, rz = a.z - cz0;
          const float n2 = __fmaf_rn(rx, rx, __fmaf_rn(ry, ry, rz * rz));
          const float ea = -2.0f * rx, eb = -2.0f * ry, ec = -2.0f * rz, ed = -2.0f * a.w;
          const float ne = -(__fmaf_rn(a.w, a.w, -n2) + (kFilterSlack * n2 + 1.0e-6f));   // the row's share of the slack
          S.RA[2 * lane] = make_ulonglong2(pack2(ea, ea), pack2(eb, eb));
          S.RA[2 * lane + 1] = make_ulonglong2(pack2(ec, ec), pack2(ed, ed));
          S.RB[lane] = pack2(ne, ne);
        }
        V3 p = v3(a.x, a.y, a.z), vel = v3(b.x, b.y, b.z);
        float irho = a.w;
        if (!((rest >> rb) & 1u)) { pgm::integrate(p, vel, dt); irho = reach(b.w, vel.x, vel.y, vel.z, dt); }
        put(CPT, p.x, p.y, p.z, irho);
#pragma unroll
        for (int m = CPT + 1; m < 2 * NP; m++) put(m, p.x, p.y, p.z, irho);
#pragma unroll
        for (int q = 0; q < NP; q++) {
          cx2[q] = pack2(fx[2 * q], fx[2 * q + 1]); cy2[q] = pack2(fy[2 * q], fy[2 * q + 1]); cz2[q] = pack2(fz[2 * q], fz[2 * q + 1]);
          cr2[q] = pack2(fr[2 * q], fr[2 * q + 1]); cc2[q] = pack2(fc[2 * q], fc[2 * q + 1]);
        }
      }
      __syncwarp();
      unsigned W[2 * NP];
#pragma unroll
      for (int m = 0; m < 2 * NP; m++) W[m] = 0;
#pragma unroll 2
      for (int rl = 0; rl < rows_here; rl++) {
        const ulonglong2 e0 = S.RA[2 * rl], e1 = S.RA[2 * rl + 1];
        const unsigned long long ne = S.RB[rl];
#pragma unroll
        for (int q = 0; q < NP; q++) {
          // (|q_c|^2 - rho_c^2) - 2 rho_r rho_c - 2 q_r.q_c - (rho_r^2 - |q_r|^2 + slack): negative <=> candidate
          unsigned long long x = fadd2(cc2[q], ne);
          x = ffma2(e1.y, cr2[q], x); x = ffma2(e1.x, cz2[q], x); x = ffma2(e0.y, cy2[q], x); x = ffma2(e0.x, cx2[q], x);
          float lo, hi; unpack2(x, lo, hi);
          W[2 * q] = __funnelshift_l(__float_as_uint(lo), W[2 * q], 1);        // shift the sign bit in
          W[2 * q + 1] = __funnelshift_l(__float_as_uint(hi), W[2 * q + 1], 1);
        }
      }
#pragma unroll
      for (int m = 0; m <= CPT; m++) W[m] = __brev(W[m]) >> (32 - rows_here);   // bit rl = row rl
      // own slot: rows before this lane's body see it as it is, rows after it see it integrated
      W[0] = (W[0] & ((1u << lane) - 1u)) | (W[CPT] & (0xfffffffeu << lane));
      if (wild) {                                  // every pair is a candidate; the rows filter for themselves
#pragma unroll
        for (int k = 0; k < CPT; k++) W[k] = 0xffffffffu >> (32 - rows_here);
        W[0] &= ~(1u << lane);                     // ... except a body with itself
      }
      unsigned *WS = S.W;                     // WS[k * 32 + lane]: the bit matrix, in shared memory from here on
#pragma unroll
      for (int k = 0; k < CPT; k++) { WS[k * 32 + lane] = W[k]; S.W0[k * 32 + lane] = W[k]; }
      __syncwarp();
      // --- closest-approach gate of every surviving (row, column) bit, one pair per lane: the bits
      // are compacted into a list (k, lane, rl), kGateList at a time, and gated 32 at a time.  With a
      // time step too small for the gate to mean anything the rows filter for themselves (`fresh`).
      bool gated = false;
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < CPT; k++) cnt += __popc(W[k]);
      // list positions are handed out by a shared-memory counter (the order of the list is irrelevant)
      int first = 0;
      if (cnt) first = atomicAdd(&S.n_list, cnt);
      __syncwarp();
      const int total = S.n_list;
      __syncwarp();
      if (lane == 0) S.n_list = 0;
      if (dt > 1.0e-9f) {
        gated = true;
        unsigned *list = S.list;
        // windows of kGateList bits (one, unless the block sits in a pile)
#pragma unroll 1
        for (int base = 0; base < total; base += kGateList) {
          if (cnt) {
            int pos = first - base;
#pragma unroll 1
            for (int k = 0; k < CPT && pos < kGateList; k++) {
              for (unsigned word = S.W0[k * 32 + lane]; word; word &= word - 1u, pos++)
                if (pos >= 0 && pos < kGateList) list[pos] = (unsigned)k | ((unsigned)lane << 8) | ((unsigned)(__ffs((int)word) - 1) << 16);
            }
          }
          __syncwarp();
          const int m = min(kGateList, total - base);
#pragma unroll 1
          for (int idx = lane; idx < m; idx += 32) {
            const unsigned e = list[idx];
            const int ek = e & 0xff, el = (e >> 8) & 0xff, erl = e >> 16;
            int sa = rb + ek; if (sa >= CPT) sa -= CPT;
            // a column of the block's own slot that comes before the row is integrated by then
            const bool integ = ek == 0 && el < erl && !((S.rest[sa] >> el) & 1u);
            if (pgs::pair_gate_far(P4, V4, row0 + erl, sa * 32 + el, integ, dt)) atomicAnd(&WS[ek * 32 + el], ~(1u << erl));
          }
          __syncwarp();
        }
      }
      unsigned rows_any = 0;
      if (total > 0) {
        unsigned any = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) any |= WS[k * 32 + lane];
        rows_any = __reduce_or_sync(kFull, any);
      }
      unsigned dirty_rows = 0;
      auto commit_to = [&](int upto) {             // integrate rows [done, upto) of this block (world.c:549-553)
        const int r = row0 + lane;
        if (r >= done && r < upto && !((rest >> rb) & 1u)) {
          const float4 a = P4[r], b = V4[r];
          V3 p = v3(a.x, a.y, a.z), vel = v3(b.x, b.y, b.z);
          pgm::integrate(p, vel, dt);
          P4[r] = make_float4(p.x, p.y, p.z, reach(b.w, vel.x, vel.y, vel.z, dt));
          V4[r] = make_float4(vel.x, vel.y, vel.z, b.w);
        }
        done = upto;
        __syncwarp();
      };
      const unsigned full = (CPT == 32) ? 0xffffffffu : ((1u << CPT) - 1u);
      while (rows_any) {
        const int rl = __ffs((int)rows_any) - 1;
        const int i = row0 + rl;
        commit_to(i);
        unsigned mask = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) mask |= ((WS[k * 32 + lane] >> rl) & 1u) << k;
        // static slot k is actual slot (rb + k) mod CPT: rotate the bits into column order
        const unsigned amask = ((mask << rb) | (mask >> ((CPT - rb) & 31))) & full;
        const int n_chg = row_visits(P4, V4, i, rb == 0 ? mask : amask, !gated || ((dirty_rows >> rl) & 1u), lane, CPT, dt, v, w, step, record, S.chg, &S.contacts);
        const unsigned later = 0xfffffffeu << rl;  // rows after rl
        if (n_chg) {
          if (n_chg > kMaxRowChanges) {
            dirty_rows |= later;                   // too many to track: every later row re-filters itself
            quiet = 0;                             // ... and every body meets the planes again
          } else {
#pragma unroll 1
            for (int e = 0; e <= n_chg; e++) {
              const int c = e < n_chg ? S.chg[e] : i;       // the changed columns, then the row body itself
              const bool c_rest = (S.rest[c >> 5] >> (c & 31)) & 1u;
              const unsigned word = recheck_column(P4, V4, row0, rl, c, c_rest, done, n, dt, lane);
              int kc = (c >> 5) - rb; if (kc < 0) kc += CPT;
              if (lane == 0) WS[kc * 32 + (c & 31)] = (WS[kc * 32 + (c & 31)] & ~later) | word;
              if (c > i && c < row0 + 32) dirty_rows |= 1u << (c - row0);   // a later row of this block changed
              if ((c & 31) == lane) quiet &= ~(1u << (c >> 5));             // a written body meets the planes again
            }
            __syncwarp();
          }
          unsigned any = 0;
#pragma unroll
          for (int k = 0; k < CPT; k++) any |= WS[k * 32 + lane];
          rows_any = (__reduce_or_sync(kFull, any) | dirty_rows) & later;
        } else {
          rows_any &= rows_any - 1u;
        }
      }
      commit_to(row0 + rows_here);
    }
    }
  }

  __syncwarp();
#pragma unroll 1
  for (int s = 0; s < n_blocks; s++) {
    const int i = s * 32 + lane;
    if (i < n) {
      const float4 a = P4[i], b = V4[i];
      v.pos_rest[base + i] = make_float4(a.x, a.y, a.z, ((rest >> s) & 1u) ? 1.0f : 0.0f);
      v.vel_rad[base + i] = b;
      if (stage.store) {
        float *sp = reinterpret_cast<float *>(stage.pos + (size_t)w * stage.stride) + 3 * i;
        float *sv = reinterpret_cast<float *>(stage.vel + (size_t)w * stage.stride) + 3 * i;
        sp[0] = a.x; sp[1] = a.y; sp[2] = a.z;
        sv[0] = b.x; sv[1] = b.y; sv[2] = b.z;
        if (stage.rest) stage.rest[(size_t)w * stage.rest_stride + i] = ((rest >> s) & 1u) ? 1 : 0;
      }
    }
  }
  if (cost && valid && lane == 0) cost[w] = (unsigned)min((clock64() - t_start) >> 10, 0x7fffffffll);
  if (stats && valid) {
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int s = 0; s < n_blocks; s++) {
      const int i = s * 32 + lane;
      if (i < n) {
        const float4 p = P4[i], b = V4[i];
        a[0] += 1.0;
        a[1] += 0.5 * ((double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z);
        a[2] += b.x; a[3] += b.y; a[4] += b.z;
        a[6] += ((rest >> s) & 1u) ? 1.0 : 0.0;
        const bool fin = isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && isfinite(b.x) && isfinite(b.y) && isfinite(b.z);
        a[7] += fin ? 0.0 : 1.0;
      }
    }
    if (lane == 0) a[5] = (double)S.contacts;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      double x = a[q];
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
      if (lane == q) stat_acc += x;
    }
  }
  __syncwarp();
  }   // next world
}

// ---------------------------------------------------------------- aggregate statistics (any-collider worlds)
// Sphere batches take their statistics in k_step_spheres' epilogue; game-sized general worlds use this
// one pass over the records: block partials -> 8 double atomics.
__global__ void k_stats_general(WorldsView v, double *out) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long total = (long long)v.n_worlds * v.cap[1];
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(k / v.cap[1]), i = (int)(k % v.cap[1]);
    if (i >= v.counts[w * 3 + 1]) continue;
    const PgBody &b = v.bodies[1][k];
    acc[0] += 1.0;
    acc[1] += 0.5 * ((double)b.velocity[0] * b.velocity[0] + (double)b.velocity[1] * b.velocity[1] + (double)b.velocity[2] * b.velocity[2]);
    acc[2] += b.velocity[0]; acc[3] += b.velocity[1]; acc[4] += b.velocity[2];
    acc[6] += b.at_rest ? 1.0 : 0.0;
    bool fin = true;
    for (int q = 0; q < 3; q++) fin = fin && isfinite(b.position[q]) && isfinite(b.velocity[q]);
    acc[7] += fin ? 0.0 : 1.0;
  }
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < v.n_worlds; w += (long long)gridDim.x * blockDim.x)
    acc[5] += (double)v.n_events[w];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    double x = acc[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(out + q, x);
  }
}

// Packed [body][3] float arrays (what a host caller holds) <-> float4 SoA hot state.
__global__ void k_pack_state(WorldsView v, int first, int count, int n, const float *pos, const float *vel,
                             const unsigned char *rest) {
  const long long total = (long long)count * n;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int wd = (int)(k / n), i = (int)(k % n);
    const size_t o = (size_t)(first + wd) * v.cap[1] + i;
    v.pos_rest[o] = make_float4(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2], (rest && rest[k]) ? 1.0f : 0.0f);
    const float r = v.vel_rad[o].w;
    v.vel_rad[o] = make_float4(vel[3 * k], vel[3 * k + 1], vel[3 * k + 2], r);
  }
}
__global__ void k_unpack_state(WorldsView v, int first, int count, int n, float *pos, float *vel, unsigned char *rest) {
  const long long total = (long long)count * n;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int wd = (int)(k / n), i = (int)(k % n);
    const size_t o = (size_t)(first + wd) * v.cap[1] + i;
    const float4 a = v.pos_rest[o], b = v.vel_rad[o];
    pos[3 * k] = a.x; pos[3 * k + 1] = a.y; pos[3 * k + 2] = a.z;
    vel[3 * k] = b.x; vel[3 * k + 1] = b.y; vel[3 * k + 2] = b.z;
    if (rest) rest[k] = a.w != 0.0f ? 1 : 0;
  }
}

}  // namespace

int pg_launch_pack_state_on(PgWorlds *w, cudaStream_t st, int first, int count, int n, const float *pos, const float *vel, const unsigned char *rest) {
  k_pack_state<<<w->sm_count * 2, 256, 0, st>>>(w->v, first, count, n, pos, vel, rest);
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}
int pg_launch_unpack_state_on(PgWorlds *w, cudaStream_t st, int first, int count, int n, float *pos, float *vel, unsigned char *rest) {
  k_unpack_state<<<w->sm_count * 2, 256, 0, st>>>(w->v, first, count, n, pos, vel, rest);
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}
int pg_launch_pack_state(PgWorlds *w, int first, int count, int n, const float *pos, const float *vel, const unsigned char *rest) {
  k_pack_state<<<w->sm_count * 4, 256, 0, w->stream>>>(w->v, first, count, n, pos, vel, rest);
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}
int pg_launch_unpack_state(PgWorlds *w, int first, int count, int n, float *pos, float *vel, unsigned char *rest) {
  k_unpack_state<<<w->sm_count * 4, 256, 0, w->stream>>>(w->v, first, count, n, pos, vel, rest);
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}

// Hand-out order for the next launch: worlds sorted by what they cost in the last one, dearest first
// (counting sort on a log-scale key, one block).  A launch ends when its slowest world ends, so the
// slow ones must start first; and a block keeps its worlds in phase with barriers, so worlds of
// similar cost belong in the same block.  Worlds are independent: the order never changes a result.
__global__ void k_order_worlds(const unsigned *cost, int n_worlds, int *order) {
  __shared__ int hist[256], cursor[256];
  auto key = [](unsigned c) {               // 5 exponent bits + 3 mantissa bits, larger cost -> smaller key
    if (c < 8u) return 255 - (int)c;
    const int e = 31 - __clz(c);
    return 255 - min(255, ((e - 2) << 3) | (int)((c >> (e - 3)) & 7u));
  };
  for (int k = threadIdx.x; k < 256; k += blockDim.x) hist[k] = 0;
  __syncthreads();
  for (int w = threadIdx.x; w < n_worlds; w += blockDim.x) atomicAdd(&hist[key(cost[w])], 1);
  __syncthreads();
  if (threadIdx.x == 0) { int acc = 0; for (int k = 0; k < 256; k++) { cursor[k] = acc; acc += hist[k]; } }
  __syncthreads();
  for (int w = threadIdx.x; w < n_worlds; w += blockDim.x) order[atomicAdd(&cursor[key(cost[w])], 1)] = w;
}

template <int CPT, int MAXW>
static int launch_spheres(PgWorlds *w, float dt, int n_steps, int first, int count, cudaStream_t stream, int *counter, double *stats,
                          const StageView &stage) {
  static bool configured[64] = {false};      // function attributes are per device
  constexpr size_t warp_bytes = sizeof(WarpSmem<CPT>);
  const int dev = w->device >= 0 && w->device < 64 ? w->device : 0;
  if (!configured[dev] || w->device >= 64) {
    cudaFuncSetAttribute(k_step_spheres<CPT, MAXW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k_step_spheres<CPT, MAXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(warp_bytes * MAXW));
    configured[dev] = true;
  }
  if (count <= 0) return PG_OK;              // configuration only (callers about to capture a graph)
  // One block per SM.  The fewest waves of (SMs x MAXW) worlds that hold the batch, then the fewest
  // warps per block that still do it in that many waves: 4096 worlds on 148 SMs run as 2 waves of
  // 148 x 14 instead of 1.73 waves of 148 x 16.
  const int sms = w->sm_count;
  const int waves = (count + sms * MAXW - 1) / (sms * MAXW);
  const int wpb = std::max(1, std::min(MAXW, (count + waves * sms - 1) / (waves * sms)));
  const int blocks = std::min((count + wpb - 1) / wpb, sms);
  // cost-ordered hand-out only for launches over the whole batch (the chunked host pipeline keeps index order)
  const bool whole = first == 0 && count == w->v.n_worlds && count > sms * wpb;
  const int *order = nullptr;
  if (whole && w->cost_valid) {
    k_order_worlds<<<1, 1024, 0, stream>>>(w->d_cost, count, w->d_order);
    w->launches++;
    order = w->d_order;
  }
  cudaMemsetAsync(counter, 0, sizeof(int), stream);
  k_step_spheres<CPT, MAXW><<<blocks, wpb * 32, warp_bytes * wpb, stream>>>(w->v, dt, n_steps, counter, first, first + count, order,
                                                                    (whole && n_steps > 0) ? w->d_cost : nullptr, stats, stage);
  if (n_steps > 0) w->cost_valid = whole;
  return PG_OK;
}

// Step worlds [first, first+count) on `stream` (the whole batch on the handle's stream by default).
int pg_launch_sphere_range(PgWorlds *w, float dt, int n_steps, int first, int count, cudaStream_t stream, int *counter, double *stats,
                           const StageView *stage_or_null) {
  const int cap = w->v.cap[1];
  StageView stage;
  memset(&stage, 0, sizeof(stage));
  if (stage_or_null) stage = *stage_or_null;
  if (w->v.cap[0] > kMaxPlanes) { pg_set_error("sphere mode supports at most %d static planes", kMaxPlanes); return PG_ERR_CAPACITY; }
  if (cap <= 32) launch_spheres<1, PG_SPH_WARPS>(w, dt, n_steps, first, count, stream, counter, stats, stage);
  else if (cap <= 64) launch_spheres<2, PG_SPH_WARPS>(w, dt, n_steps, first, count, stream, counter, stats, stage);
  else if (cap <= 128) launch_spheres<4, PG_SPH_WARPS>(w, dt, n_steps, first, count, stream, counter, stats, stage);
  else if (cap <= 256) launch_spheres<8, PG_SPH_WARPS>(w, dt, n_steps, first, count, stream, counter, stats, stage);
  else if (cap <= 512) launch_spheres<16, 8>(w, dt, n_steps, first, count, stream, counter, stats, stage);
  else { pg_set_error("sphere mode supports at most 512 spheres per world (got capacity %d)", cap); return PG_ERR_CAPACITY; }
  PG_CUDA(cudaGetLastError());
  if (count > 0) w->launches++;
  return PG_OK;
}

// The whole batch on the handle's stream; the launch's epilogue leaves the aggregate statistics in d_stats.
// n_steps == 0 is the statistics-only pass (state in, state out unchanged) for a state that was uploaded,
// not stepped.
int pg_launch_sphere_step(PgWorlds *w, float dt, int n_steps) {
  PG_CUDA(cudaMemsetAsync(w->d_stats, 0, 8 * sizeof(double), w->stream));
  int rc = pg_launch_sphere_range(w, dt, n_steps, 0, w->v.n_worlds, w->stream, w->d_work_counter, w->d_stats);
  w->stats_valid = rc == PG_OK;
  return rc;
}

// Make d_stats describe the current state (on the handle's stream, no host sync).
int pg_launch_worlds_stats(PgWorlds *w) {
  if (w->mode == PG_MODE_SPHERES) {
    if (w->stats_valid) return PG_OK;              // the last step launch's epilogue already did it
    return pg_launch_sphere_step(w, 0.0f, 0);
  }
  PG_CUDA(cudaMemsetAsync(w->d_stats, 0, 8 * sizeof(double), w->stream));
  k_stats_general<<<w->sm_count * 4, 256, 0, w->stream>>>(w->v, w->d_stats);
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}

#ifdef PG_WALK_STATS
// debug build only: read (and optionally clear) this translation unit's interval-walk counters
extern "C" int pg_debug_walk_stats(unsigned long long *out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, pgm::pg_walk_stats, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(pgm::pg_walk_stats, z, sizeof z); }
  return 0;
}
#endif
