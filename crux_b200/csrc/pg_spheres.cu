// pg_spheres.cu -- batched sphere worlds (config C3: 4096 x 256-sphere bouncehouse worlds).
//
// ONE WARP PER WORLD, state in registers, reference solve order preserved exactly.
//
// Body i of a world lives in lane (i & 31), register slot (i >> 5); a world holds up to 32*CPT
// spheres.  The whole n_steps loop runs inside one launch: positions/velocities are read from HBM
// once and written once per launch (32 B + 32 B per body), everything in between is register and
// shuffle traffic.  There are no block barriers -- a world's sequential sweep only ever needs
// warp-level votes.
//
// Pass 4 of physics_step (world.c:473-554) is N rows; row i visits every column j != i, mutating
// i and j in place, then integrates i.  Per row every lane runs a CONSERVATIVE reach test on its CPT
// columns (float, |ci-cj|^2 against (rho_i+rho_j)^2 with rho = (r + |v| dt)(1+1e-4)): a pair that
// fails it cannot pass the root test of interval_collision (world.c:559-565), so skipping it is
// exact.  Survivors are taken in increasing j (warp min-reduce), gated by a second conservative
// test (closest approach of the two linear paths over [0,dt] against rs + delta), and only then
// evaluated with the exact predicate + narrowphase + resolver of pg_math.cuh, redundantly on all
// lanes (warp-uniform control flow).  After any evaluated candidate the remaining columns are
// re-filtered against the row body's current state, which is what the sequential loop would see.
//
// Pass 3 (dynamic x static planes, world.c:369-471) touches only the dynamic body, so each lane
// runs it for its own bodies, planes in array order.
#include "pg_common.cuh"
#include "pg_math.cuh"
#include "pg_sphere_exact.cuh"

#include <algorithm>

namespace {

using pgm::V3;
using pgm::v3;

#ifndef PG_SPH_WARPS
#define PG_SPH_WARPS 12   // resident warps (worlds) per SM the register budget is sized for (measured best: 8 -> 0.41, 12 -> 0.28, 16 -> 0.44 ms/step)
#endif
constexpr int kMaxPlanes = 16;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kHuge = 1.0e8f;      // padding bodies: far from everything, squares still finite

using pgs::Sphere;
using pgs::reach;
using pgs::sphere_pair_exact;
using pgs::sphere_plane_exact;

__device__ __forceinline__ void emit_event(const WorldsView &v, int w, int step, int pass, int ac, int ai, int bc, int bi) {
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
// Lanes additionally cache (px,py,pz,rho) of their own CPT bodies in registers for the hot filter.

// Slow half of one pass-4 row: candidates of row i (bit s of `mask` = this lane's column s*32+lane
// passed the reach filter) are evaluated in increasing column order with the exact predicate.
// Warp-uniform control flow; state is read from / written to shared memory.  Returns true if any
// body changed (callers then reload their register caches).
__device__ __noinline__ bool row_slow_path(float4 *P4, float4 *V4, int i, unsigned mask, int lane, int ncols_slots,
                                           float dt, const WorldsView &v, int w, int step, bool record) {
  __syncwarp();
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;
  bool dirty = false;
  int j0 = 0;
  // Lane-parallel refinement: every lane runs the closest-approach gate on its own candidate
  // columns, so the serial part below only sees pairs that really come within contact range.
  auto refine = [&](unsigned m, float4 ap, float4 av) {
    unsigned keep = 0;
    const float na2 = __fmaf_rn(av.x, av.x, __fmaf_rn(av.y, av.y, av.z * av.z));
    for (; m; m &= m - 1) {
      const int s = __ffs((int)m) - 1;
      const float4 bp = P4[s * 32 + lane], bv = V4[s * 32 + lane];
      const float nb2 = __fmaf_rn(bv.x, bv.x, __fmaf_rn(bv.y, bv.y, bv.z * bv.z));
      const float vsum = sqrtf(2.0f * (na2 + nb2)) * 1.0001f;            // >= |vA| + |vB|
      if (!pgs::pair_gate_rejects(v3(ap.x, ap.y, ap.z), v3(av.x, av.y, av.z), v3(bp.x, bp.y, bp.z), v3(bv.x, bv.y, bv.z),
                                  av.w + bv.w, vsum, dt)) keep |= 1u << s;
    }
    return keep;
  };
  if (dt > 1.0e-9f) mask = refine(mask, P4[i], V4[i]);
  for (;;) {
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
    if (sphere_pair_exact(A, B, dt, leaf)) {
      const float rhoA = reach(A.r, A.vx, A.vy, A.vz, dt), rhoB = reach(B.r, B.vx, B.vy, B.vz, dt);
      __syncwarp();
      if (lane == 0) {
        P4[i] = make_float4(A.px, A.py, A.pz, rhoA); V4[i] = make_float4(A.vx, A.vy, A.vz, A.r);
        P4[j] = make_float4(B.px, B.py, B.pz, rhoB); V4[j] = make_float4(B.vx, B.vy, B.vz, B.r);
        if (record) emit_event(v, w, step, 4, PG_CLASS_DYNAMIC, i, PG_CLASS_DYNAMIC, j);
      }
      __syncwarp();
      dirty = true;
      // the row body changed: re-filter this lane's columns against its new state
      mask = 0;
      for (int s = 0; s < ncols_slots; s++) {
        const int c = s * 32 + lane;
        const float4 q = P4[c];
        const float dx = A.px - q.x, dy = A.py - q.y, dz = A.pz - q.z;
        const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz));
        const float T = rhoA + q.w;
        if (d2 <= T * T && c != i) mask |= 1u << s;
      }
      if (dt > 1.0e-9f) mask = refine(mask, P4[i], V4[i]);
    }
    j0 = j + 1;
  }
  return dirty;
}

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

// The register cache is kept ROTATED: while row block `rb` (rows rb*32 .. rb*32+31) is swept,
// static slot k of the cache holds the body of actual slot (rb + k) mod CPT, so the block's own
// bodies always sit in static slot 0.  That keeps every register index a compile-time constant
// with ONE copy of the row loop in the instruction stream (a fully unrolled variant was 9.5k SASS
// instructions and spent most of its time in instruction-cache misses).
template <int CPT, int WPB>
__global__ void __launch_bounds__(WPB * 32, (CPT <= 8) ? PG_SPH_WARPS / WPB : 2)
k_step_spheres(WorldsView v, float dt_in, int n_steps, int *work_counter, int first_world, int end_world) {
  __shared__ float4 s_planes[WPB][kMaxPlanes];
  __shared__ float4 s_P4[WPB][CPT * 32];
  __shared__ float4 s_V4[WPB][CPT * 32];

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  // Worlds are handed out through an atomic counter: the grid holds exactly the warps that fit on
  // the GPU, each warp keeps pulling worlds until none are left, so there is no partial last wave.
  for (;;) {
  int w = 0;
  if (lane == 0) w = first_world + atomicAdd(work_counter, 1);
  w = __shfl_sync(kFull, w, 0);
  if (w >= end_world) return;
  __syncwarp();
  const int n = v.counts[w * 3 + 1];
  const int ns = min(v.counts[w * 3 + 0], kMaxPlanes);
  const int n_blocks = (n + 31) >> 5;
  const float dt = pgm::clamp_dt(dt_in);
  const float restitution = v.restitution;
  const bool record = v.max_events > 0;
  const bool gate_ok = dt > 1.0e-9f;
  float4 *P4 = s_P4[wib], *V4 = s_V4[wib];
  pgm::LeafLen leaf; leaf.lo = v.leaf_lo; leaf.hi = v.leaf_hi;

  if (lane < ns) s_planes[wib][lane] = v.planes[(size_t)w * v.cap[0] + lane];

  unsigned rest = 0;                       // bit s: this lane's body of actual slot s is at rest
  const size_t base = (size_t)w * v.cap[1];
#pragma unroll 1
  for (int s = 0; s < CPT; s++) {
    const int i = s * 32 + lane;
    float4 a, b;
    if (i < n) {
      a = v.pos_rest[base + i]; b = v.vel_rad[base + i];
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

  // Register cache of this lane's CPT bodies for the reach filter, in coordinates relative to the
  // world's centre: q = p - centre, qr = rho, qc = |q|^2 - rho^2.  With them
  //   |ci - ck|^2 <= (rho_i + rho_k)^2   <=>   qc_k - 2 q_i.q_k - 2 rho_i rho_k <= rho_i^2 - |q_i|^2
  // is four FMAs and a compare per pair; the rounding of the expanded form is covered by `slack`
  // (>= 60 eps of the largest term), which keeps the filter a superset of the exact root test.
  float qx[CPT], qy[CPT], qz[CPT], qr[CPT], qc[CPT];
  float cx0 = 0.0f, cy0 = 0.0f, cz0 = 0.0f;
  auto cache_slot = [&](int k, const float4 a) {
    qx[k] = a.x - cx0; qy[k] = a.y - cy0; qz[k] = a.z - cz0; qr[k] = a.w;
    qc[k] = __fmaf_rn(qx[k], qx[k], __fmaf_rn(qy[k], qy[k], qz[k] * qz[k])) - a.w * a.w;
  };

  for (int step = 0; step < n_steps; step++) {
    // ---------------- pass 3: every sphere against the planes, in plane order (world.c:369-471)
    float bl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bh[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    // cheap gate: |signed distance| beyond the reach -> the root test of the interval search fails
    auto plane_near = [&](const float4 a, const float4 pl) {
      const float s0 = __fmaf_rn(a.x, pl.x, __fmaf_rn(a.y, pl.y, a.z * pl.z)) - pl.w;
      return !(gate_ok && fabsf(s0) > a.w * 1.0001f + 1.0e-5f * (fabsf(s0) + fabsf(pl.w) + 1.0f));
    };
    if (ns <= 8 && n_blocks <= 8) {
      // Every lane first lists its (body, plane) visits that pass the cheap gate, then all lanes
      // work through their own lists TOGETHER: the exact evaluations of different bodies run in
      // the same SIMT pass instead of one after the other.  Per body the planes stay in order, and
      // a contact re-gates that body's remaining planes against its new state.
      unsigned long long pend = 0;
#pragma unroll 1
      for (int s = 0; s < n_blocks; s++) {
        const int i = s * 32 + lane;
        if (i < n) {
          const float4 a = P4[i];
          for (int k = 0; k < ns; k++) if (plane_near(a, s_planes[wib][k])) pend |= 1ull << (s * 8 + k);
        }
      }
      while (__any_sync(kFull, pend != 0ull)) {
        if (pend) {
          const int bit = __ffsll((long long)pend) - 1;
          pend &= pend - 1ull;
          const int s = bit >> 3, k = bit & 7, i = s * 32 + lane;
          int came_to_rest = 0;
          if (plane_slow_path(P4, V4, i, s_planes[wib][k], restitution, dt, gate_ok, &came_to_rest, leaf)) {
            if (came_to_rest) rest |= 1u << s;
            if (record) emit_event(v, w, step, 3, PG_CLASS_DYNAMIC, i, PG_CLASS_STATIC, k);
            const float4 a = P4[i];
            for (int k2 = k + 1; k2 < ns; k2++) {
              const unsigned long long b2 = 1ull << (s * 8 + k2);
              if (plane_near(a, s_planes[wib][k2])) pend |= b2; else pend &= ~b2;
            }
          }
        }
      }
    } else {
#pragma unroll 1
      for (int s = 0; s < n_blocks; s++) {
        const int i = s * 32 + lane;
        if (i < n) {
          float4 a = P4[i];
#pragma unroll 1
          for (int k = 0; k < ns; k++) {
            const float4 pl = s_planes[wib][k];
            if (!plane_near(a, pl)) continue;
            int came_to_rest = 0;
            if (plane_slow_path(P4, V4, i, pl, restitution, dt, gate_ok, &came_to_rest, leaf)) {
              a = P4[i];
              if (came_to_rest) rest |= 1u << s;
              if (record) emit_event(v, w, step, 3, PG_CLASS_DYNAMIC, i, PG_CLASS_STATIC, k);
            }
          }
        }
      }
    }
#pragma unroll 1
    for (int s = 0; s < n_blocks; s++) {
      const int i = s * 32 + lane;
      if (i < n) {
        const float4 a = P4[i];
        bl[0] = fminf(bl[0], a.x); bl[1] = fminf(bl[1], a.y); bl[2] = fminf(bl[2], a.z);
        bh[0] = fmaxf(bh[0], a.x); bh[1] = fmaxf(bh[1], a.y); bh[2] = fmaxf(bh[2], a.z);
      }
    }
    // world centre and the magnitude the filter's rounding slack is scaled with
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        bl[k] = fminf(bl[k], __shfl_xor_sync(kFull, bl[k], o));
        bh[k] = fmaxf(bh[k], __shfl_xor_sync(kFull, bh[k], o));
      }
    }
    cx0 = 0.5f * (bl[0] + bh[0]); cy0 = 0.5f * (bl[1] + bh[1]); cz0 = 0.5f * (bl[2] + bh[2]);
    if (!(fabsf(cx0) < 1.0e30f && fabsf(cy0) < 1.0e30f && fabsf(cz0) < 1.0e30f)) { cx0 = cy0 = cz0 = 0.0f; }
    const float hx = bh[0] - cx0 + 2.0f, hy = bh[1] - cy0 + 2.0f, hz = bh[2] - cz0 + 2.0f;   // + room to move within the step
    const float maxn2 = hx * hx + hy * hy + hz * hz;
    __syncwarp();

    // ---------------- pass 4: rows in order (world.c:473-554)
#pragma unroll
    for (int k = 0; k < CPT; k++) cache_slot(k, P4[k * 32 + lane]);

#pragma unroll 1
    for (int rb = 0; rb < n_blocks; rb++) {
      // speculative integration of this lane's row body (committed to the register cache when its
      // row completes, written back to shared memory lazily; recomputed if a contact intervenes)
      float ipx, ipy, ipz, ivx, ivy, ivz, irho, irad;
      auto speculate = [&]() {
        const float4 a = P4[rb * 32 + lane], b = V4[rb * 32 + lane];
        V3 p = v3(a.x, a.y, a.z), vel = v3(b.x, b.y, b.z);
        irho = a.w; irad = b.w;
        if (!((rest >> rb) & 1u)) { pgm::integrate(p, vel, dt); irho = reach(b.w, vel.x, vel.y, vel.z, dt); }
        ipx = p.x; ipy = p.y; ipz = p.z; ivx = vel.x; ivy = vel.y; ivz = vel.z;
      };
      speculate();
      const int rows_here = min(32, n - rb * 32);
      int flushed = 0;                              // rows [0, flushed) of this block are committed in shared memory
#pragma unroll 1
      for (int rl = 0; rl < rows_here; rl++) {
        const int i = rb * 32 + rl;
        const float4 rp = P4[i];
        const float rx = rp.x - cx0, ry = rp.y - cy0, rz = rp.z - cz0;
        const float n2 = __fmaf_rn(rx, rx, __fmaf_rn(ry, ry, rz * rz));
        const float ea = -2.0f * rx, eb = -2.0f * ry, ec = -2.0f * rz, ed = -2.0f * rp.w;
        const float ee = __fmaf_rn(rp.w, rp.w, -n2) + (8.0e-6f * (n2 + maxn2) + 1.0e-6f);
        unsigned mask = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
          const float t = __fmaf_rn(ea, qx[k], __fmaf_rn(eb, qy[k], __fmaf_rn(ec, qz[k], __fmaf_rn(ed, qr[k], qc[k]))));
          mask |= (t <= ee ? 1u : 0u) << k;
        }
        if (lane == rl) mask &= ~1u;                 // the row body itself lives in static slot 0
        if (__any_sync(kFull, mask != 0)) {
          // bring shared memory up to date with the rows committed so far in this block
          if (lane >= flushed && lane < rl) {
            P4[rb * 32 + lane] = make_float4(ipx, ipy, ipz, irho);
            V4[rb * 32 + lane] = make_float4(ivx, ivy, ivz, irad);
          }
          flushed = rl;
          // static slot k is actual slot (rb + k) mod CPT: rotate the bits into column order
          const unsigned full = (CPT == 32) ? 0xffffffffu : ((1u << CPT) - 1u);
          const unsigned amask = ((mask << rb) | (mask >> ((CPT - rb) & 31))) & full;
          if (row_slow_path(P4, V4, i, rb == 0 ? mask : amask, lane, CPT, dt, v, w, step, record)) {
#pragma unroll
            for (int k = 0; k < CPT; k++) {
              int sa = rb + k; if (sa >= CPT) sa -= CPT;
              cache_slot(k, P4[sa * 32 + lane]);
            }
            if (lane >= rl) speculate();
            else {                                   // already integrated: keep what shared memory now holds
              const float4 a = P4[rb * 32 + lane], b = V4[rb * 32 + lane];
              ipx = a.x; ipy = a.y; ipz = a.z; irho = a.w; ivx = b.x; ivy = b.y; ivz = b.z; irad = b.w;
            }
          }
        }
        // integrate row i (world.c:549-553): the owner lane's cache takes the speculative state
        if (lane == rl) cache_slot(0, make_float4(ipx, ipy, ipz, irho));
      }
      __syncwarp();
      if (lane >= flushed && lane < rows_here) {
        P4[rb * 32 + lane] = make_float4(ipx, ipy, ipz, irho);
        V4[rb * 32 + lane] = make_float4(ivx, ivy, ivz, irad);
      }
      __syncwarp();
      // rotate the cache: the next block's bodies move into static slot 0
      {
        const float t0 = qx[0], t1 = qy[0], t2 = qz[0], t3 = qr[0], t4 = qc[0];
#pragma unroll
        for (int k = 0; k + 1 < CPT; k++) { qx[k] = qx[k + 1]; qy[k] = qy[k + 1]; qz[k] = qz[k + 1]; qr[k] = qr[k + 1]; qc[k] = qc[k + 1]; }
        qx[CPT - 1] = t0; qy[CPT - 1] = t1; qz[CPT - 1] = t2; qr[CPT - 1] = t3; qc[CPT - 1] = t4;
      }
    }
  }

  __syncwarp();
#pragma unroll 1
  for (int s = 0; s < n_blocks; s++) {
    const int i = s * 32 + lane;
    if (i < n) {
      const float4 a = P4[i];
      v.pos_rest[base + i] = make_float4(a.x, a.y, a.z, ((rest >> s) & 1u) ? 1.0f : 0.0f);
      v.vel_rad[base + i] = V4[i];
    }
  }
  __syncwarp();
  }   // next world
}

// ---------------------------------------------------------------- aggregate statistics
// One pass over the hot state; block partials -> 8 double atomics.  The same 8 doubles are what a
// multi-GPU run all-reduces (SUM), so the kernel can write straight into the caller's buffer.
__global__ void k_stats_spheres(WorldsView v, double *out, double *out2) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long total = (long long)v.n_worlds * v.cap[1];
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(k / v.cap[1]), i = (int)(k % v.cap[1]);
    if (i >= v.counts[w * 3 + 1]) continue;
    const float4 a = v.pos_rest[k], b = v.vel_rad[k];
    acc[0] += 1.0;
    acc[1] += 0.5 * ((double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z);
    acc[2] += b.x; acc[3] += b.y; acc[4] += b.z;
    acc[6] += a.w != 0.0f ? 1.0 : 0.0;
    const bool fin = isfinite(a.x) && isfinite(a.y) && isfinite(a.z) && isfinite(b.x) && isfinite(b.y) && isfinite(b.z);
    acc[7] += fin ? 0.0 : 1.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long ev = 0;
    for (int w = 0; w < v.n_worlds; w++) ev += v.n_events[w];
    acc[5] += (double)ev;
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    double x = acc[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) { atomicAdd(out + q, x); if (out2) atomicAdd(out2 + q, x); }
  }
}

__global__ void k_stats_general(WorldsView v, double *out, double *out2) {
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
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long ev = 0;
    for (int w = 0; w < v.n_worlds; w++) ev += v.n_events[w];
    acc[5] += (double)ev;
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    double x = acc[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) { atomicAdd(out + q, x); if (out2) atomicAdd(out2 + q, x); }
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

template <int CPT, int WPB>
static int launch_spheres(PgWorlds *w, float dt, int n_steps, int first, int count, cudaStream_t stream, int *counter) {
  static bool configured = false;
  static int resident_blocks = 0;
  if (!configured) {
    // all shared memory the SM has: the kernel wants an 8 KB world image per resident warp
    cudaFuncSetAttribute(k_step_spheres<CPT, WPB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_step_spheres<CPT, WPB>, WPB * 32, 0);
    resident_blocks = std::max(1, per_sm) * w->sm_count;
    configured = true;
  }
  const int blocks = std::min((count + WPB - 1) / WPB, resident_blocks);
  cudaMemsetAsync(counter, 0, sizeof(int), stream);
  k_step_spheres<CPT, WPB><<<blocks, WPB * 32, 0, stream>>>(w->v, dt, n_steps, counter, first, first + count);
  return PG_OK;
}

// Step worlds [first, first+count) on `stream` (the whole batch on the handle's stream by default).
int pg_launch_sphere_range(PgWorlds *w, float dt, int n_steps, int first, int count, cudaStream_t stream, int *counter) {
  const int cap = w->v.cap[1];
  if (w->v.cap[0] > kMaxPlanes) { pg_set_error("sphere mode supports at most %d static planes", kMaxPlanes); return PG_ERR_CAPACITY; }
  if (cap <= 32) launch_spheres<1, 4>(w, dt, n_steps, first, count, stream, counter);
  else if (cap <= 64) launch_spheres<2, 4>(w, dt, n_steps, first, count, stream, counter);
  else if (cap <= 128) launch_spheres<4, 4>(w, dt, n_steps, first, count, stream, counter);
  else if (cap <= 256) launch_spheres<8, 4>(w, dt, n_steps, first, count, stream, counter);
  else if (cap <= 512) launch_spheres<16, 2>(w, dt, n_steps, first, count, stream, counter);
  else { pg_set_error("sphere mode supports at most 512 spheres per world (got capacity %d)", cap); return PG_ERR_CAPACITY; }
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
}

int pg_launch_sphere_step(PgWorlds *w, float dt, int n_steps) {
  return pg_launch_sphere_range(w, dt, n_steps, 0, w->v.n_worlds, w->stream, w->d_work_counter);
}

int pg_launch_worlds_stats(PgWorlds *w, void *dev_out8) {
  PG_CUDA(cudaMemsetAsync(w->d_stats, 0, 8 * sizeof(double), w->stream));
  if (dev_out8) PG_CUDA(cudaMemsetAsync(dev_out8, 0, 8 * sizeof(double), w->stream));
  const int blocks = w->sm_count * 4;
  if (w->mode == PG_MODE_SPHERES) k_stats_spheres<<<blocks, 256, 0, w->stream>>>(w->v, w->d_stats, (double *)dev_out8);
  else k_stats_general<<<blocks, 256, 0, w->stream>>>(w->v, w->d_stats, (double *)dev_out8);
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
