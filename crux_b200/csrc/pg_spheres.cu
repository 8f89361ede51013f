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

namespace {

using pgm::V3;
using pgm::v3;

#ifndef PG_SPH_WARPS
#define PG_SPH_WARPS 12   // resident warps (worlds) per SM the register budget is sized for
#endif
constexpr int kMaxPlanes = 16;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kHuge = 1.0e30f;

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
k_step_spheres(WorldsView v, float dt_in, int n_steps) {
  __shared__ float4 s_planes[WPB][kMaxPlanes];
  __shared__ float4 s_P4[WPB][CPT * 32];
  __shared__ float4 s_V4[WPB][CPT * 32];

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int w = blockIdx.x * WPB + wib;
  if (w >= v.n_worlds) return;
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
      a = make_float4(kHuge * (float)(i + 1), kHuge, kHuge, 0.0f);
      b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    P4[i] = a; V4[i] = b;
  }
  __syncwarp();

  float px[CPT], py[CPT], pz[CPT], rho[CPT];

  for (int step = 0; step < n_steps; step++) {
    // ---------------- pass 3: every sphere against the planes, in plane order (world.c:369-471)
#pragma unroll 1
    for (int s = 0; s < n_blocks; s++) {
      const int i = s * 32 + lane;
      if (i < n) {
        float4 a = P4[i];
#pragma unroll 1
        for (int k = 0; k < ns; k++) {
          const float4 pl = s_planes[wib][k];
          // cheap gate: |signed distance| beyond the reach -> the root test of the interval search fails
          const float s0 = __fmaf_rn(a.x, pl.x, __fmaf_rn(a.y, pl.y, a.z * pl.z)) - pl.w;
          if (gate_ok && fabsf(s0) > a.w * 1.0001f + 1.0e-5f * (fabsf(s0) + fabsf(pl.w) + 1.0f)) continue;
          int came_to_rest = 0;
          if (plane_slow_path(P4, V4, i, pl, restitution, dt, gate_ok, &came_to_rest, leaf)) {
            a = P4[i];
            if (came_to_rest) rest |= 1u << s;
            if (record) emit_event(v, w, step, 3, PG_CLASS_DYNAMIC, i, PG_CLASS_STATIC, k);
          }
        }
      }
    }
    __syncwarp();

    // ---------------- pass 4: rows in order (world.c:473-554)
#pragma unroll
    for (int k = 0; k < CPT; k++) { const float4 a = P4[k * 32 + lane]; px[k] = a.x; py[k] = a.y; pz[k] = a.z; rho[k] = a.w; }

#pragma unroll 1
    for (int rb = 0; rb < n_blocks; rb++) {
      // speculative integration of this lane's row body (committed when its row completes;
      // recomputed if a contact touches the world before that)
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
#pragma unroll 1
      for (int rl = 0; rl < rows_here; rl++) {
        const int i = rb * 32 + rl;
        const float4 rp = P4[i];
        unsigned mask = 0;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
          const float dx = rp.x - px[k], dy = rp.y - py[k], dz = rp.z - pz[k];
          const float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz));
          const float T = rp.w + rho[k];
          mask |= (d2 <= T * T ? 1u : 0u) << k;
        }
        if (lane == rl) mask &= ~1u;                 // the row body itself lives in static slot 0
        if (__any_sync(kFull, mask != 0)) {
          // static slot k is actual slot (rb + k) mod CPT: rotate the bits into column order
          const unsigned full = (CPT == 32) ? 0xffffffffu : ((1u << CPT) - 1u);
          const unsigned amask = ((mask << rb) | (mask >> ((CPT - rb) & 31))) & full;
          if (row_slow_path(P4, V4, i, rb == 0 ? mask : amask, lane, CPT, dt, v, w, step, record)) {
#pragma unroll
            for (int k = 0; k < CPT; k++) {
              int sa = rb + k; if (sa >= CPT) sa -= CPT;
              const float4 a = P4[sa * 32 + lane]; px[k] = a.x; py[k] = a.y; pz[k] = a.z; rho[k] = a.w;
            }
            speculate();
          }
        }
        // integrate row i (world.c:549-553): commit the owner lane's speculative state
        if (lane == rl) {
          px[0] = ipx; py[0] = ipy; pz[0] = ipz; rho[0] = irho;
          P4[i] = make_float4(ipx, ipy, ipz, irho);
          V4[i] = make_float4(ivx, ivy, ivz, irad);
        }
      }
      __syncwarp();
      // rotate the cache: the next block's bodies move into static slot 0
      {
        const float tx = px[0], ty = py[0], tz = pz[0], tr = rho[0];
#pragma unroll
        for (int k = 0; k + 1 < CPT; k++) { px[k] = px[k + 1]; py[k] = py[k + 1]; pz[k] = pz[k + 1]; rho[k] = rho[k + 1]; }
        px[CPT - 1] = tx; py[CPT - 1] = ty; pz[CPT - 1] = tz; rho[CPT - 1] = tr;
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
static int launch_spheres(PgWorlds *w, float dt, int n_steps) {
  static bool configured = false;
  if (!configured) {
    // all shared memory the SM has: the kernel wants 24 resident warps, each with an 8 KB world image
    cudaFuncSetAttribute(k_step_spheres<CPT, WPB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  const int blocks = (w->v.n_worlds + WPB - 1) / WPB;
  k_step_spheres<CPT, WPB><<<blocks, WPB * 32, 0, w->stream>>>(w->v, dt, n_steps);
  return PG_OK;
}

int pg_launch_sphere_step(PgWorlds *w, float dt, int n_steps) {
  const int cap = w->v.cap[1];
  if (w->v.cap[0] > kMaxPlanes) { pg_set_error("sphere mode supports at most %d static planes", kMaxPlanes); return PG_ERR_CAPACITY; }
  if (cap <= 32) launch_spheres<1, 4>(w, dt, n_steps);
  else if (cap <= 64) launch_spheres<2, 4>(w, dt, n_steps);
  else if (cap <= 128) launch_spheres<4, 4>(w, dt, n_steps);
  else if (cap <= 256) launch_spheres<8, 4>(w, dt, n_steps);
  else if (cap <= 512) launch_spheres<16, 2>(w, dt, n_steps);
  else { pg_set_error("sphere mode supports at most 512 spheres per world (got capacity %d)", cap); return PG_ERR_CAPACITY; }
  PG_CUDA(cudaGetLastError());
  w->launches++;
  return PG_OK;
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
