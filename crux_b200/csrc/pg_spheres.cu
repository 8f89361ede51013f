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

namespace {

using pgm::V3;
using pgm::v3;

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxPlanes = 16;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kHuge = 1.0e30f;

struct Sphere {          // one body's hot state as scalars (register resident)
  float px, py, pz, vx, vy, vz, r;
  int rest;
};

__device__ __forceinline__ float reach(float r, float vx, float vy, float vz, float dt) {
  // >= r + |v| dt with a 1e-4 relative cushion; FMA is fine, this is only a filter
  float sp = sqrtf(__fmaf_rn(vx, vx, __fmaf_rn(vy, vy, vz * vz)));
  return __fmaf_rn(sp, dt, r) * 1.0001f + 1.0e-6f;
}

// Exact visit of two spheres (world.c:481-541 with the sphere-sphere table entries).
// All lanes call it with identical arguments.  Returns true if an event fires; A and B updated.
__device__ __noinline__ bool sphere_pair_exact(Sphere &A, Sphere &B, float dt) {
  V3 cA = v3(0.0f + A.px, 0.0f + A.py, 0.0f + A.pz);     // centre (0,0,0) + position, distance.c:283
  V3 cB = v3(0.0f + B.px, 0.0f + B.py, 0.0f + B.pz);
  V3 vA = v3(A.vx, A.vy, A.vz), vB = v3(B.vx, B.vy, B.vz);
  float rs = A.r + B.r;
  float nA = pgm::norm(vA), nB = pgm::norm(vB);
  {
    // Closest approach of the linear paths over [0,dt].  A leaf interval of interval_collision
    // (len < 1e-6 s) passes only if the computed distance is <= (nA+nB)*len <= 1e-6 (nA+nB) at
    // both ends; computed and real distances differ by a few ulps of the coordinates.  So if the
    // real minimum gap exceeds delta below, no leaf passes and the predicate is false.
    V3 s = pgm::sub(cA, cB), v = pgm::sub(vA, vB);
    float vv = __fmaf_rn(v.x, v.x, __fmaf_rn(v.y, v.y, v.z * v.z));
    float sv = __fmaf_rn(s.x, v.x, __fmaf_rn(s.y, v.y, s.z * v.z));
    float t = vv > 0.0f ? fminf(fmaxf(-sv / vv, 0.0f), dt) : 0.0f;
    float qx = __fmaf_rn(v.x, t, s.x), qy = __fmaf_rn(v.y, t, s.y), qz = __fmaf_rn(v.z, t, s.z);
    float dmin2 = __fmaf_rn(qx, qx, __fmaf_rn(qy, qy, qz * qz));
    float cmax = fmaxf(fmaxf(fmaxf(fabsf(cA.x), fabsf(cA.y)), fabsf(cA.z)),
                       fmaxf(fmaxf(fabsf(cB.x), fabsf(cB.y)), fabsf(cB.z)));
    float delta = 2.0e-5f * (nA + nB) + 8.0e-6f * (cmax + (nA + nB) * dt) + 2.0e-6f;
    float lim = (rs + delta) * 1.00001f;
    if (dmin2 > lim * lim && dmin2 < 3.0e38f) return false;
  }
  if (!pgm::ss_interval(cA, vA, nA, cB, vB, nB, rs, 0.0f, dt)) return false;
  pgm::Contact c = pgm::ss_narrow(cA, vA, cB, vB, rs);
  if (!(c.hit && c.t >= 0)) return false;
  V3 pA = v3(A.px, A.py, A.pz), pB = v3(B.px, B.py, B.pz);
  pgm::ss_resolve(pA, vA, v3(0.0f, 0.0f, 0.0f), pB, vB, v3(0.0f, 0.0f, 0.0f), rs, c.t);
  A.px = pA.x; A.py = pA.y; A.pz = pA.z; A.vx = vA.x; A.vy = vA.y; A.vz = vA.z;
  B.px = pB.x; B.py = pB.y; B.pz = pB.z; B.vx = vB.x; B.vy = vB.y; B.vz = vB.z;
  return true;
}

// Exact visit sphere x plane (per lane, divergent).  Returns true if an event fires.
__device__ __noinline__ bool sphere_plane_exact(Sphere &A, float4 pl, float restitution, float dt) {
  V3 c = v3(0.0f + A.px, 0.0f + A.py, 0.0f + A.pz);
  V3 v = v3(A.vx, A.vy, A.vz);
  V3 n = v3(pl.x, pl.y, pl.z);
  float nv = pgm::norm(v);
  if (!pgm::sp_interval(c, v, nv, n, pl.w, A.r, 0.0f, dt)) return false;
  pgm::Contact k = pgm::sp_narrow(c, v, n, pl.w, A.r, dt);
  if (!(k.hit && k.t >= 0)) return false;
  V3 p = v3(A.px, A.py, A.pz);
  bool rest = pgm::sp_resolve(p, v, v3(0.0f, 0.0f, 0.0f), A.r, restitution, n, pl.w, k.t);
  A.px = p.x; A.py = p.y; A.pz = p.z; A.vx = v.x; A.vy = v.y; A.vz = v.z;
  if (rest) A.rest = 1;
  return true;
}

__device__ __forceinline__ void emit_event(const WorldsView &v, int w, int step, int pass, int ac, int ai, int bc, int bi) {
  int slot = atomicAdd(v.n_events + w, 1);
  if (slot >= v.max_events) return;
  PgEvent e;
  e.world = w; e.step = step; e.event_type = 0;       // WORLD x WORLD -> EVENT_COLLISION
  e.a_class = ac; e.a_index = ai; e.b_class = bc; e.b_index = bi;
  e.pad_ = pass << 8;
  v.events[(size_t)w * v.max_events + slot] = e;
}

template <int CPT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_step_spheres(WorldsView v, float dt_in, int n_steps) {
  __shared__ float4 s_planes[kWarpsPerBlock][kMaxPlanes];
  __shared__ float4 s_rowp[kWarpsPerBlock][32];   // row cache: px,py,pz,rho
  __shared__ float4 s_rowv[kWarpsPerBlock][32];   //            vx,vy,vz,r

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int w = blockIdx.x * kWarpsPerBlock + wib;
  if (w >= v.n_worlds) return;
  const int n = v.counts[w * 3 + 1];
  const int ns = v.counts[w * 3 + 0];
  const float dt = pgm::clamp_dt(dt_in);
  const float restitution = v.restitution;
  const bool record = v.max_events > 0;
  const bool gate_ok = dt > 1.0e-9f;             // the conservative gates divide by dt
  const float inv_dt = gate_ok ? 1.0f / dt : 0.0f;

  if (lane < ns && lane < kMaxPlanes) s_planes[wib][lane] = v.planes[(size_t)w * v.cap[0] + lane];

  float px[CPT], py[CPT], pz[CPT], vx[CPT], vy[CPT], vz[CPT], rad[CPT], rho[CPT];
  unsigned rest = 0;
  const size_t base = (size_t)w * v.cap[1];
#pragma unroll
  for (int s = 0; s < CPT; s++) {
    const int i = s * 32 + lane;
    if (i < n) {
      float4 a = v.pos_rest[base + i], b = v.vel_rad[base + i];
      px[s] = a.x; py[s] = a.y; pz[s] = a.z; vx[s] = b.x; vy[s] = b.y; vz[s] = b.z; rad[s] = b.w;
      if (a.w != 0.0f) rest |= 1u << s;
      rho[s] = reach(rad[s], vx[s], vy[s], vz[s], dt);
    } else {
      px[s] = kHuge * (float)(i + 1); py[s] = kHuge; pz[s] = kHuge; vx[s] = vy[s] = vz[s] = 0.0f; rad[s] = 0.0f; rho[s] = 0.0f;
    }
  }
  __syncwarp();

  for (int step = 0; step < n_steps; step++) {
    // ---------------- pass 3: every sphere against the planes, in plane order
#pragma unroll
    for (int s = 0; s < CPT; s++) {
      const int i = s * 32 + lane;
      if (i < n) {
        for (int k = 0; k < ns; k++) {
          const float4 pl = s_planes[wib][k];
          // conservative gate: signed distance is linear in t; if the sphere surface stays farther
          // than delta from the plane over [0,dt] no leaf of the interval search can pass.
          float s0 = __fmaf_rn(px[s], pl.x, __fmaf_rn(py[s], pl.y, pz[s] * pl.z)) - pl.w;
          float ndv = __fmaf_rn(vx[s], pl.x, __fmaf_rn(vy[s], pl.y, vz[s] * pl.z));
          float s1 = __fmaf_rn(ndv, dt, s0);
          float gap = (s0 > 0.0f) == (s1 > 0.0f) ? fminf(fabsf(s0), fabsf(s1)) : 0.0f;
          float spd = (rho[s] - rad[s]) * inv_dt;   // >= |v| (rho >= r + |v| dt)
          float delta = 2.0e-5f * spd + 8.0e-6f * (fabsf(s0) + fabsf(pl.w) + fabsf(px[s]) + fabsf(py[s]) + fabsf(pz[s]) + spd * dt) + 2.0e-6f;
          if (gate_ok && gap > (rad[s] + delta) * 1.00001f) continue;
          Sphere A;
          A.px = px[s]; A.py = py[s]; A.pz = pz[s]; A.vx = vx[s]; A.vy = vy[s]; A.vz = vz[s]; A.r = rad[s];
          A.rest = (rest >> s) & 1;
          if (sphere_plane_exact(A, pl, restitution, dt)) {
            px[s] = A.px; py[s] = A.py; pz[s] = A.pz; vx[s] = A.vx; vy[s] = A.vy; vz[s] = A.vz;
            if (A.rest) rest |= 1u << s;
            rho[s] = reach(rad[s], vx[s], vy[s], vz[s], dt);
            if (record) emit_event(v, w, step, 3, PG_CLASS_DYNAMIC, i, PG_CLASS_STATIC, k);
          }
        }
      }
    }
    __syncwarp();

    // ---------------- pass 4: rows in order
#pragma unroll 1
    for (int rs_ = 0; rs_ < CPT; rs_++) {
      if (rs_ * 32 >= n) break;
      // stage this block of 32 row bodies in the row cache
      {
        float sx = 0, sy = 0, sz = 0, sr = 0, svx = 0, svy = 0, svz = 0, srad = 0;
#pragma unroll
        for (int s = 0; s < CPT; s++) if (s == rs_) { sx = px[s]; sy = py[s]; sz = pz[s]; sr = rho[s]; svx = vx[s]; svy = vy[s]; svz = vz[s]; srad = rad[s]; }
        s_rowp[wib][lane] = make_float4(sx, sy, sz, sr);
        s_rowv[wib][lane] = make_float4(svx, svy, svz, srad);
      }
      __syncwarp();
      const int rows_here = min(32, n - rs_ * 32);
#pragma unroll 1
      for (int rl = 0; rl < rows_here; rl++) {
        const int i = rs_ * 32 + rl;
        int j0 = 0;
        for (;;) {
          const float4 rp = s_rowp[wib][rl];
          // reach filter over this lane's columns
          int jl = 0x7fffffff;
#pragma unroll
          for (int s = CPT - 1; s >= 0; s--) {
            const int j = s * 32 + lane;
            float dx = rp.x - px[s], dy = rp.y - py[s], dz = rp.z - pz[s];
            float d2 = __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, dz * dz));
            float T = rp.w + rho[s];
            if (d2 <= T * T && j >= j0 && j != i) jl = j;
          }
          const int j = __reduce_min_sync(kFull, jl);
          if (j == 0x7fffffff) break;
          // gather column j and the row body, evaluate exactly (uniform across the warp)
          const int cs = j >> 5, cl = j & 31;
          float cx = 0, cy = 0, cz = 0, cvx = 0, cvy = 0, cvz = 0, cr = 0;
          unsigned crest_bits = 0;
#pragma unroll
          for (int s = 0; s < CPT; s++) if (s == cs) { cx = px[s]; cy = py[s]; cz = pz[s]; cvx = vx[s]; cvy = vy[s]; cvz = vz[s]; cr = rad[s]; }
          Sphere B;
          B.px = __shfl_sync(kFull, cx, cl); B.py = __shfl_sync(kFull, cy, cl); B.pz = __shfl_sync(kFull, cz, cl);
          B.vx = __shfl_sync(kFull, cvx, cl); B.vy = __shfl_sync(kFull, cvy, cl); B.vz = __shfl_sync(kFull, cvz, cl);
          B.r = __shfl_sync(kFull, cr, cl); B.rest = 0;
          (void)crest_bits;
          const float4 rv = s_rowv[wib][rl];
          Sphere A;
          A.px = rp.x; A.py = rp.y; A.pz = rp.z; A.vx = rv.x; A.vy = rv.y; A.vz = rv.z; A.r = rv.w; A.rest = 0;
          if (sphere_pair_exact(A, B, dt)) {
            const float rhoA = reach(A.r, A.vx, A.vy, A.vz, dt), rhoB = reach(B.r, B.vx, B.vy, B.vz, dt);
            // write back: row body to its owner lane + row cache; column body to its owner lane
            if (lane == rl) {
#pragma unroll
              for (int s = 0; s < CPT; s++) if (s == rs_) { px[s] = A.px; py[s] = A.py; pz[s] = A.pz; vx[s] = A.vx; vy[s] = A.vy; vz[s] = A.vz; rho[s] = rhoA; }
            }
            if (lane == cl) {
#pragma unroll
              for (int s = 0; s < CPT; s++) if (s == cs) { px[s] = B.px; py[s] = B.py; pz[s] = B.pz; vx[s] = B.vx; vy[s] = B.vy; vz[s] = B.vz; rho[s] = rhoB; }
            }
            __syncwarp();
            if (lane == 0) {
              s_rowp[wib][rl] = make_float4(A.px, A.py, A.pz, rhoA);
              s_rowv[wib][rl] = make_float4(A.vx, A.vy, A.vz, A.r);
              if (cs == rs_) {      // the column is a later (or earlier) row of this block: refresh its cache line
                s_rowp[wib][cl] = make_float4(B.px, B.py, B.pz, rhoB);
                s_rowv[wib][cl] = make_float4(B.vx, B.vy, B.vz, B.r);
              }
              if (record) emit_event(v, w, step, 4, PG_CLASS_DYNAMIC, i, PG_CLASS_DYNAMIC, j);
            }
            __syncwarp();
          }
          j0 = j + 1;
        }
        // integrate row i (world.c:549-553) in its owner lane
        if (lane == rl) {
#pragma unroll
          for (int s = 0; s < CPT; s++) {
            if (s == rs_ && !((rest >> s) & 1)) {
              V3 p = v3(px[s], py[s], pz[s]), vel = v3(vx[s], vy[s], vz[s]);
              pgm::integrate(p, vel, dt);
              px[s] = p.x; py[s] = p.y; pz[s] = p.z; vx[s] = vel.x; vy[s] = vel.y; vz[s] = vel.z;
              rho[s] = reach(rad[s], vx[s], vy[s], vz[s], dt);
            }
          }
        }
      }
      __syncwarp();
    }
  }

#pragma unroll
  for (int s = 0; s < CPT; s++) {
    const int i = s * 32 + lane;
    if (i < n) {
      v.pos_rest[base + i] = make_float4(px[s], py[s], pz[s], ((rest >> s) & 1) ? 1.0f : 0.0f);
      v.vel_rad[base + i] = make_float4(vx[s], vy[s], vz[s], rad[s]);
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

int pg_launch_sphere_step(PgWorlds *w, float dt, int n_steps) {
  const int cap = w->v.cap[1];
  const int blocks = (w->v.n_worlds + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int threads = kWarpsPerBlock * 32;
  if (w->v.cap[0] > kMaxPlanes) { pg_set_error("sphere mode supports at most %d static planes", kMaxPlanes); return PG_ERR_CAPACITY; }
  if (cap <= 32) k_step_spheres<1><<<blocks, threads, 0, w->stream>>>(w->v, dt, n_steps);
  else if (cap <= 64) k_step_spheres<2><<<blocks, threads, 0, w->stream>>>(w->v, dt, n_steps);
  else if (cap <= 128) k_step_spheres<4><<<blocks, threads, 0, w->stream>>>(w->v, dt, n_steps);
  else if (cap <= 256) k_step_spheres<8><<<blocks, threads, 0, w->stream>>>(w->v, dt, n_steps);
  else if (cap <= 512) k_step_spheres<16><<<blocks, threads, 0, w->stream>>>(w->v, dt, n_steps);
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
