// pg_common.cuh -- handle layouts and launch declarations shared by the translation units of
// libphysics_gpu.so.  Nothing in here is part of the C-ABI (include/physics_gpu.h is).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/physics_gpu.h"

void pg_set_error(const char *fmt, ...);
// Exact shortest/longest leaf interval of interval_collision's recursion from (t0, t1) (host).
void pg_leaf_lengths(float t0, float t1, float *lo, float *hi);

#define PG_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      pg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
      return PG_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define PG_CUDA_NULL(call)                                                                    \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      pg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
      return nullptr;                                                                         \
    }                                                                                         \
  } while (0)

#define PG_PIPE_MAX_CHUNKS 16
#define PG_PIPE_DMA 2

enum PgMode { PG_MODE_GENERAL = 0, PG_MODE_SPHERES = 1 };

// Device-side view of a batch of worlds (passed by value to kernels).
struct WorldsView {
  int n_worlds;
  int cap[3];                 // capacity per class (static, dynamic, player)
  PgBody *bodies[3];          // [n_worlds * cap[c]] full records (general mode; cold data in sphere mode)
  int *counts;                // [n_worlds * 3]
  PgEvent *events;            // [n_worlds * max_events]
  int *n_events;              // [n_worlds] events written by the current step call (may exceed max_events)
  int max_events;
  // sphere mode: structure-of-arrays hot state, one float4 pair per dynamic sphere
  float4 *pos_rest;           // [n_worlds * cap[1]]  x, y, z, at_rest (0.0f / 1.0f)
  float4 *vel_rad;            // [n_worlds * cap[1]]  vx, vy, vz, world radius
  float4 *planes;             // [n_worlds * cap[0]]  nx, ny, nz, d    (statics are planes in sphere mode)
  float restitution;          // uniform in sphere mode
  float leaf_lo, leaf_hi;     // extreme leaf lengths of the interval walk for (0, clamp(dt)); set per step call
};

// Host-exchange staging of a sphere batch, as the step kernel itself reads and writes it (no separate pack /
// unpack launches): body i of world w (absolute index) has its position at pos + w*stride + 12 i, its velocity
// at vel + w*stride + 12 i and its at_rest byte at rest + w*rest_stride + i.  That covers both host layouts:
// three split arrays ([world][body][3] floats, [world][body] bytes) and one packed record per world.
struct StageView {
  unsigned char *pos, *vel, *rest;     // device addresses (rest may be null)
  size_t stride, rest_stride;
  int load, store;                     // read the initial state from the stage / also write the final state to it
};

struct PgWorlds {
  int device;
  PgMode mode;
  WorldsView v;               // device pointers
  int *h_counts;              // host mirror of counts
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  long long launches;
  int last_steps;
  double *d_stats;            // 8 doubles: aggregate statistics of this handle's bodies (sphere mode: written by the step launch's epilogue)
  double *d_stats_sum;        // 8 doubles: the same summed over all ranks of a communicator (pg_worlds_stats_allreduce)
  bool stats_valid;           // d_stats describes the current state
  int *d_work_counter;        // world hand-out counter of the sphere kernel
  unsigned *d_cost;           // sphere kernel: clocks/1024 each world took in the last full launch
  int *d_order;               // sphere kernel: worlds sorted by that cost, dearest first (hand-out order)
  bool cost_valid;
  int sm_count;
  bool static_pass_parallel_ok;   // recomputed on upload (general mode)
  // device staging for packed [body][3] state exchange (sphere mode)
  float *d_stage_pos, *d_stage_vel;
  unsigned char *d_stage_rest;
  size_t stage_bodies;
  // chunked host pipeline (pg_worlds_step_host)
  cudaStream_t pipe_stream[3];    // compute: the step kernel of chunk c on stream c % 3
  cudaStream_t pipe_up[PG_PIPE_DMA], pipe_down[PG_PIPE_DMA];   // DMA queues per direction: a download never waits behind an upload,
                                                               // and one copy's set-up hides behind another's transfer
  cudaEvent_t pipe_up_ev[PG_PIPE_MAX_CHUNKS], pipe_k_ev[PG_PIPE_MAX_CHUNKS], pipe_done[2 * PG_PIPE_DMA], pipe_join[3], pipe_fork;
  unsigned char *d_stage_packed;  // device staging for the packed host layout (pg_worlds_step_host_packed)
  size_t stage_packed_bytes;
  int *d_pipe_counters;       // one hand-out counter per chunk
  bool pipe_ready;
  // single-round-trip frame path (pg_worlds_frame): pinned staging for dirty records in, all records + events out
  unsigned char *h_frame;         // pinned
  size_t h_frame_bytes;
  unsigned char *h_types[3];      // host mirror of the collider types per class of the frame path's world (licence for the parallel pass 3)
  // the whole chunk pipeline as one CUDA graph, re-captured when the host buffers or the step change
  cudaGraphExec_t pipe_graph;
  const void *pipe_key_ptr[3];
  float pipe_key_dt;
  int pipe_key_steps, pipe_key_chunks, pipe_key_packed;
  long long pipe_graph_kernels;
};

// One frame of ONE game-sized world without a single DMA (pg_worlds_frame): the step kernel itself reads the records the
// host changed out of the pinned block (mapped into the device's address space) and writes every record, the event
// count and the events of the frame back into it.  `world` < 0: not a frame launch.
#define PG_FRAME_MAX_RANGES 12
struct SmallFrame {
  int world;
  int counts[3];                       // class lengths of this frame (the kernel stores them as the device's counts)
  int n_ranges;
  int cls[PG_FRAME_MAX_RANGES], first[PG_FRAME_MAX_RANGES], count[PG_FRAME_MAX_RANGES];
  const PgBody *in;                    // the dirty records, back to back in range order (mapped pinned memory)
  PgBody *out;                         // [cap0 + cap1 + cap2] records of the three classes at their capacity offsets
  int *cnt_out;                        // events produced by the frame (may exceed max_events)
  PgEvent *ev_out;                     // min(count, max_events) events, unsorted
};
// pg_general.cu
int pg_launch_general_step(PgWorlds *w, float dt, int n_steps, int refresh_nodes, const SmallFrame *frame = nullptr);
bool pg_small_frame_fits(const PgWorlds *w, const int32_t counts[3]);
// pg_spheres.cu
int pg_launch_sphere_step(PgWorlds *w, float dt, int n_steps);
int pg_launch_sphere_range(PgWorlds *w, float dt, int n_steps, int first, int count, cudaStream_t stream, int *counter, double *stats,
                           const StageView *stage = nullptr);
int pg_launch_pack_state_on(PgWorlds *w, cudaStream_t st, int first, int count, int n, const float *pos, const float *vel, const unsigned char *rest);
int pg_launch_unpack_state_on(PgWorlds *w, cudaStream_t st, int first, int count, int n, float *pos, float *vel, unsigned char *rest);
int pg_launch_worlds_stats(PgWorlds *w);
int pg_launch_pack_state(PgWorlds *w, int first, int count, int n, const float *pos, const float *vel, const unsigned char *rest);
int pg_launch_unpack_state(PgWorlds *w, int first, int count, int n, float *pos, float *vel, unsigned char *rest);
