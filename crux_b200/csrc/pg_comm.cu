// pg_comm.cu -- the one collective of the path, inside the library: a SUM all-reduce of the 8-double aggregate
// statistics vector over NCCL (NVLink / NVSwitch), enqueued on the handle's own stream behind the step launch
// whose epilogue produced the per-GPU partials.  No host synchronisation on the way, so a C host (the reference's
// game loop, scene.c:401) gets multi-GPU runs without Python.
//
// Worlds are independent units: body data never crosses GPUs; this reduction is all the
// ranks ever exchange.  NCCL is bound at run time (dlopen "libnccl.so.2"): inside a process that already loaded a
// copy (torch bundles one) that very copy is used, and the library still loads on hosts without NCCL --
// pg_comm_* then fail with PG_ERR_UNSUPPORTED and say why.  The few NCCL declarations needed are restated below
// (nccl.h: NCCL_UNIQUE_ID_BYTES 128, ncclDouble = 8, ncclSum = 0) so that the build does not depend on the header.
#include "pg_common.cuh"

#include <dlfcn.h>

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclDouble = 8, kNcclSum = 0;

struct NcclApi {
  void *dll = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  bool tried = false;
};
NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.tried) return g_nccl.dll != nullptr;
  g_nccl.tried = true;
  const char *names[] = {getenv("PG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *name : names) {
    if (!name || !*name) continue;
    g_nccl.dll = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.dll) break;
  }
  if (!g_nccl.dll) { pg_set_error("NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return false; }
  *(void **)&g_nccl.GetUniqueId = dlsym(g_nccl.dll, "ncclGetUniqueId");
  *(void **)&g_nccl.CommInitRank = dlsym(g_nccl.dll, "ncclCommInitRank");
  *(void **)&g_nccl.CommDestroy = dlsym(g_nccl.dll, "ncclCommDestroy");
  *(void **)&g_nccl.AllReduce = dlsym(g_nccl.dll, "ncclAllReduce");
  *(void **)&g_nccl.GetErrorString = dlsym(g_nccl.dll, "ncclGetErrorString");
  *(void **)&g_nccl.GetVersion = dlsym(g_nccl.dll, "ncclGetVersion");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
    pg_set_error("NCCL library lacks a required symbol");
    dlclose(g_nccl.dll);
    g_nccl.dll = nullptr;
    return false;
  }
  return true;
}

const char *nccl_err(ncclResult_t r) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"; }

}  // namespace

struct PgComm {
  ncclComm_t comm;
  int rank, n_ranks, device;
};

extern "C" {

int pg_comm_unique_id(void *out128) {
  if (!out128) { pg_set_error("pg_comm_unique_id: null buffer"); return PG_ERR_ARG; }
  if (!nccl_load()) return PG_ERR_UNSUPPORTED;
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) { pg_set_error("ncclGetUniqueId: %s", nccl_err(r)); return PG_ERR_CUDA; }
  memcpy(out128, id.internal, 128);
  return PG_OK;
}

PgComm *pg_comm_init(const void *unique_id128, int rank, int n_ranks, int device) {
  if (!unique_id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) { pg_set_error("pg_comm_init: bad argument"); return nullptr; }
  if (!nccl_load()) return nullptr;
  PG_CUDA_NULL(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(id.internal, unique_id128, 128);
  PgComm *c = new PgComm();
  c->comm = nullptr; c->rank = rank; c->n_ranks = n_ranks; c->device = device;
  const ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
  if (r != 0) { pg_set_error("ncclCommInitRank(rank %d of %d): %s", rank, n_ranks, nccl_err(r)); delete c; return nullptr; }
  return c;
}

void pg_comm_destroy(PgComm *c) {
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) { cudaSetDevice(c->device); g_nccl.CommDestroy(c->comm); }
  delete c;
}

int pg_comm_rank(const PgComm *c) { return c ? c->rank : -1; }
int pg_comm_size(const PgComm *c) { return c ? c->n_ranks : 0; }
int pg_comm_nccl_version(void) {
  int v = 0;
  if (!nccl_load() || !g_nccl.GetVersion || g_nccl.GetVersion(&v) != 0) return 0;
  return v;
}

// The step launch's epilogue left this GPU's partial sums in d_stats; reduce them over the communicator into
// d_stats_sum, on the handle's stream (ordered behind the launch, ahead of the next one), and move the end mark
// of the step's CUDA-event pair behind the collective so that pg_worlds_last_step_ms covers launch + collective.
int pg_worlds_stats_allreduce(PgWorlds *w, PgComm *c, PgStats *host_out, void *dev_out8) {
  if (!w) { pg_set_error("pg_worlds_stats_allreduce: bad handle"); return PG_ERR_ARG; }
  PG_CUDA(cudaSetDevice(w->device));
  int rc = pg_launch_worlds_stats(w);
  if (rc) return rc;
  if (c) {
    if (c->device != w->device) { pg_set_error("pg_worlds_stats_allreduce: communicator and handle live on different devices"); return PG_ERR_ARG; }
    const ncclResult_t r = g_nccl.AllReduce(w->d_stats, w->d_stats_sum, 8, kNcclDouble, kNcclSum, c->comm, w->stream);
    if (r != 0) { pg_set_error("ncclAllReduce: %s", nccl_err(r)); return PG_ERR_CUDA; }
  } else {
    PG_CUDA(cudaMemcpyAsync(w->d_stats_sum, w->d_stats, 8 * sizeof(double), cudaMemcpyDeviceToDevice, w->stream));   // a single GPU is its own sum
  }
  if (dev_out8) PG_CUDA(cudaMemcpyAsync(dev_out8, w->d_stats_sum, 8 * sizeof(double), cudaMemcpyDeviceToDevice, w->stream));
  PG_CUDA(cudaEventRecord(w->ev1, w->stream));
  if (host_out) {
    double h[8];
    PG_CUDA(cudaMemcpyAsync(h, w->d_stats_sum, sizeof(h), cudaMemcpyDeviceToHost, w->stream));
    PG_CUDA(cudaStreamSynchronize(w->stream));
    host_out->bodies = h[0]; host_out->kinetic = h[1];
    host_out->momentum[0] = h[2]; host_out->momentum[1] = h[3]; host_out->momentum[2] = h[4];
    host_out->contacts = h[5]; host_out->at_rest = h[6]; host_out->nonfinite = h[7];
  }
  return PG_OK;
}

}  // extern "C"
