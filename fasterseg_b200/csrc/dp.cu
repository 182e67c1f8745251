// dp.cu -- native data-parallel exchange for the training units ("NCCL allreduce over NVLink for
// gradients and BN stats only").
//
// The SyncBN statistics of a training unit are 2*C floats that must cross ranks BETWEEN two kernels of the unit (conv with
// fused statistics -> all-reduce -> finalize/apply; BN-backward reduce -> all-reduce -> apply).  Doing that exchange from
// Python costs a torch.distributed call (~25 us of host time) per unit and forces the unit back onto the unfused entry
// points; the supernet runs ~3 100 units per forward pass set and is host-bound.  Here the library owns a NCCL communicator
// (created from an id the launcher broadcasts) and the fused entry points enqueue the all-reduce on the caller's stream
// themselves, so SyncBN adds no host work at all.
//
// NCCL is resolved at run time with dlopen (the copy torch already loaded, else FSB_NCCL_LIB / the system one): the
// library has no link-time dependency on it and single-process users never touch it.
#include <dlfcn.h>

#include "fsb_internal.h"

namespace fsb {

namespace {
// the slice of NCCL's C ABI used here (nccl.h: ncclUniqueId is 128 opaque bytes passed BY VALUE)
struct NcclId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*PFN_getUniqueId)(NcclId*);
typedef int (*PFN_commInitRank)(NcclComm*, int, NcclId, int);
typedef int (*PFN_allReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*PFN_commDestroy)(NcclComm);
typedef const char* (*PFN_getErrorString)(int);
constexpr int kNcclFloat32 = 7;  // ncclFloat32
constexpr int kNcclSum = 0;      // ncclSum

struct DpState {
  void* handle = nullptr;
  PFN_getUniqueId get_id = nullptr;
  PFN_commInitRank init_rank = nullptr;
  PFN_allReduce all_reduce = nullptr;
  PFN_commDestroy destroy = nullptr;
  PFN_getErrorString err_str = nullptr;
  NcclComm comm = nullptr;
  int world = 1;
  int rank = 0;
  bool enabled = true;  // fsb_dp_enable(0): keep the communicator but run the units single-process (e.g. a rank-0-only reference run)
} g_dp;

int nccl_error(int rc, const char* where) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: NCCL error %d (%s)", where, rc, g_dp.err_str ? g_dp.err_str(rc) : "?");
  return set_error(FSB_ERR_CUDA, buf);
}

int load_nccl() {
  if (g_dp.handle) return FSB_OK;
  const char* override_path = getenv("FSB_NCCL_LIB");
  void* h = nullptr;
  if (override_path && override_path[0]) h = dlopen(override_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host framework already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_error(FSB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded (set FSB_NCCL_LIB)");
  g_dp.get_id = reinterpret_cast<PFN_getUniqueId>(dlsym(h, "ncclGetUniqueId"));
  g_dp.init_rank = reinterpret_cast<PFN_commInitRank>(dlsym(h, "ncclCommInitRank"));
  g_dp.all_reduce = reinterpret_cast<PFN_allReduce>(dlsym(h, "ncclAllReduce"));
  g_dp.destroy = reinterpret_cast<PFN_commDestroy>(dlsym(h, "ncclCommDestroy"));
  g_dp.err_str = reinterpret_cast<PFN_getErrorString>(dlsym(h, "ncclGetErrorString"));
  if (!g_dp.get_id || !g_dp.init_rank || !g_dp.all_reduce || !g_dp.destroy)
    return set_error(FSB_ERR_UNSUPPORTED, "libnccl.so.2 lacks ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy");
  g_dp.handle = h;
  return FSB_OK;
}
}  // namespace

bool peer_ready();                                       // peer.cu
int peer_world();
int peer_allreduce_f32(float*, int64_t, cudaStream_t);

int dp_world() {
  if (peer_ready()) return peer_world();
  return (g_dp.comm && g_dp.enabled) ? g_dp.world : 1;
}

// in-place sum over ranks of n floats, enqueued on `stream`; no-op for a single process.  Small vectors (the per-unit SyncBN
// statistics) go through NVLink peer memory when the launcher set it up (peer.cu), everything else through NCCL.
int dp_allreduce_f32(float* buf, int64_t n, cudaStream_t stream) {
  if (peer_ready() && n <= 4096) return peer_allreduce_f32(buf, n, stream);
  if (!g_dp.comm || !g_dp.enabled || g_dp.world <= 1 || n <= 0) return FSB_OK;
  const int rc = g_dp.all_reduce(buf, buf, static_cast<size_t>(n), kNcclFloat32, kNcclSum, g_dp.comm, stream);
  if (rc != 0) return nccl_error(rc, "ncclAllReduce");
  return FSB_OK;
}

}  // namespace fsb

using namespace fsb;

extern "C" {

int fsb_dp_unique_id(void* out128) {
  if (!out128) return set_error(FSB_ERR_INVALID, "fsb_dp_unique_id: null buffer");
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  memset(&id, 0, sizeof(id));
  const int nrc = g_dp.get_id(&id);
  if (nrc != 0) return nccl_error(nrc, "ncclGetUniqueId");
  memcpy(out128, &id, sizeof(id));
  return FSB_OK;
}

int fsb_dp_init(const void* id128, int rank, int world) {
  if (!id128 || world < 1 || rank < 0 || rank >= world) return set_error(FSB_ERR_INVALID, "fsb_dp_init: bad arguments");
  if (g_dp.comm) return set_error(FSB_ERR_INVALID, "fsb_dp_init: already initialised (call fsb_dp_shutdown first)");
  if (world == 1) return FSB_OK;  // nothing to exchange
  int rc = load_nccl();
  if (rc) return rc;
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm comm = nullptr;
  const int nrc = g_dp.init_rank(&comm, world, id, rank);  // collective: every rank calls it, on its own current device
  if (nrc != 0) return nccl_error(nrc, "ncclCommInitRank");
  g_dp.comm = comm;
  g_dp.world = world;
  g_dp.rank = rank;
  return FSB_OK;
}

int fsb_dp_world(void) { return dp_world(); }

int fsb_dp_enable(int on) {
  g_dp.enabled = on != 0;
  return FSB_OK;
}

int fsb_dp_allreduce_f32(void* buf, int64_t n, void* stream) {
  if (!buf && n > 0) return set_error(FSB_ERR_INVALID, "fsb_dp_allreduce_f32: null buffer");
  return dp_allreduce_f32(static_cast<float*>(buf), n, static_cast<cudaStream_t>(stream));
}

int fsb_dp_shutdown(void) {
  if (g_dp.comm) {
    g_dp.destroy(g_dp.comm);
    g_dp.comm = nullptr;
  }
  g_dp.world = 1;
  g_dp.rank = 0;
  g_dp.enabled = true;
  return FSB_OK;
}

}  // extern "C"
