// peer.cu -- SyncBN statistics exchange over NVLink peer memory, usable inside captured CUDA graphs.
//
// A data-parallel supernet step exchanges ~7 000 vectors of <= 1 536 floats (per-channel BatchNorm sums, forward and
// backward, of every training unit).  A library collective per vector costs 10-20 us of latency each and the calls of the
// independent ops -- which run on side streams inside the captured passes -- would serialise on one communicator.  Here every
// rank owns an exchange buffer allocated with cudaMalloc and mapped into its peers through CUDA IPC (one process per GPU;
// NVSwitch gives every peer full bandwidth), and the all-reduce is a single-block kernel in the style of NCCL's LL protocol:
//     PUSH: every thread stores its elements as 8-byte {value, epoch} words straight into EVERY PEER's buffer (one NVLink write
//           each; an 8-byte store is a single transaction, so value and flag arrive together -- no fence, no separate flag)
//     POLL: then spins on its OWN buffer (local memory) until the word of every peer carries this epoch, and adds the values
//           in RANK ORDER
// so an exchange costs one one-way NVLink write latency instead of the remote-read round trips of a pull design (the first version:
// ~20 us per exchange in the step, profiles/r2_dp_2gpu_session10.log).  The sum order is the rank order on every rank, so all
// ranks hold bit-identical statistics (and the result is deterministic).  Slots are handed out in call order inside a REGION (one region per captured graph; the call order is the
// capture order and identical on every rank); `epoch` is a per-region device counter bumped by the first node of the graph, so
// a replay needs no host involvement.  Payloads are double-buffered by epoch parity: a rank can only be one replay ahead of a
// peer (it needs the peer's flags of the previous graph to finish it).
// ORDER: an exchange kernel spins until every rank has arrived, so two ranks that start two DIFFERENT exchanges first (independent
// branches of a captured pass on side streams) can each hold execution resources the other one's missing kernel needs -- the first
// 2-GPU runs dead-locked exactly like that (profiles/r2_dp_2gpu_deadlock.log).  An exchange is therefore TWO kernels:
//   push (on the caller's stream, never waits): my {value, epoch} words into every peer's buffer;
//   pull (spins on LOCAL memory until every peer's words of this epoch have landed, sums in rank order), issued on one of
//        kPullStreams internal streams (slot % kPullStreams, slot order within a stream), forked from / joined to the caller's
//        stream by events (legal under stream capture).
// At most kPullStreams kernels of a rank can be spinning at any time -- far fewer than the device can keep resident -- so a push
// (and any compute kernel) can always be scheduled; by induction over the slot order the oldest outstanding pull of every rank
// completes.  One ordered exchange stream (first fix) was dead-lock free too but serialised ~7 000 exchanges per step.
// The flat gradient all-reduce (1 GB once per step) stays on NCCL (csrc/dp.cu); this file is for the latency-bound part.
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

namespace {
constexpr int kMaxWorld = 8;
constexpr int kMaxRegions = 256;
constexpr int kMaxSlots = 32768;          // flags
constexpr size_t kPayloadWords = 384u << 20;  // 8-byte {value, epoch} words: 3 GB per rank (2 parities x world sources inside); the 16-layer
                                              // supernet's pretrain + search passes at world 8 need ~150 M of them, HBM is 180 GB
constexpr int kMaxVec = 4096;             // floats per exchange (one block, 256 threads)
constexpr int kPullStreams = 8;

struct Slot {
  uint32_t off;  // word offset inside the payload area; layout [parity][source rank][n]
  uint32_t n;
};

struct PeerState {
  int world = 1, rank = 0;
  bool ready = false, enabled = true;
  uint8_t* local = nullptr;
  uint8_t* base[kMaxWorld] = {nullptr};
  unsigned* epochs = nullptr;  // device, [kMaxRegions], private to this rank
  Slot slots[kMaxSlots];
  int n_slots = 0;
  size_t used = 0;
  int region_first[kMaxRegions];
  int region_count[kMaxRegions];
  int region = -1, cursor = 0;
  cudaStream_t xstream[kPullStreams] = {nullptr};   // the pull streams
  cudaEvent_t ev_in[kPullStreams] = {nullptr}, ev_out[kPullStreams] = {nullptr};
} g_peer;

constexpr size_t kFlagBytes = static_cast<size_t>(kMaxSlots) * sizeof(unsigned);
constexpr size_t kBufferBytes = kFlagBytes + kPayloadWords * 8;

struct PeerPtrs {
  uint8_t* base[kMaxWorld];
};

__global__ void epoch_bump_kernel(unsigned* e) {
  pdl_launch_dependents();
  pdl_wait();
  *e += 1;
}

// push: my elements, as {value, epoch} words, into slot [parity][source = rank] of every peer
__global__ void __launch_bounds__(256)
peer_push_kernel(const float* __restrict__ v, int n, PeerPtrs pp, int rank, int world, uint32_t off, const unsigned* __restrict__ epoch_ptr) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned e = *epoch_ptr;
  const size_t par = static_cast<size_t>(e & 1u) * world * n;
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    uint2* dst = reinterpret_cast<uint2*>(pp.base[p] + kFlagBytes) + off + par + static_cast<size_t>(rank) * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint2 w = make_uint2(__float_as_uint(v[i]), e);
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst + i), "r"(w.x), "r"(w.y) : "memory");
    }
  }
}
// pull: v[0..n) <- sum over ranks, in rank order; the peers' elements are polled in my own buffer
__global__ void __launch_bounds__(256)
peer_pull_kernel(float* __restrict__ v, int n, PeerPtrs pp, int rank, int world, uint32_t slot, uint32_t off,
                 const unsigned* __restrict__ epoch_ptr) {
  const unsigned e = *epoch_ptr;
  const size_t par = static_cast<size_t>(e & 1u) * world * n;
  const uint2* mine = reinterpret_cast<const uint2*>(pp.base[rank] + kFlagBytes) + off + par;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < world; ++p) {
      if (p == rank) {
        acc += v[i];
        continue;
      }
      const uint2* src = mine + static_cast<size_t>(p) * n + i;
      uint2 w;
      const long long t0 = clock64();
      for (;;) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(src) : "memory");
        if (w.y == e) break;
        if (clock64() - t0 > 20000000000LL) {  // ~10 s: a rank that never arrives must surface as a launch failure, not a hang
          printf("fsb200: peer exchange timed out (rank %d waiting for rank %d, slot %u, epoch %u)\n", rank, p, slot, e);
          __trap();
        }
      }
      acc += __uint_as_float(w.x);
    }
    v[i] = acc;
  }
}
}  // namespace

bool peer_ready() { return g_peer.ready && g_peer.enabled && g_peer.world > 1; }
int peer_world() { return peer_ready() ? g_peer.world : 1; }

int peer_allreduce_f32(float* buf, int64_t n, cudaStream_t stream) {
  if (!peer_ready() || n <= 0) return FSB_OK;
  if (n > kMaxVec) return set_error(FSB_ERR_INVALID, "peer exchange: vector longer than 4096 floats");
  if (g_peer.region < 0) return set_error(FSB_ERR_INVALID, "peer exchange outside a region (fsb_peer_begin)");
  const int r = g_peer.region;
  int slot;
  if (g_peer.cursor < g_peer.region_count[r]) {
    slot = g_peer.region_first[r] + g_peer.cursor;
    if (g_peer.slots[slot].n != static_cast<uint32_t>(n))
      return set_error(FSB_ERR_INVALID, "peer exchange: the call sequence of this region changed (vector length differs)");
  } else {
    // new slot: regions are built one at a time, so a region's slots are contiguous
    if (g_peer.region_count[r] == 0) g_peer.region_first[r] = g_peer.n_slots;
    if (g_peer.region_first[r] + g_peer.region_count[r] != g_peer.n_slots)
      return set_error(FSB_ERR_INVALID, "peer exchange: region grew after another region was started");
    if (g_peer.n_slots >= kMaxSlots || g_peer.used + 2 * static_cast<size_t>(n) * g_peer.world > kPayloadWords)
      return set_error(FSB_ERR_INVALID, "peer exchange: out of slots / payload space");
    slot = g_peer.n_slots++;
    g_peer.slots[slot].off = static_cast<uint32_t>(g_peer.used);
    g_peer.slots[slot].n = static_cast<uint32_t>(n);
    g_peer.used += 2 * static_cast<size_t>(n) * g_peer.world;
    g_peer.region_count[r]++;
  }
  g_peer.cursor++;
  PeerPtrs pp;
  for (int i = 0; i < kMaxWorld; ++i) pp.base[i] = g_peer.base[i];
  cudaError_t e;
  if (!g_peer.xstream[0]) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    for (int i = 0; i < kPullStreams; ++i) {
      e = cudaStreamCreateWithPriority(&g_peer.xstream[i], cudaStreamNonBlocking, hi);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_peer.ev_in[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_peer.ev_out[i], cudaEventDisableTiming);
      if (e != cudaSuccess) return set_cuda_error(e, "peer exchange streams");
    }
  }
  const int xs = slot % kPullStreams;
  // push on the caller's stream (after the producer of buf, never waits) ...
  FSB_LAUNCH(peer_push_kernel, dim3(1), dim3(256), 0, stream, static_cast<const float*>(buf), static_cast<int>(n), pp, g_peer.rank, g_peer.world,
             g_peer.slots[slot].off, static_cast<const unsigned*>(g_peer.epochs + r));
  e = last_launch_error();
  if (e != cudaSuccess) return set_cuda_error(e, "peer push launch");
  // ... fork: the pull waits for everything the caller's stream has issued so far ...
  e = cudaEventRecord(g_peer.ev_in[xs], stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(g_peer.xstream[xs], g_peer.ev_in[xs], 0);
  if (e != cudaSuccess) return set_cuda_error(e, "peer exchange fork");
  peer_pull_kernel<<<1, 256, 0, g_peer.xstream[xs]>>>(buf, static_cast<int>(n), pp, g_peer.rank, g_peer.world, static_cast<uint32_t>(slot),
                                                     g_peer.slots[slot].off, static_cast<const unsigned*>(g_peer.epochs + r));
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "peer pull launch");
  // ... join: the caller's stream continues after the pull
  e = cudaEventRecord(g_peer.ev_out[xs], g_peer.xstream[xs]);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, g_peer.ev_out[xs], 0);
  if (e != cudaSuccess) return set_cuda_error(e, "peer exchange join");
  return FSB_OK;
}

}  // namespace fsb

using namespace fsb;

extern "C" {

/* allocate this rank's exchange buffer; handle_out: 64 bytes (cudaIpcMemHandle_t) to be all-gathered by the launcher */
int fsb_peer_alloc(void* handle_out64) {
  if (!handle_out64) return set_error(FSB_ERR_INVALID, "fsb_peer_alloc: null handle buffer");
  if (g_peer.local) return set_error(FSB_ERR_INVALID, "fsb_peer_alloc: already allocated");
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&g_peer.local), kBufferBytes);
  if (e != cudaSuccess) return set_cuda_error(e, "fsb_peer_alloc: cudaMalloc");
  e = cudaMemset(g_peer.local, 0, kBufferBytes);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&g_peer.epochs), kMaxRegions * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(g_peer.epochs, 0, kMaxRegions * sizeof(unsigned));
  if (e != cudaSuccess) return set_cuda_error(e, "fsb_peer_alloc: init");
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, g_peer.local);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaIpcGetMemHandle");
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out64, &h, 64);
  return FSB_OK;
}

/* handles: world x 64 bytes in rank order (entry `rank` is this rank's own).  Collective in spirit: every rank calls it after
 * the all-gather and must not exchange before all ranks returned (the launcher barriers). */
int fsb_peer_open(const void* handles, int rank, int world) {
  if (!handles || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || !g_peer.local)
    return set_error(FSB_ERR_INVALID, "fsb_peer_open: bad arguments (call fsb_peer_alloc first; world <= 8)");
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      g_peer.base[p] = g_peer.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + 64 * p, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaIpcOpenMemHandle");
    g_peer.base[p] = static_cast<uint8_t*>(ptr);
  }
  g_peer.rank = rank;
  g_peer.world = world;
  for (int i = 0; i < kMaxRegions; ++i) g_peer.region_first[i] = g_peer.region_count[i] = 0;
  g_peer.ready = true;
  return FSB_OK;
}

int fsb_peer_world(void) { return peer_world(); }
int fsb_peer_enable(int on) {
  g_peer.enabled = on != 0;
  return FSB_OK;
}

/* Start (or restart) issuing the exchanges of `region` on `stream`: bumps the region's epoch on the device (captured as a
 * graph node) and rewinds its slot cursor.  Every rank must issue the same regions with the same exchange sequence. */
int fsb_peer_begin(int region, void* stream) {
  if (!peer_ready()) return FSB_OK;
  if (region < 0 || region >= kMaxRegions) return set_error(FSB_ERR_INVALID, "fsb_peer_begin: region out of range");
  g_peer.region = region;
  g_peer.cursor = 0;
  FSB_LAUNCH(epoch_bump_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), g_peer.epochs + region);
  cudaError_t e = last_launch_error();
  if (e != cudaSuccess) return set_cuda_error(e, "epoch_bump launch");
  return FSB_OK;
}

/* in-place sum over ranks of n <= 4096 floats through peer memory (rank-ordered, deterministic); no-op for a single process */
int fsb_peer_allreduce_f32(void* buf, int64_t n, void* stream) {
  if (!buf && n > 0) return set_error(FSB_ERR_INVALID, "fsb_peer_allreduce_f32: null buffer");
  return peer_allreduce_f32(static_cast<float*>(buf), n, static_cast<cudaStream_t>(stream));
}

int fsb_peer_shutdown(void) {
  for (int p = 0; p < g_peer.world; ++p)
    if (p != g_peer.rank && g_peer.base[p]) cudaIpcCloseMemHandle(g_peer.base[p]);
  if (g_peer.local) cudaFree(g_peer.local);
  if (g_peer.epochs) cudaFree(g_peer.epochs);
  for (int i = 0; i < kPullStreams; ++i) {
    if (g_peer.ev_in[i]) cudaEventDestroy(g_peer.ev_in[i]);
    if (g_peer.ev_out[i]) cudaEventDestroy(g_peer.ev_out[i]);
    if (g_peer.xstream[i]) cudaStreamDestroy(g_peer.xstream[i]);
  }
  g_peer = PeerState();
  return FSB_OK;
}

}  // extern "C"
