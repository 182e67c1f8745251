// conv_tc4.cu -- K1x: CTA-PAIR (cta_group::2) row-rolling variant of the tcgen05 implicit-GEMM conv for the large 3x3
// stride-1 layers of the student frame (reference call sites: search/seg_oprs.py:245-246 Head 3x3, search/operations.py:72-83
// ConvNorm, train/model_seg.py:193-200 stem).
//
// Why a pair kernel (tools/umma_rate.cu measures the numbers): with one CTA per MMA both operands of a 128 x N x 16 step
// stream from shared memory (A 4 KB + B N*32 B per step) WHILE TMA writes the next stages into the same memory; the per-tap
// kernel (conv_tc.cu) therefore ran the tensor pipe at ~31 % (ncu, round 1).  Here
//   * two CTAs of a cluster issue ONE 256 x N x 16 MMA (cta_group::2): every CTA holds its own 128 pixels (A) and only HALF of
//     the weight tile (B, N/2 rows) -> 25..50 % less shared-memory operand traffic per FLOP,
//   * the input is staged ONCE per CTA as whole rows (130 pixels x 64 channels, SW128): the nine taps are nine descriptor
//     start addresses into the same rows (row r -> which row slot, column s -> +s pixels = +s * 128 B; the 128B swizzle is a
//     function of the absolute shared-memory address, so a shifted start reads exactly what TMA wrote), and consecutive
//     output rows re-use two of their three input rows (rolling ring of row slots) -> A is read from L2 (R+2)/R times
//     instead of 9 times,
//   * the weights are either RESIDENT in shared memory for the whole CTA (9 * Cin/64 tiles of N/2 x 64) or streamed once per
//     group of G output rows whose MMA sequences run `lag` tiles apart (the epilogue of row j overlaps the tail of row j+1),
//   * a CTA owns R consecutive output rows of one 128-column strip (one wave: jobs <= SMs), accumulators alternate between
//     TMEM buffers so the epilogue (TMEM -> BN scale/shift -> ReLU -> fp16 -> swizzled staging -> TMA store at the channel
//     offset) of one row runs under the MMAs of the next.
// Warp roles (224 threads): warp 0 = input-row TMA producer, warp 1 = TMEM owner + MMA issuer (leader CTA only), warp 2 =
// weight TMA producer, warps 3..6 = epilogue.  All TMA loads of both CTAs complete on the LEADER's mbarriers (cta_group::2
// form); slot releases and accumulator hand-over travel by multicast tcgen05.commit / remote mbarrier arrive.
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int k4Threads = 224;
constexpr int k4MaxSlots = 12;   // input-row slots
constexpr int k4MaxKch = 4;      // 64-channel chunks of Cin
constexpr int k4MaxB = 36;       // weight tiles resident / in the ring
constexpr int k4MaxAcc = 4;
constexpr int k4Cols = 128;      // output columns per CTA (= rows of A per CTA)
constexpr int k4Halo = k4Cols + 2;
constexpr uint32_t k4RowTx = k4Halo * 128;    // bytes one row box delivers
constexpr uint32_t k4RowBytes = 17 * 1024;    // row buffer (1024-aligned)
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the pair's CTA 0

struct ConvTc4Params {
  CUtensorMap tmap_x;  // {C, W, H, N}, box {64, 130, 1, 1}, SW128
  CUtensorMap tmap_w;  // packed weights {Kpad, Npad, 9}, box {64, n_half, 1}, SW128
  CUtensorMap tmap_y;  // {C, Wo, Ho, N}, box {64, 128, 1, 1}, SW128
  int kch, T;          // 64-channel chunks, weight tiles per row (9 * kch)
  int R, G, lag;       // output rows per CTA, rows per weight pass, tile lag between the rows of a pass
  int NR, NB, nacc;    // input-row slots, weight stages, TMEM accumulators
  int resident;        // 1: all T weight tiles stay in shared memory
  int Ho, Wo, strips, row_groups, jobs;
  int Cout, n_half;
  uint32_t flags, tmem_cols;
  const float* scale;
  const float* shift;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(0));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion bytes are counted on the pair leader's barrier
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far by this thread arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

__global__ void __launch_bounds__(k4Threads, 1)
conv_tc4_kernel(const __grid_constant__ ConvTc4Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[k4MaxSlots * k4MaxKch];
  __shared__ __align__(8) uint64_t a_empty[k4MaxSlots];
  __shared__ __align__(8) uint64_t b_full[k4MaxB];
  __shared__ __align__(8) uint64_t b_empty[k4MaxB];
  __shared__ __align__(8) uint64_t tmem_full[k4MaxAcc];
  __shared__ __align__(8) uint64_t tmem_empty[k4MaxAcc];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_scale[256];
  __shared__ float s_shift[256];

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t slot_bytes = static_cast<uint32_t>(p.kch) * k4RowBytes;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_half) * 128u;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + static_cast<size_t>(p.NR) * slot_bytes;
  uint8_t* staging = smem_b + static_cast<size_t>(p.NB) * b_bytes;

  // job of this CTA: R consecutive output rows of one 128-column strip (jobs past the end are inert partners)
  const int job = blockIdx.x;
  const bool live = job < p.jobs;
  int t = live ? job : 0;
  const int strip = t % p.strips;
  t /= p.strips;
  const int rg = t % p.row_groups;
  const int img = t / p.row_groups;
  const int w0 = strip * k4Cols;
  const int h0 = live ? rg * p.R : p.Ho + 2;  // inert: every box is out of bounds (zero fill), nothing is stored
  const int in_rows = p.R + 2;
  const int groups = (p.R + p.G - 1) / p.G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_y);
    for (int i = 0; i < p.NR * p.kch; ++i) mbar_init(&a_full[i], 1);
    for (int i = 0; i < p.NR; ++i) mbar_init(&a_empty[i], 1);
    for (int i = 0; i < p.NB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < p.nacc; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);  // 4 epilogue warps x 2 CTAs
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  for (int c = threadIdx.x; c < 256; c += k4Threads) {
    const bool ok = c < p.Cout;
    s_scale[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.scale) ? p.scale[c] : 1.0f;
    s_shift[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.shift) ? p.shift[c] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote completion / arrive can target them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // Issue style: the three issuing warps stay converged; only the TMA / MMA / commit instructions are under elect.sync (inside an
  // `if (lane == 0)` region every uniform-datapath instruction becomes an ELECT / R2UR / BRA.U.ANY waterfall, ~130 cycles per MMA).
  if (warp == 0) {
    // ================= input-row producer (both CTAs, own pixels) =================
    const int first = in_rows < p.NR ? in_rows : p.NR;
    // initial fill: chunk-major so the first MMAs (chunk 0 of rows 0..2) can start early
    for (int kc = 0; kc < p.kch; ++kc)
      for (int i = 0; i < first; ++i) {
        uint64_t* bar = &a_full[i * p.kch + kc];
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(bar, 2 * k4RowTx);
          tma_load_4d_2sm(smem_a + static_cast<size_t>(i) * slot_bytes + static_cast<size_t>(kc) * k4RowBytes, &p.tmap_x, bar, kc * 64,
                          w0 - 1, h0 - 1 + i, img);
        }
        __syncwarp();
      }
    for (int i = first; i < in_rows; ++i) {
      const int slot = i % p.NR, use = i / p.NR;
      mbar_wait(&a_empty[slot], (use - 1) & 1);
      for (int kc = 0; kc < p.kch; ++kc) {
        uint64_t* bar = &a_full[slot * p.kch + kc];
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(bar, 2 * k4RowTx);
          tma_load_4d_2sm(smem_a + static_cast<size_t>(slot) * slot_bytes + static_cast<size_t>(kc) * k4RowBytes, &p.tmap_x, bar, kc * 64,
                          w0 - 1, h0 - 1 + i, img);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ================= weight producer (both CTAs, own half of the output channels) =================
    const int passes = p.resident ? 1 : groups;
    int idx = 0;
    for (int ps = 0; ps < passes; ++ps)
      for (int tt = 0; tt < p.T; ++tt, ++idx) {
        const int bs = idx % p.NB, use = idx / p.NB;
        if (use > 0) mbar_wait(&b_empty[bs], (use - 1) & 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&b_full[bs], 2 * b_bytes);
          tma_load_3d_2sm(smem_b + static_cast<size_t>(bs) * b_bytes, &p.tmap_w, &b_full[bs], (tt / 9) * 64, static_cast<int>(rank) * p.n_half,
                          tt % 9);
        }
        __syncwarp();
      }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA; converged warp, one elected lane issues) =================
    if (rank == 0) {
      const uint32_t idesc = umma_idesc_f16(256, static_cast<uint32_t>(p.Cout));
      const bool recycle = in_rows > p.NR;
      for (int grp = 0; grp < groups; ++grp) {
        const int j0 = grp * p.G;
        const int gcur = (p.R - j0) < p.G ? (p.R - j0) : p.G;
        const int steps = p.T + p.lag * (gcur - 1);
        for (int u = 0; u < steps; ++u) {
          for (int g = 0; g < gcur; ++g) {
            const int tt = u - g * p.lag;
            if (tt < 0 || tt >= p.T) continue;
            const int j = j0 + g;
            const int a = j % p.nacc;
            if (tt == 0 && j >= p.nacc) {
              mbar_wait(&tmem_empty[a], ((j / p.nacc) - 1) & 1);   // both epilogues have drained this accumulator
              tc_fence_after();
            }
            const int kc = tt / 9, tap = tt % 9;
            const int r = tap / 3, s = tap % 3;
            const int i = j + r;
            const int slot = i % p.NR;
            mbar_wait(&a_full[slot * p.kch + kc], (i / p.NR) & 1);
            const int bidx = p.resident ? tt : grp * p.T + tt;
            const int bs = bidx % p.NB;
            mbar_wait(&b_full[bs], (bidx / p.NB) & 1);
            tc_fence_after();
            const uint64_t da = umma_desc_kmajor(smem_u32(smem_a + static_cast<size_t>(slot) * slot_bytes + static_cast<size_t>(kc) * k4RowBytes) +
                                                     static_cast<uint32_t>(s) * 128u, 128);
            const uint64_t db = umma_desc_kmajor(smem_u32(smem_b + static_cast<size_t>(bs) * b_bytes), 128);
            const uint32_t acc = tmem_base + static_cast<uint32_t>(a * p.Cout);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ss_2sm(acc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (tt > 0 || k > 0) ? 1u : 0u);
              if (!p.resident && g == gcur - 1) umma_commit_2sm(&b_empty[bs]);   // last row of the pass has consumed this tile
              if (tt == p.T - 1) {
                umma_commit_2sm(&tmem_full[a]);
                if (recycle) umma_commit_2sm(&a_empty[j % p.NR]);   // input row j is not needed by any later output row
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ================= epilogue warps 3..6: thread = one output pixel (TMEM lane) =================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const bool relu = (p.flags & FSB_CONV_RELU) != 0;
    const uint32_t empty_remote = mapa_rank0(smem_u32(&tmem_empty[0]));
    const bool storer = (warp == 3 && lane == 0);
    for (int j = 0; j < p.R; ++j) {
      const int a = j % p.nacc;
      const int oh = h0 + j;
      mbar_wait(&tmem_full[a], (j / p.nacc) & 1);
      tc_fence_after();
      if (j > 0) {
        if (storer) tma_store_wait_read();   // the staging slabs still feed the previous row's bulk stores
        named_bar_sync(1, 128);
      }
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(a * p.Cout) + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < p.Cout; c0 += 64) {
        uint32_t vv[4][16];
#pragma unroll
        for (int b = 0; b < 4; ++b) tmem_ld16(taddr + c0 + 16 * b, vv[b]);
        tmem_ld_wait();
        if (c0 + 64 >= p.Cout) {
          tc_fence_before();
          if (lane == 0) mbar_arrive_cluster(empty_remote + static_cast<uint32_t>(a) * 8u);
        }
        uint8_t* slab = staging + static_cast<size_t>(c0 >> 6) * (k4Cols * 128);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int c = c0 + 16 * b;
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float x = __uint_as_float(vv[b][e]) * s_scale[c + e] + s_shift[c + e];
            f[e] = relu ? fmaxf(x, 0.f) : x;
          }
          uint4 o0, o1;
          o0.x = pack_half2(f[0], f[1]);
          o0.y = pack_half2(f[2], f[3]);
          o0.z = pack_half2(f[4], f[5]);
          o0.w = pack_half2(f[6], f[7]);
          o1.x = pack_half2(f[8], f[9]);
          o1.y = pack_half2(f[10], f[11]);
          o1.z = pack_half2(f[12], f[13]);
          o1.w = pack_half2(f[14], f[15]);
          const int ch16 = 2 * b;
          *reinterpret_cast<uint4*>(slab + m * 128 + ((ch16 ^ (m & 7)) << 4)) = o0;
          *reinterpret_cast<uint4*>(slab + m * 128 + (((ch16 + 1) ^ (m & 7)) << 4)) = o1;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (storer && live && oh < p.Ho) {
        for (int c0 = 0; c0 < p.Cout; c0 += 64)
          tma_store_4d(&p.tmap_y, staging + static_cast<size_t>(c0 >> 6) * (k4Cols * 128), c0, w0, oh, img);
        tma_store_commit();
      }
    }
    if (storer) tma_store_wait_read();
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();   // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
int encode_tiled_generic(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_bytes);

struct Tc4Plan {
  int ok;
  int kch, T, R, G, lag, NR, NB, nacc, resident, strips, row_groups, jobs, n_half;
  uint32_t tmem_cols;
  size_t smem_bytes;
};

static Tc4Plan conv_tc4_plan(const fsb_conv_desc* d) {
  Tc4Plan q;
  memset(&q, 0, sizeof(q));
  const int sms = sm_count();
  q.kch = d->Cin / 64;
  q.T = 9 * q.kch;
  q.n_half = d->Cout / 2;
  q.strips = (d->Wo + k4Cols - 1) / k4Cols;
  const int64_t units = static_cast<int64_t>(d->N) * d->Ho * q.strips;
  int R = static_cast<int>((units + sms - 1) / sms);
  if (R < 1) R = 1;
  for (;; ++R) {   // one wave: jobs (rounded up to whole pairs) must not exceed the SM count
    if (R > d->Ho) return q;   // more image strips than SMs: not a single-wave problem
    q.row_groups = (d->Ho + R - 1) / R;
    q.jobs = d->N * q.strips * q.row_groups;
    if (((q.jobs + 1) & ~1) <= (sms & ~1)) break;
  }
  q.R = R;
  const size_t slot = static_cast<size_t>(q.kch) * k4RowBytes;
  const size_t b_bytes = static_cast<size_t>(q.n_half) * 128;
  const size_t staging = static_cast<size_t>(d->Cout / 64) * (k4Cols * 128);
  const size_t budget = 220 * 1024 - staging - 1024;   // dynamic + 3.2 KB static <= 227 KB opt-in limit
  const int in_rows = R + 2;
  // resident weights when they leave room for at least 4 row slots (or every input row of the job)
  const int need_rows = in_rows < 4 ? in_rows : 4;
  if (q.T <= k4MaxB && q.T * b_bytes + need_rows * slot <= budget) {
    q.resident = 1;
    q.G = 1;
    q.lag = 0;
    q.NB = q.T;
    int nr = static_cast<int>((budget - q.T * b_bytes) / slot);
    if (nr > in_rows) nr = in_rows;
    if (nr > k4MaxSlots) nr = k4MaxSlots;
    q.NR = nr;
    q.nacc = (2 * d->Cout <= 512) ? 2 : 1;
    if (q.nacc > R) q.nacc = R;
  } else {
    q.resident = 0;
    q.G = R >= 2 ? 2 : 1;
    if (q.G * d->Cout > 512) q.G = 1;
    int nr = q.G + 2;
    if (nr > in_rows) nr = in_rows;
    if (nr * slot + 3 * b_bytes > budget) return q;  // does not fit
    // an extra row slot (prefetch of the next group's row) only if the weight ring keeps >= 4 stages
    if (nr < in_rows && (nr + 1) * slot + 4 * b_bytes <= budget) ++nr;
    q.NR = nr;
    int nb = static_cast<int>((budget - nr * slot) / b_bytes);
    if (nb > k4MaxB) nb = k4MaxB;
    if (nb > q.T * ((R + q.G - 1) / q.G)) nb = q.T * ((R + q.G - 1) / q.G);
    q.NB = nb;
    q.lag = q.G > 1 ? (nb - 2 > 5 ? 5 : nb - 2) : 0;
    if (q.G > 1 && q.lag < 1) { q.G = 1; q.lag = 0; }
    q.nacc = (2 * q.G * d->Cout <= 512) ? 2 * q.G : q.G;
    if (q.nacc > R) q.nacc = R;
    if (q.nacc > k4MaxAcc) q.nacc = k4MaxAcc;
  }
  if (q.NR < 3 && q.NR < in_rows) return q;
  if (q.NR * q.kch > k4MaxSlots * k4MaxKch) return q;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(q.nacc * d->Cout)) cols <<= 1;
  if (cols > 512) return q;
  q.tmem_cols = cols;
  q.smem_bytes = q.NR * slot + q.NB * b_bytes + staging + 1024;
  q.ok = 1;
  return q;
}

int conv_tc4_supported(const fsb_conv_desc* d, const void* y) {
  const int mode = opt(OPT_CONV_TC4);
  if (mode <= 0) return 0;   // TODO(validation): default on once the GPU suite has run with it
  if (d->ksize != 3 || d->stride != 1 || d->dil != 1 || d->pad != 1 || d->off_h || d->off_w) return 0;
  if (d->Cin % 64 != 0 || d->Cin / 64 > k4MaxKch || d->Cout % 64 != 0 || d->Cout > 256) return 0;
  if ((d->x_cstride % 8) != 0 || (d->y_cstride % 8) != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0) return 0;
  if (d->flags & (FSB_CONV_OUT_F32 | FSB_CONV_STATS | FSB_CONV_FORCE_DIRECT)) return 0;
  if (d->Wo < 96) return 0;
  if (mode != 2 && static_cast<int64_t>(d->N) * d->Ho * d->Wo < 128 * 96) return 0;  // enough strips x rows to occupy most SMs
  return conv_tc4_plan(d).ok;
}

int conv_tc4_launch(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift, void* y,
                    cudaStream_t stream) {
  const ConvGeom g = conv_geom(d);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(wpacked) & 15))
    return set_error(FSB_ERR_INVALID, "conv_tc4: x / wpacked must be 16-byte aligned");
  const Tc4Plan q = conv_tc4_plan(d);
  if (!q.ok) return set_error(FSB_ERR_UNSUPPORTED, "conv_tc4: no shared-memory plan for this shape");
  ConvTc4Params p;
  memset(&p, 0, sizeof(p));
  p.kch = q.kch; p.T = q.T; p.R = q.R; p.G = q.G; p.lag = q.lag; p.NR = q.NR; p.NB = q.NB; p.nacc = q.nacc; p.resident = q.resident;
  p.Ho = d->Ho; p.Wo = d->Wo; p.strips = q.strips; p.row_groups = q.row_groups; p.jobs = q.jobs;
  p.Cout = d->Cout; p.n_half = q.n_half; p.flags = d->flags; p.tmem_cols = q.tmem_cols;
  p.scale = scale; p.shift = shift;
  const uint64_t cs = static_cast<uint64_t>(d->x_cstride) * 2;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H), static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {cs, cs * d->W, cs * d->W * d->H};
    const uint32_t box[4] = {64u, static_cast<uint32_t>(k4Halo), 1u, 1u};
    if (int rc = encode_tiled_generic(&p.tmap_x, x, 4, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(g.kpad), static_cast<uint64_t>(g.npad), 9ull};
    const uint64_t str[2] = {static_cast<uint64_t>(g.kpad) * 2, static_cast<uint64_t>(g.kpad) * g.npad * 2};
    const uint32_t box[3] = {64u, static_cast<uint32_t>(q.n_half), 1u};
    if (int rc = encode_tiled_generic(&p.tmap_w, wpacked, 3, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t ycs = static_cast<uint64_t>(d->y_cstride) * 2;
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(d->Wo), static_cast<uint64_t>(d->Ho), static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {ycs, ycs * d->Wo, ycs * d->Wo * d->Ho};
    const uint32_t box[4] = {64u, static_cast<uint32_t>(k4Cols), 1u, 1u};
    if (int rc = encode_tiled_generic(&p.tmap_y, y, 4, dims, str, box, 128)) return rc;
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc4_kernel), 222 * 1024, "cudaFuncSetAttribute(conv_tc4)")) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>((q.jobs + 1) & ~1));
  cfg.blockDim = dim3(k4Threads);
  cfg.dynamicSmemBytes = q.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc4_kernel, p);
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc4 launch");
  return FSB_OK;
}

}  // namespace fsb
