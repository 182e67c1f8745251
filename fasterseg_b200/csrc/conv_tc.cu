// conv_tc.cu -- K1: im2col-free implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces F.conv2d (+ BatchNorm eval + ReLU) for every 3x3 / 1x1 conv of the FasterSeg hot path
// (reference call sites: search/slimmable_ops.py:47, search/operations.py:72-83,146-150,..., search/seg_oprs.py:21-39).
//
// GEMM view:  D[M = 128 output pixels (th x tw tile), N = Cout tile] += A[M, K] * B[N, K]^T,
//             K = taps * Cin, walked as (tap, 64- or 32-channel chunk).
//   A tile  : for tap (r,s) the 128 x BK slab is ONE TMA box {BK, tw, th, 1} of the NHWC input at the
//             tap-shifted coordinate; out-of-image rows/cols and channel tails are zero-filled by TMA, so
//             padding costs nothing and no im2col buffer exists.  Stride-2 convs address one of four
//             parity planes of the input (each its own tensor map), which keeps every box dense.
//   B tile  : packed fp16 weights [tap][Npad][Kpad] (K-major), one TMA box {BK, Ntile, 1}.
//   D       : fp32 accumulator in TMEM (128 lanes x Ntile columns).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2-5 = epilogue (tcgen05.ld -> BN scale/shift -> ReLU -> fp16 -> NHWC store with channel offset/stride,
// which is how torch.cat(dim=1) call sites become free).
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int kTileM = 128;
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;

struct ConvTcParams {
  CUtensorMap tmap_a[4];
  CUtensorMap tmap_b;
  CUtensorMap tmap_y[2];  // output maps for the TMA-store epilogue: [0] 64-channel slabs (SW128), [1] tail slab (dense)
  int tma_store;          // 1: fp16 tile staged in smem and written with cp.async.bulk.tensor (full-line, clipped stores)
  int tail_w;             // channels in the last slab of an N tile when n_tile % 64 != 0
  int taps;
  int tap_map[9];
  int tap_dh[9];
  int tap_dw[9];
  int tap_widx[9];  // which [tap] slice of the packed weights each tap multiplies (identity for plain convs)
  int k_chunks;
  int Ho, Wo;
  int tiles_w, tiles_h;
  int tw, th;
  int Cout;
  int n_tile;
  int stages;
  int y_cstride;
  uint32_t flags;
  uint32_t tmem_cols;
  const float* scale;
  const float* shift;
  __half* y;
  float* stats;   // partial rows (one per spatial tile = blockIdx.x), see bn.cu "Deterministic statistics"
  int stats_C;    // row half-width (row stride = 2 * stats_C floats)
  int stats_off;  // channel offset of this conv's output inside a row
  int ksplit;     // 1, or 3: a cluster of 3 CTAs (blockIdx.z) shares one output tile, CTA z accumulates kernel row z (3 of the 9 taps)
  uint32_t part_off;  // byte offset (from the 1024-aligned smem base) of the (ksplit - 1) partial-sum buffers [n_tile][128] fp32 in CTA 0
  int m_tiles;  // persistent kernel only: spatial tiles (tiles_w * tiles_h * N) ...
  int n_tiles;  // ... x output-channel tiles; a CTA walks tile = blockIdx.x, += gridDim.x (n fastest)
};

template <int BK>
__global__ void __launch_bounds__(kThreads, 2)   // <= 170 registers: two CTAs per SM (96 KB pipelines) on the large maps
conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_scale[256];
  __shared__ float s_shift[256];

  pdl_launch_dependents();  // let the next kernel's prologue overlap our main loop / tail
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kABytes = kTileM * BK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_tile) * BK * 2;
  const uint32_t stage_bytes = kABytes + b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  // tile coordinates
  int t = blockIdx.x;
  const int tile_w = t % p.tiles_w;
  t /= p.tiles_w;
  const int tile_h = t % p.tiles_h;
  const int img = t / p.tiles_h;
  const int w0 = tile_w * p.tw;
  const int h0 = tile_h * p.th;
  const int n0 = blockIdx.y * p.n_tile;
  // split-K over a cluster (K1 on small maps): a tcgen05.mma costs ~150 cycles whatever its N (tools/umma_rate.cu), so a tile's
  // main loop is a serial chain of 9 * Cin / 16 instructions however small the map is.  With ksplit = 3 the three kernel rows
  // run on three SMs; CTAs 1 and 2 hand their fp32 accumulators to CTA 0 through distributed shared memory.
  const int kz = p.ksplit > 1 ? static_cast<int>(blockIdx.z) : 0;
  const int taps_local = p.taps / p.ksplit;
  const int tap0 = kz * taps_local;
  const int k_iters = taps_local * p.k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a[0]);
    tma_prefetch_desc(&p.tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();  // predecessor's outputs (our x) and anything it still reads (our y) are safe from here on
  // epilogue constants
  for (int c = threadIdx.x; c < p.n_tile; c += kThreads) {
    const int ch = n0 + c;
    const bool ok = ch < p.Cout;
    s_scale[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.scale) ? p.scale[ch] : 1.0f;
    s_shift[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.shift) ? p.shift[ch] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // Issue style (tools/umma_rate.cu): the producer and MMA warps stay CONVERGED and only the TMA / MMA instructions themselves
  // sit under elect.sync.  Inside an `if (lane == 0)` region the compiler has to wrap every uniform-datapath instruction
  // (UTMALDG, UTCHMMA, UTCBAR) into an ELECT / R2UR / BRA.U.ANY waterfall, which cost ~130 cycles per MMA in round 1.
  if (warp == 0) {
    // ================= TMA producer =================
    RingPos rp;
    uint8_t* sa = smem;
    for (int tap = tap0; tap < tap0 + taps_local; ++tap) {
      const CUtensorMap* ma = &p.tmap_a[p.tap_map[tap]];
      const int cw = w0 + p.tap_dw[tap];
      const int chh = h0 + p.tap_dh[tap];
      const int wi = p.tap_widx[tap];
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(&empty_bar[rp.s], rp.phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[rp.s], stage_bytes);
          tma_load_4d(sa, ma, &full_bar[rp.s], kc * BK, cw, chh, img);
          tma_load_3d(sa + kABytes, &p.tmap_b, &full_bar[rp.s], kc * BK, n0, wi);
        }
        __syncwarp();
        sa += stage_bytes;
        rp.advance(p.stages);
        if (rp.s == 0) sa = smem;
      }
    }
    if (p.ksplit > 1) cluster_sync_all_threads();
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, one elected lane issues) =================
    // Lean loop: no integer division, descriptors of stage s = descriptors of stage 0 + s * (stage_bytes >> 4) in the
    // start-address field (shared-memory addresses stay below 256 KB, so the 14-bit field never carries).
    const uint32_t idesc = umma_idesc_f16(kTileM, static_cast<uint32_t>(p.n_tile));
    const uint32_t sa0 = smem_u32(smem);
    const uint64_t da0 = umma_desc_kmajor(sa0, BK * 2);
    const uint64_t db0 = umma_desc_kmajor(sa0 + kABytes, BK * 2);
    const uint32_t dstep = stage_bytes >> 4;
    RingPos rp;
    uint32_t doff = 0;
    for (int it = 0; it < k_iters; ++it) {
      mbar_wait(&full_bar[rp.s], rp.phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = da0 + doff, db = db0 + doff;
        // advance 16 fp16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
        umma_f16_ss(tmem_base, da, db, idesc, it > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < BK / 16; ++k) umma_f16_ss(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, 1u);
        umma_commit(&empty_bar[rp.s]);  // smem slot reusable once these MMAs retire
        if (it == k_iters - 1) umma_commit(&tmem_full_bar);   // accumulator complete
      }
      __syncwarp();
      doff += dstep;
      rp.advance(p.stages);
      if (rp.s == 0) doff = 0;
    }
    if (p.ksplit > 1) cluster_sync_all_threads();
  } else {
    // ================= epilogue warps 2..5 =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const int oh = h0 + m / p.tw;
    const int ow = w0 + m % p.tw;
    const bool pix_ok = (oh < p.Ho) && (ow < p.Wo);
    __half* yrow = p.y + (static_cast<size_t>(img) * p.Ho * p.Wo + static_cast<size_t>(oh) * p.Wo + ow) * p.y_cstride + n0;
    const bool out_f32 = (p.flags & FSB_CONV_OUT_F32) != 0;
    float* yrow32 = reinterpret_cast<float*>(p.y) +
                    (static_cast<size_t>(img) * p.Ho * p.Wo + static_cast<size_t>(oh) * p.Wo + ow) * p.y_cstride + n0;
    const bool vec_ok = out_f32 ? ((reinterpret_cast<uintptr_t>(yrow32) & 15) == 0) : ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0);
    const bool relu = (p.flags & FSB_CONV_RELU) != 0;
    const bool do_stats = (p.flags & FSB_CONV_STATS) != 0 && p.stats != nullptr;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float* part = reinterpret_cast<const float*>(smem + p.part_off);   // [ksplit - 1][n_tile][128] in CTA 0
    if (p.ksplit > 1) {
      if (kz > 0) {
        // hand this CTA's partial accumulator to CTA 0: column-major [channel][pixel] so that a warp writes 128 contiguous bytes
        const uint32_t remote = mapa_cluster(smem_u32(smem + p.part_off), 0) + static_cast<uint32_t>(kz - 1) * p.n_tile * 512u + m * 4u;
        for (int c = 0; c < p.n_tile; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) st_cluster_u32(remote + static_cast<uint32_t>(c + j) * 512u, v[j]);
        }
        tc_fence_before();
      }
      cluster_sync_all_threads();   // release / acquire: the partial sums are visible in CTA 0
    }
    if (kz == 0) {
    // per-warp channel statistics of this tile: [4 warps][sum | sumsq][n_tile] in the (now idle) pipeline buffers
    float* s_stat = reinterpret_cast<float*>(smem);
    // TMEM loads are latency-bound (~250 ns per dependent tcgen05.ld + wait): issue up to four 16-column loads, wait once
    for (int c0 = 0; c0 < p.n_tile; c0 += 64) {
      uint32_t vv[4][16];
      const int nb = min(4, (p.n_tile - c0) >> 4);
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < nb) tmem_ld16(taddr + c0 + 16 * b, vv[b]);
      tmem_ld_wait();
#pragma unroll
      for (int b = 0; b < 4; ++b) {
      if (b >= nb) break;
      const int c = c0 + 16 * b;
      uint32_t (&v)[16] = vv[b];
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
      if (p.ksplit > 1) {
        for (int z = 0; z < p.ksplit - 1; ++z) {
          const float* pz = part + (static_cast<size_t>(z) * p.n_tile + c) * 128 + m;
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] += pz[j * 128];
        }
      }
      if (do_stats) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = pix_ok ? f[j] : 0.f;
          float b = a * a;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          if (lane == 0) {
            s_stat[(q * 2 + 0) * p.n_tile + c + j] = a;
            s_stat[(q * 2 + 1) * p.n_tile + c + j] = b;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = f[j] * s_scale[c + j] + s_shift[c + j];
        f[j] = relu ? fmaxf(x, 0.f) : x;
      }
      if (pix_ok && out_f32) {
        const int remaining = p.Cout - (n0 + c);
        if (remaining >= 16 && vec_ok) {
          float4* dst = reinterpret_cast<float4*>(yrow32 + c);
          dst[0] = make_float4(f[0], f[1], f[2], f[3]);
          dst[1] = make_float4(f[4], f[5], f[6], f[7]);
          dst[2] = make_float4(f[8], f[9], f[10], f[11]);
          dst[3] = make_float4(f[12], f[13], f[14], f[15]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < remaining) yrow32[c + j] = f[j];
        }
      } else if (p.tma_store) {
        // stage this thread's 16 channels of pixel m: slab = 64 channels; full slabs use the 128B-swizzled layout of the
        // output tensor map, the tail slab is dense (tail_w * 2 bytes per pixel)
        uint4 o0, o1;
        o0.x = pack_half2(f[0], f[1]);
        o0.y = pack_half2(f[2], f[3]);
        o0.z = pack_half2(f[4], f[5]);
        o0.w = pack_half2(f[6], f[7]);
        o1.x = pack_half2(f[8], f[9]);
        o1.y = pack_half2(f[10], f[11]);
        o1.z = pack_half2(f[12], f[13]);
        o1.w = pack_half2(f[14], f[15]);
        uint8_t* slab = smem + static_cast<size_t>(c0 >> 6) * (kTileM * 128);
        const int ch16 = (c - c0) >> 3;  // 16-byte chunk index inside the slab row
        if (p.n_tile - c0 >= 64) {
          *reinterpret_cast<uint4*>(slab + m * 128 + ((ch16 ^ (m & 7)) << 4)) = o0;
          *reinterpret_cast<uint4*>(slab + m * 128 + (((ch16 + 1) ^ (m & 7)) << 4)) = o1;
        } else {
          uint8_t* row = slab + m * (p.tail_w * 2) + (ch16 << 4);
          *reinterpret_cast<uint4*>(row) = o0;
          *reinterpret_cast<uint4*>(row + 16) = o1;
        }
      } else if (pix_ok) {
        const int remaining = p.Cout - (n0 + c);
        if (remaining >= 16 && vec_ok) {
          uint4 o0, o1;
          o0.x = pack_half2(f[0], f[1]);
          o0.y = pack_half2(f[2], f[3]);
          o0.z = pack_half2(f[4], f[5]);
          o0.w = pack_half2(f[6], f[7]);
          o1.x = pack_half2(f[8], f[9]);
          o1.y = pack_half2(f[10], f[11]);
          o1.z = pack_half2(f[12], f[13]);
          o1.w = pack_half2(f[14], f[15]);
          uint4* dst = reinterpret_cast<uint4*>(yrow + c);
          dst[0] = o0;
          dst[1] = o1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < remaining) yrow[c + j] = __float2half_rn(f[j]);
        }
      }
      }  // 16-column chunk
      if (p.tma_store) {
        fence_proxy_async_smem();           // generic-proxy smem writes -> visible to the TMA engine
        named_bar_sync(1, 128);             // the 4 epilogue warps
        if (warp == 2 && lane == 0) {
          const bool full = (p.n_tile - c0) >= 64;
          tma_store_4d(&p.tmap_y[full ? 0 : 1], smem + static_cast<size_t>(c0 >> 6) * (kTileM * 128), n0 + c0, w0, h0, img);
          tma_store_commit();
        }
      }
    }  // 64-column batch
    if (p.tma_store && warp == 2 && lane == 0) tma_store_wait_read();  // smem must outlive the bulk reads
    if (do_stats) {
      // no atomics: the four warp partials are added in warp order and written as this tile's partial row; the consumer
      // (bn_finalize / rowsum) adds the rows in index order, so the statistics are bit-reproducible run to run
      named_bar_sync(2, 128);
      float* row = p.stats + static_cast<size_t>(blockIdx.x) * 2 * p.stats_C + p.stats_off;
      for (int ch = static_cast<int>(threadIdx.x) - 64; ch < p.n_tile; ch += 128) {
        if (n0 + ch >= p.Cout) continue;
        float a = s_stat[0 * p.n_tile + ch], b = s_stat[1 * p.n_tile + ch];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          a += s_stat[(w * 2 + 0) * p.n_tile + ch];
          b += s_stat[(w * 2 + 1) * p.n_tile + ch];
        }
        row[n0 + ch] = a;
        row[p.stats_C + n0 + ch] = b;
      }
    }
    }  // kz == 0
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// K1p: persistent variant (EXPERIMENTAL, enabled with FSB_CONV_PERSIST=1; inference epilogue only).
//
// The per-tile kernel above pays its prologue (barrier init, TMEM allocation, descriptor prefetch, first TMA round trip) and
// its epilogue (TMEM -> registers -> smem -> TMA store) once per 128-pixel tile, serially with a main loop that lasts only
// 72 MMAs for the student's largest layers -- the tensor pipe idles most of a tile's life.  Here one CTA per SM walks
// tiles tile = blockIdx.x, += gridDim.x with
//   * the TMA producer running ahead across tile boundaries (one continuous stage ring),
//   * TWO accumulators in TMEM: the MMA issuer fills buffer (t & 1) while the epilogue warps drain buffer ((t - 1) & 1),
//   * a dedicated staging buffer for the TMA store, released by cp.async.bulk.wait_group.read before it is rewritten.
// Barriers: full/empty per stage (as above), tmem_full[2] (MMA -> epilogue, tcgen05.commit), tmem_empty[2] (epilogue ->
// MMA, one arrival per epilogue warp after its last tcgen05.ld of the tile).
// ------------------------------------------------------------------------------------------
constexpr int kPersistMaxCout = 512;

template <int BK>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_persistent_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_scale[kPersistMaxCout];
  __shared__ float s_shift[kPersistMaxCout];

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kABytes = kTileM * BK * 2;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_tile) * BK * 2;
  const uint32_t stage_bytes = kABytes + b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + static_cast<size_t>(p.stages) * stage_bytes;  // stage_bytes is a multiple of 1024
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.k_chunks;
  const uint32_t acc_cols = p.tmem_cols >> 1;  // columns of ONE accumulator (the allocation holds two)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a[0]);
    tma_prefetch_desc(&p.tmap_b);
    tma_prefetch_desc(&p.tmap_y[0]);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();
  for (int c = threadIdx.x; c < kPersistMaxCout; c += kThreads) {
    const bool ok = c < p.Cout;
    s_scale[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.scale) ? p.scale[c] : 1.0f;
    s_shift[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.shift) ? p.shift[c] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================= TMA producer: one ring across all tiles of this CTA (converged warp, elected issue) =================
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n0 = (tile % p.n_tiles) * p.n_tile;
      int t = tile / p.n_tiles;
      const int w0 = (t % p.tiles_w) * p.tw;
      t /= p.tiles_w;
      const int h0 = (t % p.tiles_h) * p.th;
      const int img = t / p.tiles_h;
      for (int tap = 0; tap < p.taps; ++tap) {
        const CUtensorMap* ma = &p.tmap_a[p.tap_map[tap]];
        const int cw = w0 + p.tap_dw[tap];
        const int chh = h0 + p.tap_dh[tap];
        for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
          const int s = it % p.stages;
          const int round = it / p.stages;
          if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
          uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
          uint8_t* sb = sa + kABytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
            tma_load_4d(sa, ma, &full_bar[s], kc * BK, cw, chh, img);
            tma_load_3d(sb, &p.tmap_b, &full_bar[s], kc * BK, n0, p.tap_widx[tap]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: alternates between the two TMEM accumulators =================
    const uint32_t idesc = umma_idesc_f16(kTileM, static_cast<uint32_t>(p.n_tile));
    int it = 0;
    int lt = 0;  // tiles done by this CTA
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int use = lt >> 1;  // how often this accumulator has been used before
      if (use > 0) mbar_wait(&tmem_empty_bar[buf], (use - 1) & 1);  // epilogue of its previous tile has drained it
      tc_fence_after();
      const uint32_t acc = tmem_base + static_cast<uint32_t>(buf) * acc_cols;
      for (int ki = 0; ki < k_iters; ++ki, ++it) {
        const int s = it % p.stages;
        const int round = it / p.stages;
        mbar_wait(&full_bar[s], round & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
        const uint32_t sb = sa + kABytes;
        const uint64_t da = umma_desc_kmajor(sa, BK * 2);
        const uint64_t db = umma_desc_kmajor(sb, BK * 2);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss(acc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (ki > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (ki == k_iters - 1) umma_commit(&tmem_full_bar[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps 2..5: drain accumulator (t & 1) while the MMA warp fills the other =================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const bool relu = (p.flags & FSB_CONV_RELU) != 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const int use = lt >> 1;
      const int n0 = (tile % p.n_tiles) * p.n_tile;
      int t = tile / p.n_tiles;
      const int w0 = (t % p.tiles_w) * p.tw;
      t /= p.tiles_w;
      const int h0 = (t % p.tiles_h) * p.th;
      const int img = t / p.tiles_h;
      mbar_wait(&tmem_full_bar[buf], use & 1);
      tc_fence_after();
      if (lt > 0) {
        // the staging buffer still feeds the previous tile's bulk stores until they have READ it
        if (warp == 2 && lane == 0) tma_store_wait_read();
        named_bar_sync(1, 128);
      }
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(buf) * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < p.n_tile; c0 += 64) {
        uint32_t vv[4][16];
        const int nb = min(4, (p.n_tile - c0) >> 4);
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (b < nb) tmem_ld16(taddr + c0 + 16 * b, vv[b]);
        tmem_ld_wait();
        if (c0 + 64 >= p.n_tile) {
          // last TMEM read of this tile by this warp: hand the accumulator back to the MMA issuer
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        uint8_t* slab = staging + static_cast<size_t>(c0 >> 6) * (kTileM * 128);
        const bool full_slab = (p.n_tile - c0) >= 64;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (b >= nb) break;
          const int c = c0 + 16 * b;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = __uint_as_float(vv[b][j]) * s_scale[n0 + c + j] + s_shift[n0 + c + j];
            f[j] = relu ? fmaxf(x, 0.f) : x;
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
          const int ch16 = (c - c0) >> 3;
          if (full_slab) {
            *reinterpret_cast<uint4*>(slab + m * 128 + ((ch16 ^ (m & 7)) << 4)) = o0;
            *reinterpret_cast<uint4*>(slab + m * 128 + (((ch16 + 1) ^ (m & 7)) << 4)) = o1;
          } else {
            uint8_t* row = slab + m * (p.tail_w * 2) + (ch16 << 4);
            *reinterpret_cast<uint4*>(row) = o0;
            *reinterpret_cast<uint4*>(row + 16) = o1;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (warp == 2 && lane == 0) {
          tma_store_4d(&p.tmap_y[full_slab ? 0 : 1], slab, n0 + c0, w0, h0, img);
          tma_store_commit();
        }
      }
    }
    if (warp == 2 && lane == 0) tma_store_wait_read();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int encode_tiled(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return set_error(FSB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr,
                   bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] base %p stride0 %llu",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1],
             rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base, (unsigned long long)strides_bytes[0]);
    return set_error(FSB_ERR_CUDA, buf);
  }
  return FSB_OK;
}

int encode_tiled_generic(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_bytes) {
  return encode_tiled(map, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

ConvGeom conv_geom(const fsb_conv_desc* d) {
  ConvGeom g;
  g.taps = d->ksize * d->ksize;
  g.bk = (d->Cin % 64 == 0) ? 64 : 32;
  g.kpad = (d->Cin + g.bk - 1) / g.bk * g.bk;
  g.npad = (d->Cout + 15) / 16 * 16;
  return g;
}

// spatial tiles (= CTAs along M = partial statistic rows) of the per-tap kernel
int conv_tc_m_tiles(const fsb_conv_desc* d) {
  const int tw = d->Wo >= 16 ? 16 : 8, th = kTileM / tw;
  return ((d->Wo + tw - 1) / tw) * ((d->Ho + th - 1) / th) * d->N;
}

int conv_tc_supported(const fsb_conv_desc* d) {
  if (d->Cin < 16 || (d->x_cstride % 8) != 0) return 0;
  if (!(d->ksize == 1 || d->ksize == 3) || !(d->stride == 1 || d->stride == 2)) return 0;
  return 1;
}

int conv_tc_launch(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift,
                   void* y, float* stats, cudaStream_t stream, const ConvTcCustom* cu) {
  const ConvGeom g = conv_geom(d);
  if (cu && (d->stride != 1 || (d->flags & (FSB_CONV_OUT_F32 | FSB_CONV_STATS))))
    return set_error(FSB_ERR_INVALID, "conv_tc: custom tap tables need a stride-1 fp16 problem");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(wpacked) & 15))
    return set_error(FSB_ERR_INVALID, "conv_tc: x / wpacked must be 16-byte aligned");
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.taps = cu ? cu->ntaps : g.taps;
  p.k_chunks = g.kpad / g.bk;
  p.Ho = cu ? cu->Ho : d->Ho;
  p.Wo = cu ? cu->Wo : d->Wo;
  p.tw = p.Wo >= 16 ? 16 : 8;
  p.th = kTileM / p.tw;
  p.tiles_w = (p.Wo + p.tw - 1) / p.tw;
  p.tiles_h = (p.Ho + p.th - 1) / p.th;
  for (int i = 0; i < 9; ++i) p.tap_widx[i] = i;
  p.Cout = d->Cout;
  // Output-channel tiling.  Default: one N tile (<= 256 columns).  When the spatial tiling alone cannot fill the
  // machine (small maps at 1/16, 1/32 resolution), split N further so that more SMs pull operands from L2 in parallel
  // (each CTA then streams A 16 KB + a thinner B slab per k-step).  B boxes past npad rows are zero-filled by TMA.
  const int m_tiles = p.tiles_w * p.tiles_h * d->N;
  int n_tiles = (g.npad + 255) / 256;
  int n_tile = ((g.npad + n_tiles - 1) / n_tiles + 15) / 16 * 16;
  const int sms = sm_count();
  // FSB_CONV_NTILE_MIN: smallest N tile the split may produce (default 32).  An MMA costs about the same whatever its N,
  // so splitting N does not shorten a CTA's main loop; it only spreads the weight loads and the epilogue over more SMs.
  const int nt_min = opt(OPT_CONV_NTILE_MIN) >= 16 ? opt(OPT_CONV_NTILE_MIN) : 32;
  while (m_tiles * n_tiles < sms && n_tile > nt_min) {
    n_tile = (n_tile / 2 + 15) / 16 * 16;
    n_tiles = (g.npad + n_tile - 1) / n_tile;
  }
  p.n_tile = n_tile;
  p.y_cstride = d->y_cstride;
  p.flags = d->flags;
  p.scale = scale;
  p.shift = shift;
  p.y = static_cast<__half*>(y);
  p.stats = stats;
  p.stats_C = d->stats_C > 0 ? d->stats_C : d->Cout;
  p.stats_off = d->stats_off;
  if ((d->flags & FSB_CONV_STATS) && stats && (d->stats_off < 0 || d->stats_off + d->Cout > p.stats_C))
    return set_error(FSB_ERR_INVALID, "conv_tc: stats_off + Cout exceeds stats_C");
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(n_tile)) cols <<= 1;
  p.tmem_cols = cols;

  const size_t stage_bytes = static_cast<size_t>(kTileM) * g.bk * 2 + static_cast<size_t>(n_tile) * g.bk * 2;
  const int k_iters = p.taps * p.k_chunks;
  // <= one CTA per SM anyway -> give the pipeline (almost) the whole shared memory; otherwise keep 2 CTAs/SM resident
  const size_t smem_budget = (m_tiles * n_tiles <= sms) ? 192 * 1024 : 96 * 1024;
  int stages = static_cast<int>(smem_budget / stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > k_iters) stages = k_iters;
  // ---- split-K over a 3-CTA cluster (one kernel row per CTA) for small maps: the serial MMA chain of a tile shrinks 3x ----
  p.ksplit = 1;
  // Default OFF (FSB_CONV_KSPLIT=1 turns it on): measured on B200 it shortens an isolated small conv by 15-30 % (cell9-0.conv2
  // 10.3 -> 7.1 us) but inside the frame and the supernet step the small convs of independent branches already overlap on side
  // streams, and tripling their CTA count costs more concurrency than the shorter chains win (frame 0.393 -> 0.411 ms, pretrain step
  // 170 -> 191 ms; profiles/r2_conv_bench_ksplit.log).
  if (!cu && g.taps == 9 && opt(OPT_CONV_KSPLIT) > 0 && m_tiles * n_tiles * 3 <= sms && n_tile <= 128) {
    const size_t part_bytes = static_cast<size_t>(2) * n_tile * 512;
    const size_t room = 216 * 1024 - part_bytes;
    int st3 = static_cast<int>(room / stage_bytes);
    const int k_local = 3 * p.k_chunks;
    if (st3 > kMaxStages) st3 = kMaxStages;
    if (st3 > k_local) st3 = k_local;
    if (st3 >= 2 || (st3 >= 1 && k_local == 1)) {
      p.ksplit = 3;
      stages = st3;
      p.part_off = static_cast<uint32_t>(stage_bytes * stages);
    }
  }
  p.stages = stages;
  size_t smem_bytes = stage_bytes * stages + 1024;
  if (p.ksplit > 1) smem_bytes += static_cast<size_t>(2) * n_tile * 512;
  // ---- TMA-store epilogue: fp16 output whose pixels start on 16 B and whose channel count is a multiple of 8 ----
  p.tma_store = 0;
  if (!(d->flags & (FSB_CONV_OUT_F32 | FSB_CONV_STATS)) && d->Cout % 8 == 0 && d->y_cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
      n_tile % 8 == 0 && (cu || opt(OPT_NO_TMA_STORE) <= 0)) {
    const uint64_t ycs = static_cast<uint64_t>(d->y_cstride) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(d->Wo), static_cast<uint64_t>(d->Ho),
                        static_cast<uint64_t>(d->N)};
    uint64_t str[3] = {ycs, ycs * d->Wo, ycs * d->Wo * d->Ho};
    if (cu) {  // output = a strided sub-lattice (parity plane) of the destination tensor
      y = const_cast<void*>(cu->y_base);
      for (int i = 0; i < 4; ++i) dims[i] = cu->y_dims[i];
      for (int i = 0; i < 3; ++i) str[i] = cu->y_strides[i];
    }
    const uint32_t box64[4] = {64u, static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), 1u};
    int rc = 0;
    if (n_tile >= 64) rc = encode_tiled(&p.tmap_y[0], y, 4, dims, str, box64, 128);
    p.tail_w = n_tile % 64;
    if (!rc && p.tail_w) {
      const uint32_t boxt[4] = {static_cast<uint32_t>(p.tail_w), static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), 1u};
      rc = encode_tiled(&p.tmap_y[1], y, 4, dims, str, boxt, 0);
    }
    if (rc) return rc;
    if (n_tile < 64) p.tmap_y[0] = p.tmap_y[1];
    p.tma_store = 1;
    const size_t staging = static_cast<size_t>((n_tile + 63) / 64) * kTileM * 128 + 1024;
    if (smem_bytes < staging) smem_bytes = staging;
  }
  if (cu && !p.tma_store) return set_error(FSB_ERR_UNSUPPORTED, "conv_tc: custom output lattice needs the TMA-store epilogue");

  // ---- A tensor maps ----
  const __half* xb = static_cast<const __half*>(x);
  const uint64_t cs = static_cast<uint64_t>(d->x_cstride) * 2;  // bytes per pixel step
  const uint32_t boxA[4] = {static_cast<uint32_t>(g.bk), static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), 1u};
  if (d->stride == 1) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {cs, cs * d->W, cs * d->W * d->H};
    int rc = encode_tiled(&p.tmap_a[0], xb, 4, dims, str, boxA, g.bk * 2);
    if (rc) return rc;
    for (int r = 0; r < d->ksize; ++r)
      for (int s = 0; s < d->ksize; ++s) {
        const int tp = r * d->ksize + s;
        p.tap_map[tp] = 0;
        p.tap_dh[tp] = r * d->dil - d->pad + d->off_h;
        p.tap_dw[tp] = s * d->dil - d->pad + d->off_w;
      }
    if (cu)
      for (int i = 0; i < cu->ntaps; ++i) {
        p.tap_map[i] = 0;
        p.tap_dh[i] = cu->dh[i];
        p.tap_dw[i] = cu->dw[i];
        p.tap_widx[i] = cu->widx[i];
      }
  } else {
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < d->ksize; ++r)
      for (int s = 0; s < d->ksize; ++s) {
        const int tp = r * d->ksize + s;
        const int qh = r * d->dil - d->pad + d->off_h;
        const int qw = s * d->dil - d->pad + d->off_w;
        const int ph = ((qh % 2) + 2) % 2, pw = ((qw % 2) + 2) % 2;
        p.tap_map[tp] = ph * 2 + pw;
        p.tap_dh[tp] = floordiv(qh, 2);
        p.tap_dw[tp] = floordiv(qw, 2);
        used[ph * 2 + pw] = true;
      }
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        if (!used[ph * 2 + pw]) continue;
        const int Hp = (d->H - ph + 1) / 2, Wp = (d->W - pw + 1) / 2;
        if (Hp <= 0 || Wp <= 0) return set_error(FSB_ERR_INVALID, "conv_tc: empty parity plane");
        const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(Wp), static_cast<uint64_t>(Hp),
                                  static_cast<uint64_t>(d->N)};
        const uint64_t str[3] = {2 * cs, 2 * cs * d->W, cs * d->W * d->H};
        const __half* base = xb + (static_cast<size_t>(ph) * d->W + pw) * d->x_cstride;
        int rc = encode_tiled(&p.tmap_a[ph * 2 + pw], base, 4, dims, str, boxA, g.bk * 2);
        if (rc) return rc;
      }
    // prefetch target must be a valid map
    if (!used[0]) p.tmap_a[0] = p.tmap_a[p.tap_map[0]];
  }
  // ---- B tensor map ----
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(g.kpad), static_cast<uint64_t>(g.npad), static_cast<uint64_t>(g.taps)};
    const uint64_t str[2] = {static_cast<uint64_t>(g.kpad) * 2, static_cast<uint64_t>(g.kpad) * g.npad * 2};
    const uint32_t boxB[3] = {static_cast<uint32_t>(g.bk), static_cast<uint32_t>(n_tile), 1u};
    int rc = encode_tiled(&p.tmap_b, wpacked, 3, dims, str, boxB, g.bk * 2);
    if (rc) return rc;
  }
  cudaError_t e;
  // ---- experimental persistent variant (see K1p above): inference epilogue, plain taps, several tiles per SM ----
  const bool persist_on = opt(OPT_CONV_PERSIST) == 1;
  if (persist_on && !cu && p.tma_store && d->Cout <= kPersistMaxCout && !(d->flags & (FSB_CONV_OUT_F32 | FSB_CONV_STATS)) &&
      m_tiles * n_tiles > sms && cols <= 256) {
    const size_t staging = static_cast<size_t>((n_tile + 63) / 64) * kTileM * 128;
    // tuning knobs for the first measurements: CTAs per SM (1 = whole shared memory for one pipeline, 2 = two issuers
    // sharing the tensor core, TMEM 2 x 2 x cols <= 512) and a cap on the pipeline depth
    const int occ_env = opt(OPT_PERSIST_OCC) == 2 ? 2 : 1;
    const int stages_env = opt(OPT_PERSIST_STAGES) > 0 ? opt(OPT_PERSIST_STAGES) : 0;
    const int occ = (occ_env == 2 && cols * 4 <= 512) ? 2 : 1;
    const size_t budget = (occ == 2 ? 100 : 200) * 1024;
    int pst = budget > staging ? static_cast<int>((budget - staging) / stage_bytes) : 0;
    if (pst > kMaxStages) pst = kMaxStages;
    if (stages_env >= 2 && pst > stages_env) pst = stages_env;
    if (pst >= 2) {
      p.m_tiles = m_tiles;
      p.n_tiles = n_tiles;
      p.tmem_cols = cols * 2;  // two accumulators
      p.stages = pst;
      const size_t psmem = stage_bytes * pst + staging + 1024;
      const unsigned ctas = static_cast<unsigned>(m_tiles * n_tiles < sms * occ ? m_tiles * n_tiles : sms * occ);
      if (g.bk == 64) {
        if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc_persistent_kernel<64>), 216 * 1024, "cudaFuncSetAttribute(conv_tc_persistent<64>)")) return rc;
        e = launch_kernel(conv_tc_persistent_kernel<64>, dim3(ctas), dim3(kThreads), psmem, stream, p);
      } else {
        if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc_persistent_kernel<32>), 216 * 1024, "cudaFuncSetAttribute(conv_tc_persistent<32>)")) return rc;
        e = launch_kernel(conv_tc_persistent_kernel<32>, dim3(ctas), dim3(kThreads), psmem, stream, p);
      }
      if (e != cudaSuccess) return set_cuda_error(e, "conv_tc persistent launch");
      return FSB_OK;
    }
  }
  dim3 grid(static_cast<unsigned>(m_tiles), static_cast<unsigned>(n_tiles), static_cast<unsigned>(p.ksplit));
  if (int rc = ensure_dyn_smem(g.bk == 64 ? reinterpret_cast<const void*>(conv_tc_kernel<64>) : reinterpret_cast<const void*>(conv_tc_kernel<32>),
                               220 * 1024, "cudaFuncSetAttribute(conv_tc)")) return rc;
  if (p.ksplit > 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = static_cast<unsigned>(p.ksplit);
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    e = g.bk == 64 ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<64>, p) : cudaLaunchKernelEx(&cfg, conv_tc_kernel<32>, p);
  } else if (g.bk == 64) {
    e = launch_kernel(conv_tc_kernel<64>, grid, dim3(kThreads), smem_bytes, stream, p);
  } else {
    e = launch_kernel(conv_tc_kernel<32>, grid, dim3(kThreads), smem_bytes, stream, p);
  }
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc launch");
  return FSB_OK;
}

}  // namespace fsb
