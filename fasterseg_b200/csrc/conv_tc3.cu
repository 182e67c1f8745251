// conv_tc3.cu -- K1c: "channel-major accumulator" variant of the tcgen05 implicit-GEMM conv for wide output-channel counts.
//
// conv_tc.cu computes D[128 pixels, N = Cout] with both operands streamed from shared memory.  For Cout = 128 that is a
// 128 x 128 x 16 MMA per k-step: 4 KB of A + 4 KB of B per 64 tensor-pipe cycles = 128 B/clk, i.e. 100 % of the shared-memory
// read bandwidth -- ncu shows the tensor pipe 31 % active and an MMA issuing every ~130 cycles (profiles/r1_prof_conv_heads8_*).
// Here the roles are swapped: the WEIGHTS are the M operand (M = 128 output channels) and a 256-PIXEL tile is the N
// operand, so one MMA is 128 x 256 x 16: 4 KB + 8 KB per 128 cycles = 96 B/clk (75 %), the same ratio as the 128 x 256 tiles
// library GEMMs use.  Both operands are still K-major (weights [co][ci], NHWC pixels [pixel][ci]), fed by the same TMA boxes.
//   D (TMEM)  : lane = output channel, column = pixel of the 256-pixel tile (tw x th, tw * th = 256).
//   epilogue  : thread = one output channel: scale/shift are per-thread constants; the fp16 results are transposed through
//               the (idle) pipeline buffers into the NHWC [pixel][64-channel slab] SW128 layout and leave by TMA store
//               (channel-offset / strided destinations = zero-copy concat, exactly like conv_tc.cu).
// Used for inference-epilogue convs (and the data-gradient convs, which have no epilogue) with Cin % 64 == 0 and
// Cout % 64 == 0, Cout >= 128 on maps with at least kMinPixels output pixels; everything else stays on conv_tc / conv_tc2.
// Reference call sites: search/seg_oprs.py:245-246 (Head 3x3), search/operations.py:72-83 (refine ConvNorm 3x3).
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int k3Threads = 192;
constexpr int k3MaxStages = 4;
constexpr int k3TilePix = 256;              // N of the MMA
constexpr int k3BK = 64;                    // channels per k-chunk (128-byte rows, SW128)
constexpr uint32_t k3WBytes = 128 * k3BK * 2;       // weight tile  [128 co][64 ci]  = 16 KB
constexpr uint32_t k3XBytes = k3TilePix * k3BK * 2;  // pixel tile   [256 px][64 ci]  = 32 KB
constexpr uint32_t k3StageBytes = k3WBytes + k3XBytes;

struct ConvTc3Params {
  CUtensorMap tmap_x[4];  // NHWC input (parity planes for stride 2), box {64, tw, th, 1}
  CUtensorMap tmap_w;     // packed weights {Kpad, Npad, taps}, box {64, 128, 1}
  CUtensorMap tmap_y;     // output, box {64, tw, th, 1}, SW128
  int taps;
  int tap_map[9], tap_dh[9], tap_dw[9], tap_widx[9];
  int k_chunks;
  int tiles_w, tiles_h;
  int tw, th;
  int Cout;
  int stages;
  uint32_t flags;
  const float* scale;
  const float* shift;
};

__global__ void __launch_bounds__(k3Threads, 1)
conv_tc3_kernel(const __grid_constant__ ConvTc3Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[k3MaxStages];
  __shared__ __align__(8) uint64_t empty_bar[k3MaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  int t = blockIdx.x;
  const int tile_w = t % p.tiles_w;
  t /= p.tiles_w;
  const int tile_h = t % p.tiles_h;
  const int img = t / p.tiles_h;
  const int w0 = tile_w * p.tw;
  const int h0 = tile_h * p.th;
  const int m0 = blockIdx.y * 128;  // first output channel of this CTA
  const int k_iters = p.taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x[0]);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_y);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 256);
    tmem_relinquish();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================= TMA producer (converged warp, elected issue, division-free ring: see conv_tc.cu) =================
    RingPos rp;
    uint8_t* sw = smem;
    for (int tap = 0; tap < p.taps; ++tap) {
      const CUtensorMap* mx = &p.tmap_x[p.tap_map[tap]];
      const int cw = w0 + p.tap_dw[tap];
      const int chh = h0 + p.tap_dh[tap];
      const int wi = p.tap_widx[tap];
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(&empty_bar[rp.s], rp.phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[rp.s], k3StageBytes);
          tma_load_3d(sw, &p.tmap_w, &full_bar[rp.s], kc * k3BK, m0, wi);
          tma_load_4d(sw + k3WBytes, mx, &full_bar[rp.s], kc * k3BK, cw, chh, img);
        }
        __syncwarp();
        sw += k3StageBytes;
        rp.advance(p.stages);
        if (rp.s == 0) sw = smem;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: D[co, pixel] += W[co, k] * X[pixel, k]^T =================
    const uint32_t idesc = umma_idesc_f16(128, k3TilePix);
    const uint32_t s0 = smem_u32(smem);
    const uint64_t da0 = umma_desc_kmajor(s0, 128);
    const uint64_t db0 = umma_desc_kmajor(s0 + k3WBytes, 128);
    RingPos rp;
    uint32_t doff = 0;
    for (int it = 0; it < k_iters; ++it) {
      mbar_wait(&full_bar[rp.s], rp.phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = da0 + doff, db = db0 + doff;
        umma_f16_ss(tmem_base, da, db, idesc, it > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < k3BK / 16; ++k)
          umma_f16_ss(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, 1u);
        umma_commit(&empty_bar[rp.s]);
        if (it == k_iters - 1) umma_commit(&tmem_full_bar);
      }
      __syncwarp();
      doff += k3StageBytes >> 4;
      rp.advance(p.stages);
      if (rp.s == 0) doff = 0;
    }
  } else {
    // ================= epilogue warps 2..5: thread = output channel (TMEM lane), columns = pixels =================
    const int q = warp & 3;
    const int cl = q * 32 + lane;  // channel inside this CTA's 128
    const int co = m0 + cl;
    const bool ch_ok = co < p.Cout;
    const bool relu = (p.flags & FSB_CONV_RELU) != 0;
    const float sc = (ch_ok && (p.flags & FSB_CONV_AFFINE) && p.scale) ? p.scale[co] : 1.0f;
    const float sh = (ch_ok && (p.flags & FSB_CONV_AFFINE) && p.shift) ? p.shift[co] : 0.0f;
    // staging: slab (cl / 64) of [256 pixels][128 B], 16-byte chunk index XOR (pixel & 7) (SW128, as the tensor map expects)
    uint8_t* slab = smem + static_cast<size_t>(cl >> 6) * (k3TilePix * 128);
    const int cbyte = (cl & 63) * 2;
    const int chunk = cbyte >> 4, within = cbyte & 15;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c0 = 0; c0 < k3TilePix; c0 += 64) {
      uint32_t vv[4][16];
#pragma unroll
      for (int b = 0; b < 4; ++b) tmem_ld16(taddr + c0 + 16 * b, vv[b]);
      tmem_ld_wait();
#pragma unroll
      for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int px = c0 + 16 * b + j;
          float v = __uint_as_float(vv[b][j]) * sc + sh;
          v = relu ? fmaxf(v, 0.f) : v;
          *reinterpret_cast<__half*>(slab + px * 128 + ((chunk ^ (px & 7)) << 4) + within) = __float2half_rn(v);
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (warp == 2 && lane == 0) {
      const int slabs = min(2, (p.Cout - m0 + 63) / 64);
      for (int s = 0; s < slabs; ++s) tma_store_4d(&p.tmap_y, smem + static_cast<size_t>(s) * (k3TilePix * 128), m0 + s * 64, w0, h0, img);
      tma_store_commit();
      tma_store_wait_read();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------
int encode_tiled_generic(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_bytes);

static inline int floordiv3(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

constexpr int64_t kMinPixels = 256 * 48;  // below this the grid cannot occupy a third of the machine with 256-pixel tiles

int conv_tc3_supported(const fsb_conv_desc* d, const void* y) {
  const int mode = opt(OPT_CONV_TC3);
  if (mode == 0) return 0;
  if (!(d->ksize == 1 || d->ksize == 3) || !(d->stride == 1 || d->stride == 2) || d->dil != 1) return 0;
  if (d->Cin % 64 != 0 || d->Cout % 64 != 0 || d->Cout < 128) return 0;
  if ((d->x_cstride % 8) != 0 || (d->y_cstride % 8) != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0) return 0;
  if (d->flags & (FSB_CONV_OUT_F32 | FSB_CONV_STATS | FSB_CONV_FORCE_DIRECT)) return 0;
  if (mode == 2) return 1;  // force (tests / tuning)
  return static_cast<int64_t>(d->N) * d->Ho * d->Wo >= kMinPixels;
}

int conv_tc3_launch(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift, void* y,
                    cudaStream_t stream) {
  const ConvGeom g = conv_geom(d);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(wpacked) & 15))
    return set_error(FSB_ERR_INVALID, "conv_tc3: x / wpacked must be 16-byte aligned");
  ConvTc3Params p;
  memset(&p, 0, sizeof(p));
  p.taps = g.taps;
  p.k_chunks = g.kpad / k3BK;
  p.tw = d->Wo >= 32 ? 32 : (d->Wo >= 16 ? 16 : 8);
  p.th = k3TilePix / p.tw;
  p.tiles_w = (d->Wo + p.tw - 1) / p.tw;
  p.tiles_h = (d->Ho + p.th - 1) / p.th;
  p.Cout = d->Cout;
  p.flags = d->flags;
  p.scale = scale;
  p.shift = shift;
  const int k_iters = p.taps * p.k_chunks;
  p.stages = k_iters < k3MaxStages ? k_iters : k3MaxStages;
  if (p.stages < 2) p.stages = 2;  // staging for the epilogue needs 2 x 32 KB
  const size_t smem_bytes = static_cast<size_t>(p.stages) * k3StageBytes + 1024;

  const uint32_t box[4] = {static_cast<uint32_t>(k3BK), static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), 1u};
  const __half* xb = static_cast<const __half*>(x);
  const uint64_t cs = static_cast<uint64_t>(d->x_cstride) * 2;
  if (d->stride == 1) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H), static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {cs, cs * d->W, cs * d->W * d->H};
    int rc = encode_tiled_generic(&p.tmap_x[0], xb, 4, dims, str, box, 128);
    if (rc) return rc;
    for (int r = 0; r < d->ksize; ++r)
      for (int s = 0; s < d->ksize; ++s) {
        const int tp = r * d->ksize + s;
        p.tap_map[tp] = 0;
        p.tap_dh[tp] = r - d->pad + d->off_h;
        p.tap_dw[tp] = s - d->pad + d->off_w;
        p.tap_widx[tp] = tp;
      }
  } else {
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < d->ksize; ++r)
      for (int s = 0; s < d->ksize; ++s) {
        const int tp = r * d->ksize + s;
        const int qh = r - d->pad + d->off_h, qw = s - d->pad + d->off_w;
        const int ph = ((qh % 2) + 2) % 2, pw = ((qw % 2) + 2) % 2;
        p.tap_map[tp] = ph * 2 + pw;
        p.tap_dh[tp] = floordiv3(qh, 2);
        p.tap_dw[tp] = floordiv3(qw, 2);
        p.tap_widx[tp] = tp;
        used[ph * 2 + pw] = true;
      }
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        if (!used[ph * 2 + pw]) continue;
        const int Hp = (d->H - ph + 1) / 2, Wp = (d->W - pw + 1) / 2;
        if (Hp <= 0 || Wp <= 0) return set_error(FSB_ERR_INVALID, "conv_tc3: empty parity plane");
        const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(Wp), static_cast<uint64_t>(Hp), static_cast<uint64_t>(d->N)};
        const uint64_t str[3] = {2 * cs, 2 * cs * d->W, cs * d->W * d->H};
        int rc = encode_tiled_generic(&p.tmap_x[ph * 2 + pw], xb + (static_cast<size_t>(ph) * d->W + pw) * d->x_cstride, 4, dims, str, box, 128);
        if (rc) return rc;
      }
    if (!used[0]) p.tmap_x[0] = p.tmap_x[p.tap_map[0]];
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(g.kpad), static_cast<uint64_t>(g.npad), static_cast<uint64_t>(g.taps)};
    const uint64_t str[2] = {static_cast<uint64_t>(g.kpad) * 2, static_cast<uint64_t>(g.kpad) * g.npad * 2};
    const uint32_t boxw[3] = {static_cast<uint32_t>(k3BK), 128u, 1u};
    int rc = encode_tiled_generic(&p.tmap_w, wpacked, 3, dims, str, boxw, 128);
    if (rc) return rc;
  }
  {
    const uint64_t ycs = static_cast<uint64_t>(d->y_cstride) * 2;
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(d->Wo), static_cast<uint64_t>(d->Ho), static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {ycs, ycs * d->Wo, ycs * d->Wo * d->Ho};
    int rc = encode_tiled_generic(&p.tmap_y, y, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc3_kernel), 200 * 1024, "cudaFuncSetAttribute(conv_tc3)")) return rc;
  dim3 grid(static_cast<unsigned>(p.tiles_w * p.tiles_h * d->N), static_cast<unsigned>((d->Cout + 127) / 128));
  cudaError_t e = launch_kernel(conv_tc3_kernel, grid, dim3(k3Threads), smem_bytes, stream, p);
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc3 launch");
  return FSB_OK;
}

}  // namespace fsb
