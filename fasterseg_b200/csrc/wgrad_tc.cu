// wgrad_tc.cu -- K7: convolution weight gradient on tcgen05 tensor cores.
//
//   dW[co, ci, r, s] = sum over output pixels p of  dy[p, co] * x[p * stride + (r, s) - pad, ci]
//
// Per filter tap this is a GEMM  D[M = co, N = ci] = A^T[M, K] * B[K, N]  with K = output pixels.  Both operands live in
// memory as NHWC, i.e. pixel-major rows of channel-contiguous data, which is exactly the UMMA "MN-major" canonical layout:
// a TMA box {64 channels, tw, th, 1} lands in shared memory as 128 rows (pixels = K) x 128 bytes (64 channels = M or N),
// 128B-swizzled; 8-pixel groups are 1024 B apart (stride byte offset) and successive 64-channel blocks of the same pixels are
// separate boxes, one tile (16 KB) apart (leading byte offset).  The x tile of tap (r,s) is the dy tile's pixel block
// shifted by the tap (out-of-image pixels zero-filled by TMA; stride-2 convs read the matching parity plane), exactly like
// the forward kernel's A operand.  No transposes, no im2col.
// A CTA owns (co tile of 128, ci tile <= 256, tap, pixel chunk): the 128 x ci_tile fp32 accumulator sits in TMEM while it
// walks its pixel chunk 128 pixels per pipeline stage; the epilogue adds the tile into the fp32 master-layout gradient with
// atomics (pixel chunks of the same tile race benignly), scaled by 1/loss-scale.
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 6;
constexpr int kWgPix = 128;           // K per stage
constexpr int kWgSub = kWgPix * 128;  // bytes of one [128 pixels x 64 channels] sub-tile

struct WgradTcParams {
  CUtensorMap tmap_dy;     // {Cout, Wo, Ho, N}, box {64, tw, th, 1}
  CUtensorMap tmap_x[4];   // input (parity planes for stride 2), box {64, tw, th, 1}
  int taps, ksize;
  int tap_map[9], tap_dh[9], tap_dw[9];
  int tiles_w, tiles_h, n_img;  // pixel tiling of the OUTPUT map
  int tw, th;
  int Cout, Cin;
  int co_tiles, ci_tiles, ci_tile;  // ci_tile: multiple of 64, <= 256
  int chunks;                       // pixel-tile chunks (grid.y)
  int stages;
  uint32_t tmem_cols;
  float inv_gscale;
  float* dw;
  long long so, si;
};

// K-major-in-memory-is-pixels: MN-major operand descriptor, 128B swizzle.
//   LBO = distance between consecutive 64-element (128 B) blocks along M/N, SBO = distance between 8-row groups along K.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);  // both operands MN-major
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_sub = p.ci_tile / 64;                   // x sub-tiles per stage
  const uint32_t stage_bytes = static_cast<uint32_t>(2 + n_sub) * kWgSub;

  int b = blockIdx.x;
  const int co_t = b % p.co_tiles;
  b /= p.co_tiles;
  const int ci_t = b % p.ci_tiles;
  b /= p.ci_tiles;
  const int tap = b;
  const int co0 = co_t * 128, ci0 = ci_t * p.ci_tile;
  const int total_tiles = p.tiles_w * p.tiles_h * p.n_img;
  const int per = (total_tiles + p.chunks - 1) / p.chunks;
  const int t_begin = blockIdx.y * per;
  const int t_end = min(total_tiles, t_begin + per);
  const int n_iters = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_dy);
    tma_prefetch_desc(&p.tmap_x[p.tap_map[tap]]);
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
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (n_iters <= 0) {
    // nothing to do for this chunk (more chunks than tiles): fall through to the common teardown
  } else if (warp == 0) {
    // converged warp, only the TMA issue under elect.sync (see conv_tc.cu: a divergent single-lane region turns every
    // uniform-datapath instruction into an ELECT / R2UR / BRA.U.ANY waterfall)
    const CUtensorMap* mx = &p.tmap_x[p.tap_map[tap]];
    RingPos rp;
    uint8_t* st = smem;
    int tile_w = t_begin % p.tiles_w, tile_h = (t_begin / p.tiles_w) % p.tiles_h, img = t_begin / (p.tiles_w * p.tiles_h);
    for (int it = 0; it < n_iters; ++it) {
      const int w0 = tile_w * p.tw, h0 = tile_h * p.th;
      mbar_wait(&empty_bar[rp.s], rp.phase ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[rp.s], stage_bytes);
        tma_load_4d(st, &p.tmap_dy, &full_bar[rp.s], co0, w0, h0, img);
        tma_load_4d(st + kWgSub, &p.tmap_dy, &full_bar[rp.s], co0 + 64, w0, h0, img);
        for (int j = 0; j < n_sub; ++j)
          tma_load_4d(st + static_cast<size_t>(2 + j) * kWgSub, mx, &full_bar[rp.s], ci0 + j * 64, w0 + p.tap_dw[tap],
                      h0 + p.tap_dh[tap], img);
      }
      __syncwarp();
      st += stage_bytes;
      rp.advance(p.stages);
      if (rp.s == 0) st = smem;
      if (++tile_w == p.tiles_w) {   // next spatial tile without integer division
        tile_w = 0;
        if (++tile_h == p.tiles_h) {
          tile_h = 0;
          ++img;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16_mn(128, static_cast<uint32_t>(p.ci_tile));
    const uint32_t s0 = smem_u32(smem);
    RingPos rp;
    uint32_t soff = 0;
    for (int it = 0; it < n_iters; ++it) {
      mbar_wait(&full_bar[rp.s], rp.phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = s0 + soff;
        const uint32_t sb = sa + 2 * kWgSub;
#pragma unroll
        for (int k = 0; k < kWgPix / 16; ++k) {
          // 16 pixels = two 8-row groups = 2048 B further down each sub-tile
          const uint64_t da = umma_desc_mnmajor_sw128(sa + k * 2048, kWgSub);
          const uint64_t db = umma_desc_mnmajor_sw128(sb + k * 2048, kWgSub);
          umma_f16_ss(tmem_base, da, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[rp.s]);
        if (it == n_iters - 1) umma_commit(&tmem_full_bar);
      }
      __syncwarp();
      soff += stage_bytes;
      rp.advance(p.stages);
      if (rp.s == 0) soff = 0;
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c = 0; c < p.ci_tile; c += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c, v);
      tmem_ld_wait();
      if (co < p.Cout) {
        float* row = p.dw + static_cast<long long>(co) * p.so + tap;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ci = ci0 + c + j;
          if (ci < p.Cin) atomicAdd(row + static_cast<long long>(ci) * p.si, __uint_as_float(v[j]) * p.inv_gscale);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int encode_tiled_generic(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_bytes);

static inline int floordiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }

int conv_wgrad_tc_supported(const fsb_conv_desc* d, int dy_cstride) {
  if (!(d->ksize == 1 || d->ksize == 3) || !(d->stride == 1 || d->stride == 2) || d->dil != 1) return 0;
  if (d->Cin < 16 || d->Cout < 16 || (d->x_cstride % 8) != 0 || (dy_cstride % 8) != 0) return 0;
  if (opt(OPT_WGRAD_TC) == 0) return 0;
  return 1;
}

int conv_wgrad_tc_launch(const fsb_conv_desc* d, const void* x, const void* dy, int dcs, float* dw, int64_t so, int64_t si,
                         float gscale, cudaStream_t stream) {
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15))
    return set_error(FSB_ERR_INVALID, "wgrad_tc: x / dy must be 16-byte aligned");
  WgradTcParams p;
  memset(&p, 0, sizeof(p));
  p.ksize = d->ksize;
  p.taps = d->ksize * d->ksize;
  p.tw = d->Wo >= 16 ? 16 : 8;
  p.th = kWgPix / p.tw;
  p.tiles_w = (d->Wo + p.tw - 1) / p.tw;
  p.tiles_h = (d->Ho + p.th - 1) / p.th;
  p.n_img = d->N;
  p.Cout = d->Cout;
  p.Cin = d->Cin;
  p.co_tiles = (d->Cout + 127) / 128;
  const int ci64 = (d->Cin + 63) / 64 * 64;
  p.ci_tile = ci64 > 256 ? (ci64 % 256 == 0 ? 256 : (ci64 % 192 == 0 ? 192 : (ci64 % 128 == 0 ? 128 : 64))) : ci64;
  p.ci_tiles = (ci64 + p.ci_tile - 1) / p.ci_tile;
  p.inv_gscale = 1.0f / gscale;
  p.dw = dw;
  p.so = so;
  p.si = si;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.ci_tile)) cols <<= 1;
  p.tmem_cols = cols;
  const int total_tiles = p.tiles_w * p.tiles_h * d->N;
  const int fixed = p.co_tiles * p.ci_tiles * p.taps;
  int chunks = (sm_count() * 2 + fixed - 1) / fixed;
  if (chunks > total_tiles) chunks = total_tiles;
  // deterministic mode (FSB_DETERMINISTIC=1 / fsb_set_option): no split over pixels, so every gradient element has exactly one
  // writer and the fp32 accumulation order is fixed (slower: co_tiles * ci_tiles * taps CTAs only)
  if (chunks < 1 || opt(OPT_DETERMINISTIC) == 1) chunks = 1;
  p.chunks = chunks;
  const size_t stage_bytes = static_cast<size_t>(2 + p.ci_tile / 64) * kWgSub;
  int stages = static_cast<int>((192 * 1024) / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  const int per = (total_tiles + chunks - 1) / chunks;
  if (stages > per) stages = per;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const size_t smem_bytes = stage_bytes * stages + 1024;

  const uint32_t box[4] = {64u, static_cast<uint32_t>(p.tw), static_cast<uint32_t>(p.th), 1u};
  {
    const uint64_t cs = static_cast<uint64_t>(dcs) * 2;
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(d->Wo), static_cast<uint64_t>(d->Ho),
                              static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {cs, cs * d->Wo, cs * d->Wo * d->Ho};
    int rc = encode_tiled_generic(&p.tmap_dy, dy, 4, dims, str, box, 128);
    if (rc) return rc;
  }
  const __half* xb = static_cast<const __half*>(x);
  const uint64_t cs = static_cast<uint64_t>(d->x_cstride) * 2;
  if (d->stride == 1) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {cs, cs * d->W, cs * d->W * d->H};
    int rc = encode_tiled_generic(&p.tmap_x[0], xb, 4, dims, str, box, 128);
    if (rc) return rc;
    for (int r = 0; r < d->ksize; ++r)
      for (int s = 0; s < d->ksize; ++s) {
        const int tp = r * d->ksize + s;
        p.tap_map[tp] = 0;
        p.tap_dh[tp] = r - d->pad + d->off_h;
        p.tap_dw[tp] = s - d->pad + d->off_w;
      }
  } else {
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < d->ksize; ++r)
      for (int s = 0; s < d->ksize; ++s) {
        const int tp = r * d->ksize + s;
        const int qh = r - d->pad + d->off_h, qw = s - d->pad + d->off_w;
        const int ph = ((qh % 2) + 2) % 2, pw = ((qw % 2) + 2) % 2;
        p.tap_map[tp] = ph * 2 + pw;
        p.tap_dh[tp] = floordiv2(qh);
        p.tap_dw[tp] = floordiv2(qw);
        used[ph * 2 + pw] = true;
      }
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        if (!used[ph * 2 + pw]) continue;
        const int Hp = (d->H - ph + 1) / 2, Wp = (d->W - pw + 1) / 2;
        if (Hp <= 0 || Wp <= 0) return set_error(FSB_ERR_INVALID, "wgrad_tc: empty parity plane");
        const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(Wp), static_cast<uint64_t>(Hp),
                                  static_cast<uint64_t>(d->N)};
        const uint64_t str[3] = {2 * cs, 2 * cs * d->W, cs * d->W * d->H};
        int rc = encode_tiled_generic(&p.tmap_x[ph * 2 + pw], xb + (static_cast<size_t>(ph) * d->W + pw) * d->x_cstride, 4, dims,
                                      str, box, 128);
        if (rc) return rc;
      }
  }
  cudaError_t e;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_wgrad_tc_kernel), 220 * 1024, "cudaFuncSetAttribute(conv_wgrad_tc)")) return rc;
  dim3 grid(static_cast<unsigned>(fixed), static_cast<unsigned>(chunks));
  e = launch_kernel(conv_wgrad_tc_kernel, grid, dim3(kWgThreads), smem_bytes, stream, p);
  if (e != cudaSuccess) return set_cuda_error(e, "conv_wgrad_tc launch");
  return FSB_OK;
}

}  // namespace fsb
