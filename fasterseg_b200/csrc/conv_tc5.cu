// conv_tc5.cu -- K1n: "tap-concatenated" tcgen05 implicit-GEMM conv for 3x3 stride-1 layers with FEW output channels
// (Cout <= 64: stem.1.conv2 / stem.2.conv2 64->64, cell convs 64->32 of the student; reference call sites
// search/operations.py:72-83,146-150, train/model_seg.py:193-200).
//
// Measured on B200 (tools/umma_rate.cu, profiles/r2_umma_rate_*.log): one tcgen05.mma kind::f16 costs ~140-180 cycles per
// SM whether N is 64, 128 or 256 -- the instruction streams its 128 A rows at about one row per cycle and N only decides how
// much of the array is used.  A pixel-major conv tile with N = Cout = 64 therefore runs the tensor pipe at <= 25 % however the
// operands are staged (conv_tc2: 36 MMAs per 128 output pixels).  Here the three HORIZONTAL taps of a kernel row share one
// instruction: for kernel row r the B operand is the 3*Cout x 64 stack [W(r,0); W(r,1); W(r,2)] and the A operand is the
// UN-shifted input row, so
//      D[p, (s, c)] = sum_{r, ci} X[pixel p of input row y+r-1, ci] * W[c, ci, r, s]          (12 MMAs instead of 36 for Cin = 64)
// and the horizontal shift moves into the epilogue:  y[p, c] = D[p-1, (0,c)] + D[p, (1,c)] + D[p+1, (2,c)]
// (thread = pixel = TMEM lane; the two neighbour terms come from the adjacent lanes by warp shuffle).  Pixels are addressed on
// the FLATTENED H*W axis of the NHWC tensor.  A tile is four groups of 32 consecutive pixels, one group per TMEM lane quarter
// = per epilogue warp, consecutive groups overlapping by two pixels: every warp holds its own halo pixel on either side, so
// the shifted sum needs NO data from another warp (a first version that exchanged the border terms through shared memory
// spent 8 us per tile in divergent edge code and named barriers; profiles/r2_tc5_timeline_v1.log).  120 outputs per tile.
// Kernel row r is the same window shifted by (r-1)*W pixels; TMA zero-fills what falls off either end of the image (= top /
// bottom padding) and the epilogue drops the left/right neighbour term of pixels in the first/last image column (= left /
// right padding), so tiles need not align with image rows.
// Persistent CTAs (one per SM) walk tiles; all 3 * Cin/64 weight stacks stay resident in shared memory; two TMEM accumulators
// and TWO sets of four epilogue warps (even / odd tiles, each warp with its own staging slab and TMA store, software-pipelined
// tcgen05.ld) keep the epilogue off the critical path.  Issue loops follow conv_tc.cu (converged warps, elect.sync, division-free rings).
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int k5Threads = 320;     // warp 0 TMA producer, warp 1 TMEM + MMA issuer, warps 2..5 / 6..9 epilogue of even / odd tiles
constexpr int k5MaxStages = 10;
constexpr int k5Lanes = 128;       // TMEM lanes = rows of the A operand
constexpr int k5GroupOut = 30;     // outputs per 32-pixel group (one halo pixel on either side)
constexpr int k5Out = 4 * k5GroupOut;   // outputs per tile
constexpr uint32_t k5StageSlab = 4096;  // per-warp staging slab (30 rows x <= 128 B, 1024-aligned)
constexpr uint32_t k5ABytes = k5Lanes * 128;   // one (kernel row, 64-channel chunk) slab of the input window

struct ConvTc5Params {
  CUtensorMap tmap_x;   // {Cin, H*W, N}, box {64, 32, 1}, SW128
  CUtensorMap tmap_w;   // packed weights {Kpad, Npad, 9}, box {64, Cout, 3}, SW128
  CUtensorMap tmap_y;   // {Cout, H*W, N}, box {Cout, 30, 1}, SW128 when Cout == 64 else dense rows
  int kch;              // Cin / 64
  int W, P;             // image width, pixels per image (H * W)
  int tiles_per_img, tiles;
  int Cout, NT;         // NT = 3 * Cout (MMA N)
  int stages;
  int y_swizzled;       // staging rows are 128 B and XOR-swizzled (Cout == 64)
  uint32_t flags, tmem_cols, acc_cols;
  const float* scale;
  const float* shift;
  unsigned long long* dbg;   // optional timeline buffer (fsb_debug_set_buffer): 64 stamps of CTA 0
};

extern unsigned long long* g_dbg_buffer;
__device__ __forceinline__ unsigned long long gtimer5() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define T5_STAMP(i) do { if (p.dbg && blockIdx.x == 0 && (i) < 64) p.dbg[(i)] = gtimer5(); } while (0)

// One warp drains its 32 lanes of an accumulator: CH channels per pass (three tcgen05.ld of CH columns, ONE wait), the shifted
// three-term sum by warp shuffle, BN scale/shift, ReLU, fp16, into the warp's staging slab (rows = output pixels).
template <int CH>
__device__ __forceinline__ void tc5_drain(const ConvTc5Params& p, uint32_t taddr, uint8_t* slab, int lane, float keep0, float keep2,
                                          bool relu, const float* s_scale, const float* s_shift, uint64_t* empty_bar) {
  const uint32_t row_bytes = static_cast<uint32_t>(p.Cout) * 2u;
  for (int c0 = 0; c0 < p.Cout; c0 += CH) {
    uint32_t v0[CH], v1[CH], v2[CH];
    if constexpr (CH == 32) {
      tmem_ld32(taddr + c0, v0);
      tmem_ld32(taddr + p.Cout + c0, v1);
      tmem_ld32(taddr + 2 * p.Cout + c0, v2);
    } else {
      tmem_ld16(taddr + c0, v0);
      tmem_ld16(taddr + p.Cout + c0, v1);
      tmem_ld16(taddr + 2 * p.Cout + c0, v2);
    }
    tmem_ld_wait();
    if (c0 + CH >= p.Cout) {   // that was this warp's last TMEM read of the tile: hand the accumulator back before the arithmetic
      tc_fence_before();
      if (lane == 0) mbar_arrive(empty_bar);
    }
#pragma unroll
    for (int h = 0; h < CH / 16; ++h) {
      float f[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int c = h * 16 + e;
        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[c]), 1);
        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[c]), 1);
        const float acc = fmaf(left, keep0, fmaf(right, keep2, __uint_as_float(v1[c])));
        const float x = fmaf(acc, s_scale[c0 + c], s_shift[c0 + c]);
        f[e] = relu ? fmaxf(x, 0.f) : x;
      }
      if (lane >= 1 && lane <= k5GroupOut) {
        uint4 o0, o1;
        o0.x = pack_half2(f[0], f[1]);
        o0.y = pack_half2(f[2], f[3]);
        o0.z = pack_half2(f[4], f[5]);
        o0.w = pack_half2(f[6], f[7]);
        o1.x = pack_half2(f[8], f[9]);
        o1.y = pack_half2(f[10], f[11]);
        o1.z = pack_half2(f[12], f[13]);
        o1.w = pack_half2(f[14], f[15]);
        const int row = lane - 1;
        uint8_t* dst = slab + static_cast<size_t>(row) * row_bytes;
        const int ch16 = (c0 >> 3) + 2 * h;
        if (p.y_swizzled) {
          *reinterpret_cast<uint4*>(dst + ((ch16 ^ (row & 7)) << 4)) = o0;
          *reinterpret_cast<uint4*>(dst + (((ch16 + 1) ^ (row & 7)) << 4)) = o1;
        } else {
          *reinterpret_cast<uint4*>(dst + (ch16 << 4)) = o0;
          *reinterpret_cast<uint4*>(dst + ((ch16 + 1) << 4)) = o1;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(k5Threads, 1)
conv_tc5_kernel(const __grid_constant__ ConvTc5Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[k5MaxStages];
  __shared__ __align__(8) uint64_t a_empty[k5MaxStages];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_scale[64];
  __shared__ float s_shift[64];

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) T5_STAMP(0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t w_tile = static_cast<uint32_t>(p.NT) * 128u;            // one (kernel row, chunk) weight stack
  const uint32_t w_bytes = 3u * static_cast<uint32_t>(p.kch) * w_tile;
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + w_bytes;
  uint8_t* staging = smem_a + static_cast<size_t>(p.stages) * k5ABytes;
  const int k_steps = 3 * p.kch;   // (kernel row, chunk) steps per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_y);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(&w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 4);   // one arrival per epilogue warp of the set that drains this accumulator
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();
  for (int c = threadIdx.x; c < 64; c += k5Threads) {
    const bool ok = c < p.Cout;
    s_scale[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.scale) ? p.scale[c] : 1.0f;
    s_shift[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.shift) ? p.shift[c] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) T5_STAMP(1);

  if (warp == 0) {
    // ================= TMA producer: resident weights once, then the input windows of every tile of this CTA =================
    if (elect_one()) {
      mbar_arrive_expect_tx(&w_full, w_bytes);
      for (int r = 0; r < 3; ++r)
        for (int kc = 0; kc < p.kch; ++kc)
          tma_load_3d(smem_w + static_cast<size_t>(r * p.kch + kc) * w_tile, &p.tmap_w, &w_full, kc * 64, 0, r * 3);
    }
    __syncwarp();
    RingPos rp;
    uint8_t* sa = smem_a;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int img = tile / p.tiles_per_img;
      const int i0 = (tile - img * p.tiles_per_img) * k5Out;
      for (int r = 0; r < 3; ++r) {
        const int pix = i0 - 1 + (r - 1) * p.W;   // flattened coordinate of lane 0's pixel for this kernel row
        for (int kc = 0; kc < p.kch; ++kc) {
          mbar_wait(&a_empty[rp.s], rp.phase ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[rp.s], k5ABytes);
#pragma unroll
            for (int g = 0; g < 4; ++g)   // lane quarter g holds pixels pix + 30 g .. + 31
              tma_load_3d(sa + g * 4096, &p.tmap_x, &a_full[rp.s], kc * 64, pix + g * k5GroupOut, img);
          }
          __syncwarp();
          sa += k5ABytes;
          rp.advance(p.stages);
          if (rp.s == 0) sa = smem_a;
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: alternates between the two TMEM accumulators =================
    const uint32_t idesc = umma_idesc_f16(128, static_cast<uint32_t>(p.NT));
    const uint64_t da0 = umma_desc_kmajor(smem_u32(smem_a), 128);
    const uint64_t db0 = umma_desc_kmajor(smem_u32(smem_w), 128);
    const uint32_t wstep = w_tile >> 4;
    mbar_wait(&w_full, 0);
    tc_fence_after();
    RingPos rp;
    uint32_t aoff = 0;
    int buf = 0;
    uint32_t acc_phase = 0;   // parity of the accumulator round: tmem_empty is waited with (acc_phase ^ 1) like a producer ring
    int lt_m = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt_m) {
      mbar_wait(&tmem_empty[buf], acc_phase ^ 1u);   // epilogue of the tile that used this accumulator two tiles ago
      tc_fence_after();
      if (lane == 0) T5_STAMP(2 + lt_m * 8 + 0);
      const uint32_t acc = tmem_base + static_cast<uint32_t>(buf) * p.acc_cols;
      uint32_t woff = 0;
      for (int ks = 0; ks < k_steps; ++ks) {
        mbar_wait(&a_full[rp.s], rp.phase);
        tc_fence_after();
        if (lane == 0 && ks == 0) T5_STAMP(2 + lt_m * 8 + 1);
        if (elect_one()) {
          const uint64_t da = da0 + aoff, db = db0 + woff;
          umma_f16_ss(acc, da, db, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_f16_ss(acc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, 1u);
          umma_commit(&a_empty[rp.s]);
          if (ks == k_steps - 1) umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
        woff += wstep;
        aoff += k5ABytes >> 4;
        rp.advance(p.stages);
        if (rp.s == 0) aoff = 0;
      }
      if (lane == 0) T5_STAMP(2 + lt_m * 8 + 2);
      buf ^= 1;
      if (buf == 0) acc_phase ^= 1u;
    }
  } else {
    // ================= epilogue: warps 2..5 take the even tiles of this CTA, warps 6..9 the odd ones; thread = pixel lane ===========
    // out[l] = D[l-1](s=0) + D[l](s=1) + D[l+1](s=2) inside the warp's own 32 lanes (lanes 0 and 31 are the group's halo pixels).
    // (Splitting the channel chunks of ONE tile over both warp sets was measured slower: 2.0 vs 1.4 us per tile -- eight warps
    // reading the accumulator that the tensor pipe is about to need contend for TMEM; profiles/r2_tc5_timeline_v3.log.)
    const int set = (warp - 2) >> 2;   // 0: even local tiles (accumulator 0), 1: odd local tiles (accumulator 1)
    const int q = warp & 3;            // TMEM lane quarter
    const bool relu = (p.flags & FSB_CONV_RELU) != 0;
    uint8_t* slab = staging + static_cast<size_t>(set * 4 + q) * k5StageSlab;
    const uint32_t taddr = tmem_base + static_cast<uint32_t>(set) * p.acc_cols + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t acc_phase = 0;
    int lt = set;
    for (int tile = blockIdx.x + set * gridDim.x; tile < p.tiles; tile += 2 * gridDim.x, lt += 2) {
      const int img = tile / p.tiles_per_img;
      const int i0 = (tile - img * p.tiles_per_img) * k5Out;
      const int g0 = i0 + q * k5GroupOut;         // first output pixel of this warp's group
      const int idx = g0 + lane - 1;              // flattened pixel of this lane
      const int xcol = idx >= 0 ? idx % p.W : 0;
      const float keep0 = xcol == 0 ? 0.f : 1.f;          // no left neighbour in the first image column
      const float keep2 = xcol == p.W - 1 ? 0.f : 1.f;    // no right neighbour in the last image column
      mbar_wait(&tmem_full[set], acc_phase);
      tc_fence_after();
      if (threadIdx.x == 64) T5_STAMP(2 + lt * 8 + 3);
      if (lt >= 2) {   // this warp's slab still feeds its previous bulk store until that store has READ it
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      if ((p.Cout & 31) == 0)
        tc5_drain<32>(p, taddr, slab, lane, keep0, keep2, relu, s_scale, s_shift, &tmem_empty[set]);
      else
        tc5_drain<16>(p, taddr, slab, lane, keep0, keep2, relu, s_scale, s_shift, &tmem_empty[set]);
      if (threadIdx.x == 64) T5_STAMP(2 + lt * 8 + 4);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && g0 < p.P) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&p.tmap_y)), "r"(smem_u32(slab)), "r"(0), "r"(g0), "r"(img) : "memory");
        tma_store_commit();
      }
      if (threadIdx.x == 64) T5_STAMP(2 + lt * 8 + 5);
      acc_phase ^= 1u;
    }
    if (lane == 0) tma_store_wait_read();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
int encode_tiled_generic(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_bytes);

struct Tc5Plan {
  int ok, kch, stages, tiles_per_img, tiles;
  uint32_t tmem_cols, acc_cols;
  size_t smem_bytes;
};

static Tc5Plan conv_tc5_plan(const fsb_conv_desc* d) {
  Tc5Plan q;
  memset(&q, 0, sizeof(q));
  q.kch = d->Cin / 64;
  const int NT = 3 * d->Cout;
  const size_t w_bytes = static_cast<size_t>(3) * q.kch * NT * 128;
  const size_t staging = 8 * k5StageSlab;
  const size_t budget = 220 * 1024 - 1024;
  if (w_bytes + staging + 3 * k5ABytes > budget) return q;
  int st = static_cast<int>((budget - w_bytes - staging) / k5ABytes);
  if (st > k5MaxStages) st = k5MaxStages;
  q.stages = st;
  const int64_t P = static_cast<int64_t>(d->H) * d->W;
  q.tiles_per_img = static_cast<int>((P + k5Out - 1) / k5Out);
  q.tiles = q.tiles_per_img * d->N;
  uint32_t acc = 32;
  while (acc < static_cast<uint32_t>(NT)) acc <<= 1;
  q.acc_cols = acc;
  q.tmem_cols = 2 * acc;
  if (q.tmem_cols > 512) return q;
  q.smem_bytes = w_bytes + static_cast<size_t>(st) * k5ABytes + staging + 1024;
  q.ok = 1;
  return q;
}

int conv_tc5_supported(const fsb_conv_desc* d, const void* y) {
  const int mode = opt(OPT_CONV_TC5);
  if (mode == 0) return 0;
  if (d->ksize != 3 || d->stride != 1 || d->dil != 1 || d->pad != 1 || d->off_h || d->off_w) return 0;
  if (d->Cin % 64 != 0 || d->Cin > 256) return 0;
  if (!(d->Cout == 16 || d->Cout == 32 || d->Cout == 48 || d->Cout == 64)) return 0;
  if ((d->x_cstride % 8) != 0 || (d->y_cstride % 8) != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0) return 0;
  if (d->flags & (FSB_CONV_OUT_F32 | FSB_CONV_STATS | FSB_CONV_FORCE_DIRECT)) return 0;
  if (static_cast<int64_t>(d->H) * d->W >= (1ll << 31)) return 0;
  if (mode != 2 && static_cast<int64_t>(d->N) * d->H * d->W < k5Out * 96) return 0;   // fewer tiles than ~2/3 of the SMs
  return conv_tc5_plan(d).ok;
}

int conv_tc5_launch(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift, void* y,
                    cudaStream_t stream) {
  const ConvGeom g = conv_geom(d);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(wpacked) & 15))
    return set_error(FSB_ERR_INVALID, "conv_tc5: x / wpacked must be 16-byte aligned");
  const Tc5Plan q = conv_tc5_plan(d);
  if (!q.ok) return set_error(FSB_ERR_UNSUPPORTED, "conv_tc5: no shared-memory plan for this shape");
  ConvTc5Params p;
  memset(&p, 0, sizeof(p));
  p.kch = q.kch;
  p.W = d->W;
  p.P = d->H * d->W;
  p.tiles_per_img = q.tiles_per_img;
  p.tiles = q.tiles;
  p.Cout = d->Cout;
  p.NT = 3 * d->Cout;
  p.stages = q.stages;
  p.y_swizzled = d->Cout == 64 ? 1 : 0;
  p.flags = d->flags;
  p.tmem_cols = q.tmem_cols;
  p.acc_cols = q.acc_cols;
  p.scale = scale;
  p.shift = shift;
  p.dbg = g_dbg_buffer;
  const uint64_t P = static_cast<uint64_t>(d->H) * d->W;
  {
    const uint64_t cs = static_cast<uint64_t>(d->x_cstride) * 2;
    const uint64_t dims[3] = {static_cast<uint64_t>(d->Cin), P, static_cast<uint64_t>(d->N)};
    const uint64_t str[2] = {cs, cs * P};
    const uint32_t box[3] = {64u, 32u, 1u};
    if (int rc = encode_tiled_generic(&p.tmap_x, x, 3, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(g.kpad), static_cast<uint64_t>(g.npad), 9ull};
    const uint64_t str[2] = {static_cast<uint64_t>(g.kpad) * 2, static_cast<uint64_t>(g.kpad) * g.npad * 2};
    const uint32_t box[3] = {64u, static_cast<uint32_t>(d->Cout), 3u};
    if (int rc = encode_tiled_generic(&p.tmap_w, wpacked, 3, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t ycs = static_cast<uint64_t>(d->y_cstride) * 2;
    const uint64_t dims[3] = {static_cast<uint64_t>(d->Cout), P, static_cast<uint64_t>(d->N)};
    const uint64_t str[2] = {ycs, ycs * P};
    const uint32_t box[3] = {static_cast<uint32_t>(d->Cout), static_cast<uint32_t>(k5GroupOut), 1u};
    if (int rc = encode_tiled_generic(&p.tmap_y, y, 3, dims, str, box, p.y_swizzled ? 128 : 0)) return rc;
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc5_kernel), 222 * 1024, "cudaFuncSetAttribute(conv_tc5)")) return rc;
  const int sms = sm_count();
  const unsigned ctas = static_cast<unsigned>(q.tiles < sms ? q.tiles : sms);
  cudaError_t e = launch_kernel(conv_tc5_kernel, dim3(ctas), dim3(k5Threads), q.smem_bytes, stream, p);
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc5 launch");
  return FSB_OK;
}

}  // namespace fsb
