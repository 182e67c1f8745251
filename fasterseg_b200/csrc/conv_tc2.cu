// conv_tc2.cu -- K1 "row-strip" variant of the tcgen05 implicit-GEMM conv for 3x3 stride-1 convs on wide maps.
//
// conv_tc.cu re-reads the input once per filter tap (9 TMA boxes per 128-pixel tile) and streams the weights once per tile;
// ncu shows that layout saturating the L2->SM fabric (7.8 TB/s, ~70 % of the practical cap) long before HBM or the tensor
// pipe (profiles/r1_prof_conv_heads8_v1_summary.txt).  Here a CTA owns R consecutive output rows x 128 output columns:
//   * A (input): the R+2 input rows (130 pixels each: 128 + left/right halo) of one 64/32-channel chunk are loaded ONCE by
//     TMA into shared memory as K-major, 128B/64B-swizzled rows.  The operand of tap (r,s) for output row j is the SAME
//     buffer read through a UMMA descriptor whose start address is shifted by (j+r) rows and s pixels -- no data movement
//     per tap at all (halo re-read factor (R+2)/R instead of 9).
//   * B (weights): each (tap, chunk) tile is streamed once per CTA and reused by the R accumulators (R x fewer weight bytes
//     per output).
//   * D: R accumulators of 128 x N fp32 in TMEM (R*N <= 512 columns).
// Swizzle note (verified on B200, tests/test_kernels_gpu.py wide cases): TMA and the MMA unit both apply the 128B/64B XOR
// pattern to ABSOLUTE shared-memory address bits, so a descriptor that starts at a 128B (64B) multiple inside a
// 1024B-aligned buffer reads back exactly what TMA wrote; the descriptor's base_offset field must stay 0 (setting it to
// (addr >> 7) & 7 produces wrong results).
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

constexpr int kT2Threads = 192;
constexpr int kT2MaxB = 12;  // weight-tile ring depth
constexpr int kT2Cols = 128; // output columns per CTA (= UMMA M)
constexpr int kT2Halo = kT2Cols + 2;

struct ConvTc2Params {
  CUtensorMap tmap_a;  // {C, W, H, N}, box {BK, 130, 1, 1}
  CUtensorMap tmap_b;  // packed weights {Kpad, Npad, taps}, box {BK, n_tile, 1}
  int R;               // output rows per CTA
  int k_chunks;
  int Ho, Wo;
  int strips, row_groups;
  int Cout, n_tile;
  int a_stages, b_stages;
  int rowb;            // bytes per row buffer (1024-aligned)
  int y_cstride;
  uint32_t flags;
  uint32_t tmem_cols;
  const float* scale;
  const float* shift;
  __half* y;
  float* stats;   // partial rows, one per CTA (bn.cu "Deterministic statistics")
  int stats_C, stats_off;
  unsigned long long* dbg;  // optional timeline buffer (fsb_debug_set_buffer): 64 stamps per traced CTA
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define T2_STAMP(i) do { if (p.dbg && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) p.dbg[(blockIdx.x ? 64 : 0) + (i)] = gtimer(); } while (0)

template <int BK>
__global__ void __launch_bounds__(kT2Threads, 1)
conv_tc2_kernel(const __grid_constant__ ConvTc2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[2];
  __shared__ __align__(8) uint64_t a_empty[2];
  __shared__ __align__(8) uint64_t b_full[kT2MaxB];
  __shared__ __align__(8) uint64_t b_empty[kT2MaxB];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_scale[256];
  __shared__ float s_shift[256];

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) T2_STAMP(0);
  constexpr int kPix = BK * 2;  // bytes per pixel in a row buffer
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int rows_in = p.R + 2;
  const uint32_t a_stage_bytes = static_cast<uint32_t>(rows_in) * p.rowb;
  uint8_t* smem_b = smem + static_cast<size_t>(p.a_stages) * a_stage_bytes;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_tile) * kPix;

  int t = blockIdx.x;
  const int strip = t % p.strips;
  t /= p.strips;
  const int rg = t % p.row_groups;
  const int img = t / p.row_groups;
  const int w0 = strip * kT2Cols;
  const int h0 = rg * p.R;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_b);
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();
  for (int c = threadIdx.x; c < p.n_tile; c += kT2Threads) {
    const bool ok = c < p.Cout;
    s_scale[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.scale) ? p.scale[c] : 1.0f;
    s_shift[c] = (ok && (p.flags & FSB_CONV_AFFINE) && p.shift) ? p.shift[c] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) T2_STAMP(1);

  if (warp == 0) {
    // ================= TMA producer (converged warp; only the TMA issue is under elect.sync, see conv_tc.cu) =================
    RingPos ra, rb;
    for (int kc = 0; kc < p.k_chunks; ++kc) {
      mbar_wait(&a_empty[ra.s], ra.phase ^ 1u);
      uint8_t* sa = smem + static_cast<size_t>(ra.s) * a_stage_bytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&a_full[ra.s], static_cast<uint32_t>(rows_in) * kT2Halo * kPix);
        for (int i = 0; i < rows_in; ++i)
          tma_load_4d(sa + static_cast<size_t>(i) * p.rowb, &p.tmap_a, &a_full[ra.s], kc * BK, w0 - 1, h0 - 1 + i, img);
      }
      __syncwarp();
      ra.advance(p.a_stages);
      for (int tap = 0; tap < 9; ++tap) {
        mbar_wait(&b_empty[rb.s], rb.phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[rb.s], b_bytes);
          tma_load_3d(smem_b + static_cast<size_t>(rb.s) * b_bytes, &p.tmap_b, &b_full[rb.s], kc * BK, 0, tap);
        }
        __syncwarp();
        rb.advance(p.b_stages);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (converged warp, elected lane issues, division-free rings) =================
    const uint32_t idesc = umma_idesc_f16(128, static_cast<uint32_t>(p.n_tile));
    const uint64_t db0 = umma_desc_kmajor(smem_u32(smem_b), kPix);
    const uint32_t bstep = b_bytes >> 4;
    RingPos ra, rb;
    for (int kc = 0; kc < p.k_chunks; ++kc) {
      mbar_wait(&a_full[ra.s], ra.phase);
      tc_fence_after();
      if (lane == 0) T2_STAMP(2 + kc * 10);
      const uint32_t sa = smem_u32(smem + static_cast<size_t>(ra.s) * a_stage_bytes);
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          mbar_wait(&b_full[rb.s], rb.phase);
          tc_fence_after();
          const uint64_t db = db0 + static_cast<uint32_t>(rb.s) * bstep;
          if (elect_one()) {
            uint32_t a_addr = sa + static_cast<uint32_t>(r) * p.rowb + static_cast<uint32_t>(s) * kPix;
            uint32_t tm = tmem_base;
            for (int j = 0; j < p.R; ++j, a_addr += p.rowb, tm += p.n_tile) {
              const uint64_t da = umma_desc_kmajor(a_addr, kPix);
              umma_f16_ss(tm, da, db, idesc, (kc > 0 || r > 0 || s > 0) ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < BK / 16; ++k)
                umma_f16_ss(tm, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, 1u);
            }
            umma_commit(&b_empty[rb.s]);
            if (r == 2 && s == 2) {
              umma_commit(&a_empty[ra.s]);
              if (kc == p.k_chunks - 1) umma_commit(&tmem_full_bar);
            }
          }
          __syncwarp();
          rb.advance(p.b_stages);
        }
      ra.advance(p.a_stages);
    }
    if (lane == 0) T2_STAMP(40);
  } else {
    // ================= epilogue warps 2..5 =================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int ow = w0 + m;
    const bool relu = (p.flags & FSB_CONV_RELU) != 0;
    const bool out_f32 = (p.flags & FSB_CONV_OUT_F32) != 0;
    const bool do_stats = (p.flags & FSB_CONV_STATS) != 0 && p.stats != nullptr;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) T2_STAMP(41);
    // per-warp channel statistics [4 warps][sum | sumsq][n_tile], accumulated over the R rows in row order (lane 0 only)
    float* s_stat = reinterpret_cast<float*>(smem);
    if (do_stats) {
      for (int i = lane; i < 2 * p.n_tile; i += 32) s_stat[q * 2 * p.n_tile + i] = 0.f;
      __syncwarp();
    }
    for (int j = 0; j < p.R; ++j) {
      const int oh = h0 + j;
      const bool pix_ok = (oh < p.Ho) && (ow < p.Wo);
      const size_t pix = (static_cast<size_t>(img) * p.Ho + oh) * p.Wo + ow;
      __half* yrow = p.y + pix * p.y_cstride;
      float* yrow32 = reinterpret_cast<float*>(p.y) + pix * p.y_cstride;
      const bool vec_ok = out_f32 ? ((reinterpret_cast<uintptr_t>(yrow32) & 15) == 0) : ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * p.n_tile);
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
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (do_stats) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = pix_ok ? f[i] : 0.f;
            float b = a * a;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane == 0) {
              s_stat[(q * 2 + 0) * p.n_tile + c + i] += a;
              s_stat[(q * 2 + 1) * p.n_tile + c + i] += b;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x = f[i] * s_scale[c + i] + s_shift[c + i];
          f[i] = relu ? fmaxf(x, 0.f) : x;
        }
        if (!pix_ok) continue;
        const int remaining = p.Cout - c;
        if (out_f32) {
          if (remaining >= 16 && vec_ok) {
            float4* dst = reinterpret_cast<float4*>(yrow32 + c);
            dst[0] = make_float4(f[0], f[1], f[2], f[3]);
            dst[1] = make_float4(f[4], f[5], f[6], f[7]);
            dst[2] = make_float4(f[8], f[9], f[10], f[11]);
            dst[3] = make_float4(f[12], f[13], f[14], f[15]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < remaining) yrow32[c + i] = f[i];
          }
        } else if (remaining >= 16 && vec_ok) {
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
          for (int i = 0; i < 16; ++i)
            if (i < remaining) yrow[c + i] = __float2half_rn(f[i]);
        }
        }
      }
    }
    if (do_stats) {
      named_bar_sync(2, 128);
      float* row = p.stats + static_cast<size_t>(blockIdx.x) * 2 * p.stats_C + p.stats_off;
      for (int ch = static_cast<int>(threadIdx.x) - 64; ch < p.n_tile; ch += 128) {
        if (ch >= p.Cout) continue;
        float a = s_stat[0 * p.n_tile + ch], b = s_stat[1 * p.n_tile + ch];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          a += s_stat[(w * 2 + 0) * p.n_tile + ch];
          b += s_stat[(w * 2 + 1) * p.n_tile + ch];
        }
        row[ch] = a;
        row[p.stats_C + ch] = b;
      }
    }
    tc_fence_before();
    if (threadIdx.x == 64) T2_STAMP(42);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

unsigned long long* g_dbg_buffer = nullptr;

// ------------------------------------------------------------------------------------------
int encode_tiled_generic(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, int swizzle_bytes);

int conv_tc2_supported(const fsb_conv_desc* d) {
  if (d->ksize != 3 || d->stride != 1 || d->dil != 1 || d->pad != 1 || d->off_h || d->off_w) return 0;
  if (d->Cin < 32 || (d->x_cstride % 8) != 0 || d->Cout > 256) return 0;
  if (d->Wo < 96) return 0;  // strips are 128 output columns wide
  const int mode = opt(OPT_CONV_TC2);
  if (mode == 0) return 0;
  if (mode == 2) return 1;  // force (tests / tuning)
  // Measured on B200 (profiles/r1_conv_bench_v3.log): with 1-CTA SS-mode MMAs both kernels are bound by the tensor pipe's
  // operand fetch (~130 cycles per 128x128x16 MMA); the row-strip layout only wins where the per-tap kernel's L2->SM
  // traffic is the limiter: narrow N (<= 64 output channels) on maps large enough for >= 1 wave of 8-row strips.
  if (d->Cout > 64 || d->Cin % 64 != 0) return 0;
  if (static_cast<int64_t>(d->N) * d->Ho * d->Wo < 98304) return 0;
  return 1;
}

// rows per CTA: as many as TMEM (R * N <= 512) and shared memory allow while still producing >= ~0.8 waves of CTAs
int conv_tc2_rows_per_cta(const fsb_conv_desc* d) {
  const ConvGeom g = conv_geom(d);
  const int n_tile = g.npad;
  const int strips = (d->Wo + kT2Cols - 1) / kT2Cols;
  const int pix_bytes = g.bk * 2;
  const int rowb = (kT2Halo * pix_bytes + 1023) / 1024 * 1024;
  const int b_bytes = n_tile * pix_bytes;
  const int sms = sm_count();
  const int smem_cap = 200 * 1024;
  int bestR = 1;
  for (int R = 8; R >= 1; R >>= 1) {
    if (R * n_tile > 512) continue;
    const int a_stage = (R + 2) * rowb;
    if (a_stage + 2 * b_bytes > smem_cap) continue;
    const int tiles = ((d->Ho + R - 1) / R) * strips * d->N;
    bestR = R;
    if (tiles * 10 >= sms * 8) break;  // enough CTAs: stop shrinking R
  }
  const int R = opt(OPT_TC2_R);  // tuning override
  if (R >= 1 && R <= 8 && R * n_tile <= 512 && (R + 2) * rowb + 2 * b_bytes <= smem_cap) bestR = R;
  return bestR;
}
// CTAs (= partial statistic rows) of the row-strip kernel
int conv_tc2_ctas(const fsb_conv_desc* d) {
  const int R = conv_tc2_rows_per_cta(d);
  return ((d->Wo + kT2Cols - 1) / kT2Cols) * ((d->Ho + R - 1) / R) * d->N;
}

int conv_tc2_launch(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift, void* y,
                    float* stats, cudaStream_t stream) {
  const ConvGeom g = conv_geom(d);
  ConvTc2Params p;
  memset(&p, 0, sizeof(p));
  p.k_chunks = g.kpad / g.bk;
  p.Ho = d->Ho;
  p.Wo = d->Wo;
  p.Cout = d->Cout;
  p.n_tile = g.npad;  // single N tile (Cout <= 256)
  p.strips = (d->Wo + kT2Cols - 1) / kT2Cols;
  const int pix_bytes = g.bk * 2;
  p.rowb = (kT2Halo * pix_bytes + 1023) / 1024 * 1024;
  p.y_cstride = d->y_cstride;
  p.flags = d->flags;
  p.scale = scale;
  p.shift = shift;
  p.y = static_cast<__half*>(y);
  p.stats = stats;
  p.stats_C = d->stats_C > 0 ? d->stats_C : d->Cout;
  p.stats_off = d->stats_off;
  p.dbg = g_dbg_buffer;
  const int b_bytes = p.n_tile * pix_bytes;
  const int smem_cap = 200 * 1024;
  const int bestR = conv_tc2_rows_per_cta(d);
  p.R = bestR;
  p.row_groups = (d->Ho + p.R - 1) / p.R;
  const int a_stage = (p.R + 2) * p.rowb;
  // one A stage + a deep weight ring beats two A stages + a shallow ring: the weight tiles are the latency-critical stream
  const int want_a = opt(OPT_TC2_ASTAGES) == 2 ? 2 : 1;
  p.a_stages = (want_a == 2 && p.k_chunks > 1 && 2 * a_stage + 2 * b_bytes <= smem_cap) ? 2 : 1;
  int bst = (smem_cap - p.a_stages * a_stage) / b_bytes;
  if (bst > kT2MaxB) bst = kT2MaxB;
  if (bst < 2) bst = 2;
  p.b_stages = bst;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.R * p.n_tile)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem_bytes = static_cast<size_t>(p.a_stages) * a_stage + static_cast<size_t>(p.b_stages) * b_bytes + 1024;

  const uint64_t cs = static_cast<uint64_t>(d->x_cstride) * 2;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->N)};
    const uint64_t str[3] = {cs, cs * d->W, cs * d->W * d->H};
    const uint32_t box[4] = {static_cast<uint32_t>(g.bk), static_cast<uint32_t>(kT2Halo), 1u, 1u};
    int rc = encode_tiled_generic(&p.tmap_a, x, 4, dims, str, box, pix_bytes);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(g.kpad), static_cast<uint64_t>(g.npad), 9ull};
    const uint64_t str[2] = {static_cast<uint64_t>(g.kpad) * 2, static_cast<uint64_t>(g.kpad) * g.npad * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(g.bk), static_cast<uint32_t>(p.n_tile), 1u};
    int rc = encode_tiled_generic(&p.tmap_b, wpacked, 3, dims, str, box, pix_bytes);
    if (rc) return rc;
  }
  dim3 grid(static_cast<unsigned>(p.strips * p.row_groups * d->N));
  cudaError_t e;
  if (g.bk == 64) {
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc2_kernel<64>), 220 * 1024, "cudaFuncSetAttribute(conv_tc2<64>)")) return rc;
    e = launch_kernel(conv_tc2_kernel<64>, grid, dim3(kT2Threads), smem_bytes, stream, p);
  } else {
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc2_kernel<32>), 220 * 1024, "cudaFuncSetAttribute(conv_tc2<32>)")) return rc;
    e = launch_kernel(conv_tc2_kernel<32>, grid, dim3(kT2Threads), smem_bytes, stream, p);
  }
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc2 launch");
  return FSB_OK;
}

}  // namespace fsb
