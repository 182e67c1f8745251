// umma_rate.cu -- B200 micro-benchmark behind the K1 design: how many cycles does one
// tcgen05.mma kind::f16 (K = 16, fp16 x fp16 -> fp32, both operands from shared memory) take per SM as a function of
//   * the instruction shape (M = 128 with cta_group::1, M = 256 with cta_group::2; N in {64, 128, 256}),
//   * the shared-memory WRITE traffic that TMA adds while the tensor pipe reads its operands (per-stage A and/or B loads,
//     A every `a_every`-th stage = the row-strip reuse of the conv kernels).
// Issue style matters: a tcgen05.mma inside an `if (lane == 0)` region is compiled into a per-instruction "waterfall" (ELECT / R2UR /
// BRA.U.ANY loop, ~130-230 cycles per MMA -- the figure the first version of this benchmark and every kernel of round 1
// measured); here the issuing warp stays converged and only the instruction itself sits under elect.sync, so the operands
// live in uniform registers and consecutive UTCHMMAs are back to back in the SASS.
// One CTA (pair) per SM, synthetic mainloop with the same mbarrier ring a real kernel uses.  Standalone: built with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu -lcuda
// Output: one line per configuration with cycles per MMA (median over CTAs) and the implied TFLOP/s at the SM clock.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(tx) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  const long long t0 = clock64();
  while (!mbar_try(b, par)) {
    if (clock64() - t0 > 2000000000LL) { printf("umma_rate: barrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ bool elect() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= 1ull << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24); }

struct Params {
  CUtensorMap tm_a;   // [128 rows][64 fp16], box {64,128}
  CUtensorMap tm_b;   // [256 rows][64 fp16], box {64, nb_rows}
  int N;              // instruction N
  int stages;
  int kblocks;        // pipeline iterations; 4 MMAs each
  int load_a;         // 0: never, n>0: every n-th k-block
  int load_b;         // 0/1
  int nacc;           // accumulators used round-robin by consecutive MMAs (1 = one dependent chain)
  int commit_every;   // no-load mode only: tcgen05.commit to a barrier nobody waits on every n MMAs (0 = never)
  int kblock;         // MMAs per pipeline stage (4 = one 64-wide k-block; 8 re-reads the stage twice)
  int M;              // instruction M per CTA: 128 or 64
  int flags;          // 1: no tcgen05.fence::after_thread_sync after the full-barrier wait; 2: handshake only (producer arrives, no TMA)
  long long* cycles;  // per CTA
};

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t full[8];
  __shared__ __align__(8) uint64_t empty[8];
  __shared__ __align__(8) uint64_t done;
  __shared__ __align__(8) uint64_t dummy;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_rank() : 0;
  const int nb_rows = p.N / CG;                    // B rows held by this CTA
  const uint32_t a_bytes = 128 * 128, b_bytes = static_cast<uint32_t>(nb_rows) * 128;
  const uint32_t stage_bytes = a_bytes + 256 * 128;
  // fill the operand buffers with small finite numbers (the loads may be switched off)
  for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(p.stages) * stage_bytes / 2; i += blockDim.x)
    reinterpret_cast<__half*>(smem)[i] = __float2half(0.001f * static_cast<float>((i * 37u) & 255u) - 0.1f);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    mbar_init(&dummy, 1 << 20);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const bool loads = p.load_a || p.load_b;

  if (warp == 0 && loads) {
    // ---------------- producer (every CTA loads its own operands; completion lands on the leader's barrier) ----------
    for (int kb = 0; kb < p.kblocks; ++kb) {
      const int s = kb % p.stages, round = kb / p.stages;
      if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
      const bool hs = (p.flags & 2) != 0;
      const bool la = !hs && p.load_a && (kb % p.load_a) == 0;
      const bool lbb = !hs && p.load_b;
      const uint32_t tx = (la ? a_bytes : 0) + (lbb ? b_bytes : 0);
      if (elect()) {
      if (rank == 0) mbar_expect_tx(&full[s], tx * CG);
      uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
      const uint32_t bar = CG == 2 ? (smem_u32(&full[s]) & kPeerMask) : smem_u32(&full[s]);
      if (la) {
        if (CG == 2)
          asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(sa)), "l"(reinterpret_cast<uint64_t>(&p.tm_a)), "r"(bar), "r"(0), "r"(0) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(sa)), "l"(reinterpret_cast<uint64_t>(&p.tm_a)), "r"(bar), "r"(0), "r"(0) : "memory");
      }
      if (lbb) {
        if (CG == 2)
          asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(sa + a_bytes)), "l"(reinterpret_cast<uint64_t>(&p.tm_b)), "r"(bar), "r"(0), "r"(0) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(sa + a_bytes)), "l"(reinterpret_cast<uint64_t>(&p.tm_b)), "r"(bar), "r"(0), "r"(0) : "memory");
      }
      }
      __syncwarp();
    }
  } else if (warp == 1 && rank == 0) {
    // ---------------- MMA issuer (leader CTA) ----------------
    const uint32_t idesc = idesc_f16(static_cast<uint32_t>(p.M) * CG, static_cast<uint32_t>(p.N));
    const long long t0 = clock64();
    int n_blocks = 0;
    const int nacc_mask = p.nacc - 1;   // nacc in {1, 2, 4}
    const bool commit_each = !loads && p.commit_every == 1;
    const bool commit_blocks = !loads && p.commit_every >= 4;
    const int commit_block_mask = p.commit_every >= 4 ? p.commit_every / 4 - 1 : 0;
    for (int kb = 0; kb < p.kblocks; ++kb) {
      const int s = kb % p.stages;
      if (loads) {
        mbar_wait(&full[s], (kb / p.stages) & 1);
        if (!(p.flags & 1)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t sa = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
      const uint64_t da = desc_sw128(sa), db = desc_sw128(sa + a_bytes);
      for (int j = 0; j < p.kblock / 4; ++j, ++n_blocks) {
        if (elect()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // accumulator of MMA number n = 4 * n_blocks + k: n % nacc (nacc in {1, 2, 4}) = k & (nacc - 1), a per-k constant
          const uint32_t td = tmem + static_cast<uint32_t>((k & nacc_mask) * p.N);
          const uint32_t acc = (n_blocks > 0 || k > nacc_mask) ? 1u : 0u;
          if (CG == 1) {
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}\n"
                         ::"r"(td), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, q;\n\t}\n"
                         ::"r"(td), "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
          }
          if (commit_each) {
            if (CG == 1)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy)) : "memory");
            else
              asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                           ::"r"(smem_u32(&dummy)), "h"(static_cast<uint16_t>(3)) : "memory");
          }
        }
        if (commit_blocks && ((n_blocks + 1) & commit_block_mask) == 0) {
          if (CG == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy)) : "memory");
          else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&dummy)), "h"(static_cast<uint16_t>(3)) : "memory");
        }
        }
        __syncwarp();
      }
      if (loads && elect()) {
        if (CG == 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(smem_u32(&empty[s])), "h"(static_cast<uint16_t>(3)) : "memory");
      }
      __syncwarp();
    }
    if (elect()) {
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&done)), "h"(static_cast<uint16_t>(3)) : "memory");
    }
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (lane == 0) p.cycles[blockIdx.x] = t1 - t0;
  }
  if (!(warp == 1 && rank == 0) && threadIdx.x == 64) {
    mbar_wait(&done, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 1) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeFn enc, CUtensorMap* m, void* base, int rows, int box_rows) {
  cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(rows)};
  cuuint64_t str[1] = {128};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  printf("# %s, %d SMs, max SM clock %d MHz\n", prop.name, sms, clk_khz / 1000);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
  __half* src;
  CK(cudaMalloc(&src, 256 * 128));
  std::vector<__half> h(256 * 64);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half(0.01f * static_cast<float>(i % 61) - 0.3f);
  CK(cudaMemcpy(src, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  long long* cyc;
  CK(cudaMalloc(&cyc, sizeof(long long) * 512));
  const int smem_bytes = 4 * (128 * 128 + 256 * 128) + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int kblocks = 2048;
  printf("# cg  M   N  loadA(every) loadB | cycles/MMA (median CTA, min..max) | floor M*N/256/cg' | smem read B/cyc/SM | tma write B/cyc/SM | TFLOP/s chip @event time\n");
  struct Cfg { int cg, N, la, lb, nacc, ce, kblock, stages; int M = 128; int flags = 0; };
  std::vector<Cfg> cfgs;
  const int set = argc > 1 ? atoi(argv[1]) : 0;
  if (set == 1) {
    // (5) M = 64 per CTA: does the instruction time follow M (rows streamed) rather than N?
    for (int cg = 1; cg <= 2; ++cg)
      for (int N : {64, 128, 256}) {
        Cfg c{cg, N, 0, 0, 1, 0, 4, 4};
        c.M = 64;
        cfgs.push_back(c);
        c.M = 128;
        cfgs.push_back(c);
      }
    // (6) where does the per-stage bubble come from?  handshake only (no TMA) / no fence / deeper k-blocks
    for (int N : {128, 256})
      for (int kblock : {4, 8, 16})
        for (int flags : {0, 1, 2, 3}) {
          Cfg c{1, N, 1, 1, 1, 0, kblock, 4};
          c.flags = flags;
          cfgs.push_back(c);
        }
    for (int kblock : {4, 8, 16})
      for (int flags : {0, 2}) {
        Cfg c{2, 256, 1, 1, 1, 0, kblock, 4};
        c.flags = flags;
        cfgs.push_back(c);
      }
  }
  if (set == 0) {
    // (1) is the ~131-cycle figure a dependency latency?  round-robin over independent accumulators, no loads, no commits
    for (int cg = 1; cg <= 2; ++cg)
      for (int N : {64, 128, 256})
        for (int nacc : {1, 2, 4}) {
          if (nacc * N > 512) continue;
          cfgs.push_back(Cfg{cg, N, 0, 0, nacc, 0, 4, 4});
        }
    // (2) what does a tcgen05.commit cost?  no loads, commit every n MMAs to a barrier nobody waits on
    for (int N : {128, 256})
      for (int ce : {1, 4, 8, 16})
        cfgs.push_back(Cfg{1, N, 0, 0, 1, ce, 4, 4});
    for (int ce : {4, 16}) cfgs.push_back(Cfg{1, 128, 0, 0, 2, ce, 4, 4});
    // (3) the full producer/consumer ring: stages x MMAs per stage, operands re-loaded every stage
    for (int cg = 1; cg <= 2; ++cg)
      for (int N : {128, 256})
        for (int nacc : {1, 2}) {
          if (nacc * N > 512) continue;
          for (int kblock : {4, 8})
            for (int stages : {2, 4})
              cfgs.push_back(Cfg{cg, N, 1, 1, nacc, 0, kblock, stages});
        }
    // (4) conv-like: A every 9th stage only (row-strip reuse), B every stage / never (resident weights)
    for (int cg = 1; cg <= 2; ++cg)
      for (int N : {64, 128})
        for (int lb : {0, 1})
          cfgs.push_back(Cfg{cg, N, 9, lb, 2, 0, 8, 4});
  }
  for (const Cfg& c : cfgs) {
    Params p;
    memset(&p, 0, sizeof(p));
    make_map(enc, &p.tm_a, src, 128, 128);
    make_map(enc, &p.tm_b, src, 256, c.N / c.cg);
    p.N = c.N; p.stages = c.stages; p.kblocks = kblocks * 4 / c.kblock; p.load_a = c.la; p.load_b = c.lb; p.cycles = cyc;
    p.nacc = c.nacc; p.commit_every = c.ce; p.kblock = c.kblock; p.M = c.M; p.flags = c.flags;
    CK(cudaMemset(cyc, 0, sizeof(long long) * 512));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int grid = c.cg == 2 ? (sms / 2) * 2 : sms;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c.cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best_ms = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      if (c.cg == 1) CK(cudaLaunchKernelEx(&cfg, rate_kernel<1>, p)); else CK(cudaLaunchKernelEx(&cfg, rate_kernel<2>, p));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      best_ms = std::min(best_ms, ms);
    }
    std::vector<long long> hc(grid);
    CK(cudaMemcpy(hc.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    std::vector<double> per;
    for (int i = 0; i < grid; i += c.cg) per.push_back(static_cast<double>(hc[i]) / (kblocks * 4.0));
    std::sort(per.begin(), per.end());
    const double med = per[per.size() / 2];
    const double read_b = (128 * 32 + (c.N / c.cg) * 32) / med;
    const double wr_b = ((c.la ? 128.0 * 128 / c.la : 0.0) + (c.lb ? (c.N / c.cg) * 128.0 : 0.0)) / c.kblock / med;
    const double flop = 2.0 * c.M * c.cg * c.N * 16 * 4.0 * kblocks * (grid / c.cg);
    printf("  %d  %3d %3d   %d        %d  nacc %d commit/%-2d kblk %2d st %d fl %d | %7.1f (%6.1f..%6.1f) | %5.0f | %6.1f | %6.1f | %8.1f\n", c.cg, c.M * c.cg, c.N, c.la, c.lb,
           c.nacc, c.ce, c.kblock, c.stages, c.flags, med, per.front(), per.back(),
           128.0 * c.N / 256.0, read_b, wr_b, flop / (best_ms * 1e-3) / 1e12);
    fflush(stdout);
  }
  return 0;
}
