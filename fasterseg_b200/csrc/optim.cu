// optim.cu -- the step's tail on the flat gradient buffer: clip_grad_norm_ + SGD(momentum, weight decay) over ~5 000 parameter tensors.
//
// The captured supernet passes leave every weight gradient in ONE flat fp32 buffer (graphed.FlatGrads); the drivers then call
// nn.utils.clip_grad_norm_(model.parameters(), 5) and optimizer.step() (search/train_search.py:249-250), whose torch
// implementations walk all ~5 000 Parameter objects in Python every step: 17 + 42 ms of host time per 165 ms step with the GPU
// idle (profiles/r2_step_census_pretrain.log).  Here the same arithmetic runs as three table-driven kernels over the flat buffer:
//   sqnorm : per-block sum of squares of the LIVE segments (parameters that received a gradient this step), fixed-order final sum
//            -> total norm and the clip coefficient min(1, max_norm / (norm + 1e-6)) on the device (no host sync)
//   scale  : g *= coef                                   (what clip_grad_norm_ leaves in param.grad)
//   sgd    : d = g + wd * p;  m = momentum * m + d;  p -= lr * m      (torch.optim.SGD, dampening 0, nesterov off; a zero-initialised
//            momentum buffer is arithmetically the same as torch's "first step: buf = d")
// A segment = one parameter tensor (device pointer of its storage, offset into the flat gradient / momentum buffers, numel); a block
// map assigns 4096-element chunks to CTAs once; parameters without a gradient this step are skipped exactly like torch skips them.
#include "fsb_common.cuh"
#include "fsb_internal.h"

namespace fsb {

struct FlatSeg {
  float* p;        // parameter storage (fp32, contiguous)
  uint32_t off;    // float offset into the flat gradient and momentum buffers (multiple of 4)
  uint32_t n;      // elements
};
constexpr int kFlatChunk = 4096;   // elements per CTA (256 threads x 4 x float4)

__device__ __forceinline__ bool flat_range(const int2* __restrict__ map, const FlatSeg* __restrict__ segs, const uint8_t* __restrict__ live,
                                           FlatSeg& s, uint32_t& lo, uint32_t& hi) {
  const int2 m = map[blockIdx.x];
  if (!live[m.x]) return false;
  s = segs[m.x];
  lo = static_cast<uint32_t>(m.y) * kFlatChunk;
  hi = min(lo + kFlatChunk, s.n);
  return true;
}

__global__ void __launch_bounds__(256)
flat_sqnorm_kernel(const int2* __restrict__ map, const FlatSeg* __restrict__ segs, const uint8_t* __restrict__ live,
                   const float* __restrict__ G, float* __restrict__ partial) {
  FlatSeg s;
  uint32_t lo, hi;
  float acc = 0.f;
  if (flat_range(map, segs, live, s, lo, hi)) {
    const float* g = G + s.off;
    for (uint32_t i = lo + threadIdx.x * 4; i < hi; i += 1024) {
      if (i + 4 <= hi) {
        const float4 v = *reinterpret_cast<const float4*>(g + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      } else {
        for (uint32_t j = i; j < hi; ++j) acc += g[j] * g[j];
      }
    }
  }
  __shared__ float sw[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += sw[w];
    partial[blockIdx.x] = a;
  }
}

// out[0] = total norm, out[1] = clip coefficient; partial sums added in a fixed order (double)
__global__ void __launch_bounds__(1024)
flat_norm_final_kernel(const float* __restrict__ partial, int n, const float* __restrict__ extra_sq, float max_norm, float* __restrict__ out) {
  __shared__ double sd[1024];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) a += static_cast<double>(partial[i]);
  sd[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sd[threadIdx.x] += sd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double sq = sd[0] + (extra_sq ? static_cast<double>(*extra_sq) : 0.0);
    const float norm = static_cast<float>(sqrt(sq));
    out[0] = norm;
    out[1] = fminf(1.0f, max_norm / (norm + 1e-6f));
  }
}

__global__ void __launch_bounds__(256)
flat_scale_kernel(const int2* __restrict__ map, const FlatSeg* __restrict__ segs, const uint8_t* __restrict__ live, float* __restrict__ G,
                  const float* __restrict__ coef) {
  FlatSeg s;
  uint32_t lo, hi;
  if (!flat_range(map, segs, live, s, lo, hi)) return;
  const float c = *coef;
  float* g = G + s.off;
  for (uint32_t i = lo + threadIdx.x * 4; i < hi; i += 1024) {
    if (i + 4 <= hi) {
      float4 v = *reinterpret_cast<float4*>(g + i);
      v.x *= c; v.y *= c; v.z *= c; v.w *= c;
      *reinterpret_cast<float4*>(g + i) = v;
    } else {
      for (uint32_t j = i; j < hi; ++j) g[j] *= c;
    }
  }
}

__global__ void __launch_bounds__(256)
flat_sgd_kernel(const int2* __restrict__ map, const FlatSeg* __restrict__ segs, const uint8_t* __restrict__ live, const float* __restrict__ G,
                float* __restrict__ M, float lr, float momentum, float wd) {
  FlatSeg s;
  uint32_t lo, hi;
  if (!flat_range(map, segs, live, s, lo, hi)) return;
  const float* g = G + s.off;
  float* m = M + s.off;
  float* p = s.p;
  const bool vec = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
  for (uint32_t i = lo + threadIdx.x * 4; i < hi; i += 1024) {
    if (vec && i + 4 <= hi) {
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 pv = *reinterpret_cast<float4*>(p + i);
      mv.x = momentum * mv.x + (gv.x + wd * pv.x);
      mv.y = momentum * mv.y + (gv.y + wd * pv.y);
      mv.z = momentum * mv.z + (gv.z + wd * pv.z);
      mv.w = momentum * mv.w + (gv.w + wd * pv.w);
      pv.x -= lr * mv.x; pv.y -= lr * mv.y; pv.z -= lr * mv.z; pv.w -= lr * mv.w;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(p + i) = pv;
    } else {
      const uint32_t e = min(i + 4, hi);
      for (uint32_t j = i; j < e; ++j) {
        const float mj = momentum * m[j] + (g[j] + wd * p[j]);
        m[j] = mj;
        p[j] -= lr * mj;
      }
    }
  }
}

int flat_check(const void* map, int nblocks, const void* segs, const void* live) {
  if (!map || !segs || !live || nblocks <= 0) return set_error(FSB_ERR_INVALID, "flat optimizer: bad table");
  return FSB_OK;
}

}  // namespace fsb

using namespace fsb;

extern "C" {

int fsb_flat_chunk(void) { return kFlatChunk; }

int fsb_flat_grad_norm(const void* map, int nblocks, const void* segs, const uint8_t* live, const float* G, float* partial,
                       const float* extra_sq, float max_norm, float* out2, void* stream) {
  if (int rc = flat_check(map, nblocks, segs, live)) return rc;
  if (!G || !partial || !out2) return set_error(FSB_ERR_INVALID, "fsb_flat_grad_norm: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  flat_sqnorm_kernel<<<nblocks, 256, 0, st>>>(static_cast<const int2*>(map), static_cast<const FlatSeg*>(segs), live, G, partial);
  flat_norm_final_kernel<<<1, 1024, 0, st>>>(partial, nblocks, extra_sq, max_norm, out2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "fsb_flat_grad_norm launch");
  return FSB_OK;
}

int fsb_flat_scale(const void* map, int nblocks, const void* segs, const uint8_t* live, float* G, const float* coef, void* stream) {
  if (int rc = flat_check(map, nblocks, segs, live)) return rc;
  if (!G || !coef) return set_error(FSB_ERR_INVALID, "fsb_flat_scale: null pointer");
  flat_scale_kernel<<<nblocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const int2*>(map), static_cast<const FlatSeg*>(segs), live, G,
                                                                          coef);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "fsb_flat_scale launch");
  return FSB_OK;
}

int fsb_flat_sgd(const void* map, int nblocks, const void* segs, const uint8_t* live, const float* G, float* M, float lr, float momentum,
                 float weight_decay, void* stream) {
  if (int rc = flat_check(map, nblocks, segs, live)) return rc;
  if (!G || !M) return set_error(FSB_ERR_INVALID, "fsb_flat_sgd: null pointer");
  flat_sgd_kernel<<<nblocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const int2*>(map), static_cast<const FlatSeg*>(segs), live, G, M,
                                                                        lr, momentum, weight_decay);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "fsb_flat_sgd launch");
  return FSB_OK;
}

}  // extern "C"
