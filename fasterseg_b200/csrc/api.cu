// api.cu -- the extern "C" surface of libfsb200.so (declared in include/fsb200.h).
#include <stdlib.h>

#include <mutex>

#include "fsb_internal.h"

namespace fsb {

static thread_local char g_err[512] = "";
thread_local cudaError_t g_launch_err = cudaSuccess;

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  return FSB_ERR_CUDA;
}

static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("FSB_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// defined in the other translation units
int pack_conv_weight(const fsb_conv_desc*, const float*, int64_t, int64_t, void*, cudaStream_t);
int stem_conv_nchw_launch(int, int, int, int, const void*, int, const float*, const float*, const float*, void*, int, uint32_t,
                          cudaStream_t);
int bilinear_launch(int, int, int, int, int, int, const void*, int, void*, int, uint32_t, cudaStream_t);
int upsample_logits_launch(int, int, int, int, int, int, const void*, int, void*, int, cudaStream_t);
int upsample_argmax_launch(int, int, int, int, int, int, const void*, int, uint8_t*, cudaStream_t);
int nchw_to_nhwc_launch(int, int, int, int, const void*, int, void*, int, cudaStream_t);
int nhwc_to_nchw_launch(int, int, int, int, const void*, int, void*, int, cudaStream_t);
int copy_channels_launch(int64_t, int, const void*, int, void*, int, cudaStream_t);
int bn_fold_launch(int, const float*, const float*, const float*, const float*, float, const float*, float*, float*, cudaStream_t);
int bn_stats_launch(int64_t, int, const void*, int, float*, cudaStream_t);
int bn_finalize_launch(int, const float*, double, const float*, const float*, float, float, float*, float*, float*, float*,
                       float*, float*, cudaStream_t);
int affine_act_launch(int64_t, int, const void*, int, const float*, const float*, void*, int, uint32_t, cudaStream_t);

static int check_desc(const fsb_conv_desc* d) {
  if (!d) return set_error(FSB_ERR_INVALID, "null conv desc");
  if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0)
    return set_error(FSB_ERR_INVALID, "conv desc: non-positive dimension");
  if (!(d->ksize == 1 || d->ksize == 3)) return set_error(FSB_ERR_INVALID, "conv desc: ksize must be 1 or 3");
  if (!(d->stride == 1 || d->stride == 2)) return set_error(FSB_ERR_INVALID, "conv desc: stride must be 1 or 2");
  if (d->dil < 1) return set_error(FSB_ERR_INVALID, "conv desc: dil must be >= 1");
  if (d->x_cstride < d->Cin || d->y_cstride < d->Cout) return set_error(FSB_ERR_INVALID, "conv desc: channel stride < channels");
  const int He = d->H - d->off_h, We = d->W - d->off_w;
  const int ext = d->dil * (d->ksize - 1) + 1;
  const int Ho = (He + 2 * d->pad - ext) / d->stride + 1, Wo = (We + 2 * d->pad - ext) / d->stride + 1;
  if (Ho != d->Ho || Wo != d->Wo) {
    char buf[160];
    snprintf(buf, sizeof(buf), "conv desc: Ho/Wo (%d,%d) inconsistent with geometry (expected %d,%d)", d->Ho, d->Wo, Ho, Wo);
    return set_error(FSB_ERR_INVALID, buf);
  }
  return FSB_OK;
}

}  // namespace fsb

using namespace fsb;

extern "C" {

int fsb_abi_version(void) { return FSB_ABI_VERSION; }
const char* fsb_last_error_string(void) { return g_err; }

int fsb_set_pdl(int enabled) {
  g_pdl = enabled ? 1 : 0;
  return FSB_OK;
}

int fsb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(FSB_ERR_NO_DEVICE, "no CUDA device");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDeviceProperties");
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return FSB_OK;
}

size_t fsb_conv_packed_bytes(const fsb_conv_desc* d) {
  if (!d || d->Cin <= 0 || d->Cout <= 0) return 0;
  const ConvGeom g = conv_geom(d);
  return static_cast<size_t>(g.taps) * g.npad * g.kpad * 2;
}

int fsb_pack_conv_weight(const fsb_conv_desc* d, const float* w, int64_t so, int64_t si, void* packed, void* stream) {
  if (!d || !w || !packed) return set_error(FSB_ERR_INVALID, "pack_conv_weight: null argument");
  return pack_conv_weight(d, w, so, si, packed, static_cast<cudaStream_t>(stream));
}

int fsb_bn_fold(int C, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                const float* conv_bias, float* scale, float* shift, void* stream) {
  if (C <= 0 || !mean || !var || !scale || !shift) return set_error(FSB_ERR_INVALID, "bn_fold: bad argument");
  return bn_fold_launch(C, gamma, beta, mean, var, eps, conv_bias, scale, shift, static_cast<cudaStream_t>(stream));
}

int fsb_conv_fwd(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift, void* y,
                 float* stats, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (!x || !wpacked || !y) return set_error(FSB_ERR_INVALID, "conv_fwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((d->flags & FSB_CONV_FORCE_DIRECT) || !conv_tc_supported(d)) return conv_direct_launch(d, x, wpacked, scale, shift, y, stats, st);
  return conv_tc_launch(d, x, wpacked, scale, shift, y, stats, st);
}

int fsb_stem_conv_nchw(int N, int H, int W, int Cout, const void* x, int x_is_f32, const float* w, const float* scale,
                       const float* shift, void* y, int y_cstride, uint32_t flags, void* stream) {
  if (N <= 0 || H <= 0 || W <= 0 || Cout <= 0 || !x || !w || !y || y_cstride < Cout)
    return set_error(FSB_ERR_INVALID, "stem_conv_nchw: bad argument");
  return stem_conv_nchw_launch(N, H, W, Cout, x, x_is_f32, w, scale, shift, y, y_cstride, flags, static_cast<cudaStream_t>(stream));
}

int fsb_bilinear_fwd(int N, int C, int Hi, int Wi, int Ho, int Wo, const void* x, int xcs, void* y, int ycs, uint32_t flags,
                     void* stream) {
  if (N <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || !x || !y) return set_error(FSB_ERR_INVALID, "bilinear: bad argument");
  return bilinear_launch(N, C, Hi, Wi, Ho, Wo, x, xcs, y, ycs, flags, static_cast<cudaStream_t>(stream));
}

int fsb_upsample_logits_nchw(int N, int C, int Hi, int Wi, int Ho, int Wo, const void* x, int xcs, void* y, int out_dtype,
                             void* stream) {
  if (N <= 0 || C <= 0 || !x || !y || xcs < C) return set_error(FSB_ERR_INVALID, "upsample_logits: bad argument");
  return upsample_logits_launch(N, C, Hi, Wi, Ho, Wo, x, xcs, y, out_dtype, static_cast<cudaStream_t>(stream));
}

int fsb_upsample_argmax(int N, int C, int Hi, int Wi, int Ho, int Wo, const void* x, int xcs, uint8_t* labels, void* stream) {
  if (N <= 0 || C <= 0 || C > 255 || !x || !labels || xcs < C) return set_error(FSB_ERR_INVALID, "upsample_argmax: bad argument");
  return upsample_argmax_launch(N, C, Hi, Wi, Ho, Wo, x, xcs, labels, static_cast<cudaStream_t>(stream));
}

int fsb_nchw_to_nhwc_f16(int N, int C, int H, int W, const void* x, int x_is_f32, void* y, int ycs, void* stream) {
  if (N <= 0 || C <= 0 || !x || !y || ycs < C) return set_error(FSB_ERR_INVALID, "nchw_to_nhwc: bad argument");
  return nchw_to_nhwc_launch(N, C, H, W, x, x_is_f32, y, ycs, static_cast<cudaStream_t>(stream));
}
int fsb_nhwc_f16_to_nchw(int N, int C, int H, int W, const void* x, int xcs, void* y, int y_is_f32, void* stream) {
  if (N <= 0 || C <= 0 || !x || !y || xcs < C) return set_error(FSB_ERR_INVALID, "nhwc_to_nchw: bad argument");
  return nhwc_to_nchw_launch(N, C, H, W, x, xcs, y, y_is_f32, static_cast<cudaStream_t>(stream));
}
int fsb_copy_channels(int64_t pixels, int C, const void* x, int xcs, void* y, int ycs, void* stream) {
  if (pixels <= 0 || C <= 0 || !x || !y) return set_error(FSB_ERR_INVALID, "copy_channels: bad argument");
  return copy_channels_launch(pixels, C, x, xcs, y, ycs, static_cast<cudaStream_t>(stream));
}

int fsb_bn_stats(int64_t pixels, int C, const void* x, int xcs, float* stats, void* stream) {
  if (pixels <= 0 || C <= 0 || !x || !stats) return set_error(FSB_ERR_INVALID, "bn_stats: bad argument");
  return bn_stats_launch(pixels, C, x, xcs, stats, static_cast<cudaStream_t>(stream));
}
int fsb_bn_finalize(int C, const float* stats, double count, const float* gamma, const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                    void* stream) {
  if (C <= 0 || !stats || count <= 0) return set_error(FSB_ERR_INVALID, "bn_finalize: bad argument");
  return bn_finalize_launch(C, stats, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, save_mean,
                            save_invstd, static_cast<cudaStream_t>(stream));
}
int fsb_affine_act(int64_t pixels, int C, const void* x, int xcs, const float* scale, const float* shift, void* y, int ycs,
                   uint32_t flags, void* stream) {
  if (pixels <= 0 || C <= 0 || !x || !y || !scale || !shift) return set_error(FSB_ERR_INVALID, "affine_act: bad argument");
  return affine_act_launch(pixels, C, x, xcs, scale, shift, y, ycs, flags, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
