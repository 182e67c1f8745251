// fsb_internal.h -- declarations shared between the translation units of libfsb200.so (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fsb200.h"

namespace fsb {

int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// derived geometry shared by the packer and the kernels
struct ConvGeom {
  int taps;     // ksize^2
  int bk;       // K chunk (channels per TMA box / smem row): 64 if Cin % 64 == 0 else 32
  int kpad;     // Cin rounded up to bk
  int n_tiles;  // output-channel tiles (grid.y)
  int n_tile;   // UMMA N per tile (multiple of 16, <= 256)
  int npad;     // n_tile * n_tiles
};
ConvGeom conv_geom(const fsb_conv_desc* d);

int conv_tc_supported(const fsb_conv_desc* d);
int conv_tc_launch(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift,
                   void* y, float* stats, cudaStream_t stream);
int conv_direct_launch(const fsb_conv_desc* d, const void* x, const void* wpacked, const float* scale, const float* shift,
                       void* y, float* stats, cudaStream_t stream);

}  // namespace fsb
