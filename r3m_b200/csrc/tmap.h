// Host-side TMA descriptor (CUtensorMap) encoders.  The driver entry points are resolved at run time through
// cudaGetDriverEntryPoint so that the shared library has no link-time dependency on libcuda (it must load on a
// GPU-less build host for the symbol check).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace r3m {

// NHWC bf16 (elem_bytes 2) or fp32 (4) activation tensor viewed as rank-4 (C, W, H, N) for im2col-mode loads.
//   lower_* / upper_* : bounding-box corners of the base pixel (CUDA driver API convention)
//   channels          : channels per load (<= 64 -> one 128-byte swizzle row)
//   pixels            : base pixels per load (GEMM rows of one stage)
//   trav_stride       : traversal stride of the base pixel
// Returns an empty string on success, else a diagnostic.
std::string encode_im2col_map(CUtensorMap* out, const void* base, int C, int W, int H, int N, int lower_w, int lower_h,
                              int upper_w, int upper_h, int channels, int pixels, int trav_stride, int elem_bytes = 2);

// Row-major 2-D bf16 matrix [outer][inner] (inner contiguous), swizzled boxes of box_inner x box_outer.
// swizzle_bytes: 128 (box_inner * 2 <= 128) or 64 (box_inner * 2 <= 64).
std::string encode_tiled_2d_map(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                                uint64_t row_stride_bytes, int box_inner, int box_outer, int swizzle_bytes = 128,
                                int elem_bytes = 2);

// NHWC bf16 tensor viewed as rank-4 (C, W, H, N) for TILED loads / stores of a (box_c, box_w, box_h, 1) box, 128-byte
// swizzle (box_c * 2 == 128); coordinates may lie outside the tensor (zero fill on load, dropped on store).
std::string encode_tiled_4d_map(CUtensorMap* out, const void* base, int C, int W, int H, int N, int box_c, int box_w,
                                int box_h);

}  // namespace r3m
