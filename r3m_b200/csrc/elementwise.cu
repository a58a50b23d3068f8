// HBM-bound kernels of the R3M pretraining step: input normalisation + stem re-layout, BatchNorm apply (train / eval)
// with fused residual + ReLU, stem BN+ReLU+maxpool, global average pool, BatchNorm backward, pooling backward,
// fused Adam and the filter re-packers.  All activations are NHWC bf16, 16-byte vector accesses (8 channels per
// thread), fp32 arithmetic.  Reference semantics: torchvision resnet.py (BasicBlock.forward :89-105,
// Bottleneck.forward :143-163, _forward_impl :266-282) and torch.nn.BatchNorm2d / MaxPool2d / Adam defaults.
#include "elementwise.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "launch.h"
#include "ptx.cuh"

#ifndef R3M_STREAM_HINTS
#define R3M_STREAM_HINTS 0  // bit 0: evict-first loads, bit 1: streaming stores in the HBM-bound kernels (A/B knob)
#endif
#define R3M_PRAGMA_(x) _Pragma(#x)
#define R3M_UNROLL(n) R3M_PRAGMA_(unroll n)
#ifndef R3M_BN_APPLY_UNROLL
#define R3M_BN_APPLY_UNROLL 2
#endif
#ifndef R3M_BN_BWD_UNROLL
#define R3M_BN_BWD_UNROLL 2
#endif

namespace r3m {

namespace {

typedef __nv_bfloat16 bf16;

struct F8 {
  float v[8];
};
__device__ __forceinline__ uint4 ld8_raw(const bf16* p) {
#if R3M_STREAM_HINTS & 1
  return __ldcs(reinterpret_cast<const uint4*>(p));  // streaming (evict-first) read
#else
  return __ldg(reinterpret_cast<const uint4*>(p));  // every ld8 source is read-only within its kernel
#endif
}
__device__ __forceinline__ F8 cvt8(const uint4& u) {
  F8 f;
  f.v[0] = bf16lo(u.x);
  f.v[1] = bf16hi(u.x);
  f.v[2] = bf16lo(u.y);
  f.v[3] = bf16hi(u.y);
  f.v[4] = bf16lo(u.z);
  f.v[5] = bf16hi(u.z);
  f.v[6] = bf16lo(u.w);
  f.v[7] = bf16hi(u.w);
  return f;
}
__device__ __forceinline__ F8 ld8(const bf16* p) {
  const uint4 u = ld8_raw(p);
  F8 f;
  f.v[0] = bf16lo(u.x);
  f.v[1] = bf16hi(u.x);
  f.v[2] = bf16lo(u.y);
  f.v[3] = bf16hi(u.y);
  f.v[4] = bf16lo(u.z);
  f.v[5] = bf16hi(u.z);
  f.v[6] = bf16lo(u.w);
  f.v[7] = bf16hi(u.w);
  return f;
}
// last-use read: the line is marked evict-first so that it does not displace data the next kernels will read
// (R3M_STREAM_HINTS bit 2; used by the BatchNorm apply passes, whose inputs are not touched again before the backward /
// at all)
__device__ __forceinline__ F8 ld8_last(const bf16* p) {
#if R3M_STREAM_HINTS & 4
  const uint4 u = __ldcs(reinterpret_cast<const uint4*>(p));
  F8 f;
  f.v[0] = bf16lo(u.x);
  f.v[1] = bf16hi(u.x);
  f.v[2] = bf16lo(u.y);
  f.v[3] = bf16hi(u.y);
  f.v[4] = bf16lo(u.z);
  f.v[5] = bf16hi(u.z);
  f.v[6] = bf16lo(u.w);
  f.v[7] = bf16hi(u.w);
  return f;
#else
  return ld8(p);
#endif
}
__device__ __forceinline__ void st8(bf16* p, const F8& f) {
  uint4 u;
  u.x = pack_bf16x2(f.v[0], f.v[1]);
  u.y = pack_bf16x2(f.v[2], f.v[3]);
  u.z = pack_bf16x2(f.v[4], f.v[5]);
  u.w = pack_bf16x2(f.v[6], f.v[7]);
#if R3M_STREAM_HINTS & 2
  __stcs(reinterpret_cast<uint4*>(p), u);
#else
  *reinterpret_cast<uint4*>(p) = u;
#endif
}

// Per-channel statistics arrive either as fp32 arrays or "raw": the fixed-point accumulators the producing kernel added
// into (fx_add, ptx.cuh) — the engine's path: no finalize pass in the producer, the consumer converts on read.
//   conv statistics  raw layout: channel c -> kFxWords words of the sum, then kFxWords of the sum of squares
//                    (sum == sq == base)
//   backward sums    raw layout: entry i   -> kFxWords words
__device__ __forceinline__ float bsum_at(const float* p, int raw, int i) {
  if (!raw) return p[i];
  return fx_to_float(reinterpret_cast<const unsigned long long*>(p) + i * kFxWords);
}

// Batch mean and biased variance of channel c from the per-channel sum / sum of squares of M elements.  The textbook
// E[x^2] - mean^2 loses eps * mean^2 / var of the variance to cancellation when |mean| >> std, so the subtraction is
// carried out in fp64: on the engine's path ("raw") the sums are exact (fixed-point accumulators, ptx.cuh) and are
// converted straight to fp64, so the variance is good to ~1e-16 * mean^2 / var; with fp32 sums (kernel-level C ABI) the
// inputs are already rounded and fp64 only keeps the subtraction itself from adding to that.  bn_apply / stem_pool
// (bn_coeffs) and bn_publish (the statistics the backward pass reads) share this function: identical bits.
__device__ __forceinline__ void bn_batch_moments(const float* sum, const float* sq, int raw, int c, int M, float& mean,
                                                 float& var) {
  if (raw == 2) {  // A/B aid: fp32 conversion of the sums and fp32 subtraction, as in round 1
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(sum);
    const float inv_mf = 1.0f / (float)M;
    mean = fx_to_float(p + (2 * c + 0) * kFxWords) * inv_mf;
    var = fmaxf(fx_to_float(p + (2 * c + 1) * kFxWords) * inv_mf - mean * mean, 0.f);
    return;
  }
  if (raw == 3) {  // the producing conv's last CTAs have already published (mean, variance) by the same law (fx_moments)
    mean = sum[c];
    var = sq[c];
    return;
  }
  if (raw) {
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(sum);
    fx_moments(p + (2 * c + 0) * kFxWords, p + (2 * c + 1) * kFxWords, M, mean, var);
    return;
  }
  const double inv_m = 1.0 / (double)M;
  const double m = (double)sum[c] * inv_m;
  mean = (float)m;
  var = fmaxf((float)((double)sq[c] * inv_m - m * m), 0.f);
}

__device__ __forceinline__ void bn_coeffs(int train, const float* sum, const float* sq, int raw, int M,
                                          const float* gamma, const float* beta, const float* rm, const float* rv, int c,
                                          float& scale, float& shift, float& mean, float& var) {
  if (train) {
    bn_batch_moments(sum, sq, raw, c, M, mean, var);
  } else {
    mean = rm[c];
    var = rv[c];
  }
  const float rstd = 1.0f / sqrtf(var + kBnEps);
  scale = gamma[c] * rstd;
  shift = beta[c] - mean * scale;
}

// block 0 publishes the batch statistics for backward and folds them into the running estimates
__device__ __forceinline__ void bn_publish(int C, int M, const float* sum, const float* sq, int raw, float* save_mean,
                                           float* save_rstd, float* rm, float* rv, int update_running) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, var;
    bn_batch_moments(sum, sq, raw, c, M, mean, var);
    if (save_mean) save_mean[c] = mean;
    if (save_rstd) save_rstd[c] = 1.0f / sqrtf(var + kBnEps);
    if (update_running && rm && rv) {
      const float unbiased = (M > 1) ? var * ((float)M / (float)(M - 1)) : var;
      rm[c] = (1.f - kBnMomentum) * rm[c] + kBnMomentum * mean;
      rv[c] = (1.f - kBnMomentum) * rv[c] + kBnMomentum * unbiased;
    }
  }
}

// ------------------------------------------------------------------------------------------------ ordered reduce
// Deterministic grid-wide sum of per-block partial vectors (V floats each).  Every block adds its partial into V
// fixed-point accumulators (fx_add, ptx.cuh: exact integer arithmetic, so the arrival order does not matter — fp32
// atomics would make the result depend on it); the last block to arrive (ticket) converts the totals to fp32, hands them
// to emit(i, value) and clears the accumulators for the next launch.  One short tail instead of a multi-level tree.
// d.scratch: V * kFxWords 64-bit words, zero on entry and on exit; d.tickets: one int, likewise.  Blocks along x reduce
// together (a caller with several blockIdx.y slices passes each slice its own DetScratch).
template <class Emit>
__device__ __forceinline__ void det_grid_reduce(const float* partial, int V, const DetScratch d, Emit emit) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(d.scratch);
  for (int i = tid; i < V; i += blockDim.x) fx_add(acc + kFxWords * i, partial[i]);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&d.tickets[0], 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = tid; i < V; i += blockDim.x) {
    emit(i, fx_to_float(acc + kFxWords * i));
    fx_clear(acc + kFxWords * i);
  }
  if (tid == 0) d.tickets[0] = 0;
}

// ------------------------------------------------------------------------------------------------ preprocess
// Input formats (enum ObsFormat in elementwise.cuh): fp32 NCHW (the reference loader's contract), uint8 NCHW
// (torchvision.io.read_image frames: 4x fewer PCIe / HBM bytes) and uint8 NHWC (decoder output order).  A thread
// produces one 16-channel group of the stem operand from 2 rows x 2 columns x 3 colours of the frame.
template <int FMT>
__device__ __forceinline__ void load_px2(const void* __restrict__ obs, int n, int c, int row, int col, float& a, float& b) {
  if (FMT == kObsF32NCHW) {
    const float2 px = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(obs) +
                                                       (((long long)n * 3 + c) * 224 + row) * 224 + col);
    a = px.x;
    b = px.y;
  } else if (FMT == kObsU8NCHW) {
    const uchar2 px = *reinterpret_cast<const uchar2*>(reinterpret_cast<const uint8_t*>(obs) +
                                                       (((long long)n * 3 + c) * 224 + row) * 224 + col);
    a = (float)px.x;
    b = (float)px.y;
  } else {  // uint8 NHWC: the two pixels are 3 bytes apart
    const uint8_t* p = reinterpret_cast<const uint8_t*>(obs) + (((long long)n * 224 + row) * 224 + col) * 3 + c;
    a = (float)p[0];
    b = (float)p[3];
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) preprocess_stem_kernel(const void* __restrict__ obs, bf16* __restrict__ xs,
                                                              int N) {
  pdl_sync();
  const long long total = (long long)N * 112 * 112 * 4;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float istd[3] = {1.f / 0.229f, 1.f / 0.224f, 1.f / 0.225f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(idx & 3);
    long long t = idx >> 2;
    const int q = (int)(t % 112);
    t /= 112;
    const int i = (int)(t % 112);
    const int n = (int)(t / 112);
    const int col = 2 * (q - 2 + kw);  // even, so a pixel pair covers dx = 0, 1
    uint32_t w[8];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      float v[2][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float p0 = 0.f, p1 = 0.f;
        const bool inside = (col >= 0 && col < 224);
        if (inside) {
          load_px2<FMT>(obs, n, c, 2 * i + dy, col, p0, p1);
          p0 = (p0 * (1.f / 255.f) - mean[c]) * istd[c];
          p1 = (p1 * (1.f / 255.f) - mean[c]) * istd[c];
        }
        v[0][c] = p0;
        v[1][c] = p1;
      }
      v[0][3] = 0.f;
      v[1][3] = 0.f;
      // element order inside the 16-wide group: (dy*2+dx)*4 + c
      w[dy * 4 + 0] = pack_bf16x2(v[0][0], v[0][1]);
      w[dy * 4 + 1] = pack_bf16x2(v[0][2], v[0][3]);
      w[dy * 4 + 2] = pack_bf16x2(v[1][0], v[1][1]);
      w[dy * 4 + 3] = pack_bf16x2(v[1][2], v[1][3]);
    }
    uint4* dst = reinterpret_cast<uint4*>(xs + idx * 16);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// ------------------------------------------------------------------------------------------------ crop + resize
// torchvision.transforms.RandomResizedCrop's arithmetic on the GPU (r3m/utils/data_loaders.py:47-50,81-102): crop box
// (top, left, h, w) of a uint8 frame -> 224 x 224 by bilinear interpolation WITH antialiasing, i.e. ATen's
// upsample_bilinear2d_aa (separable triangle filter whose support grows with the down-scale factor; weights
// normalised per output index; borders clamp to the CROP, as the reference crops first and resizes second).
// out: fp32 NCHW [N,3,224,224] in [0,255] — the reference loader's contract.
struct AaAxis {
  int lo, n;      // first source index (inside the crop) and tap count
  float center, inv, scale_w;
};
__device__ __forceinline__ AaAxis aa_axis(int o, int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  const float support = (scale >= 1.f) ? scale : 1.f;
  AaAxis a;
  a.center = scale * ((float)o + 0.5f);
  a.inv = (scale >= 1.f) ? 1.f / scale : 1.f;
  a.lo = max((int)(a.center - support + 0.5f), 0);
  a.n = min((int)(a.center + support + 0.5f), in_size) - a.lo;
  float total = 0.f;
  for (int j = 0; j < a.n; ++j) total += fmaxf(0.f, 1.f - fabsf(((float)(j + a.lo) - a.center + 0.5f) * a.inv));
  a.scale_w = total != 0.f ? 1.f / total : 0.f;
  return a;
}
__device__ __forceinline__ float aa_weight(const AaAxis& a, int j) {
  return fmaxf(0.f, 1.f - fabsf(((float)(j + a.lo) - a.center + 0.5f) * a.inv)) * a.scale_w;
}

template <bool NHWC>
__global__ void __launch_bounds__(256) crop_resize_kernel(const uint8_t* __restrict__ src, const int* __restrict__ boxes,
                                                          float* __restrict__ out, int N, int H, int W) {
  pdl_sync();
  const long long total = (long long)N * 224 * 224;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % 224);
    const int y = (int)((idx / 224) % 224);
    const int n = (int)(idx / (224 * 224));
    const int top = boxes[4 * n], left = boxes[4 * n + 1], h = boxes[4 * n + 2], w = boxes[4 * n + 3];
    const AaAxis ax = aa_axis(x, w, 224), ay = aa_axis(y, h, 224);
    float acc[3] = {0.f, 0.f, 0.f};
    for (int jy = 0; jy < ay.n; ++jy) {
      const float wy = aa_weight(ay, jy);
      const int yy = top + ay.lo + jy;
      float row[3] = {0.f, 0.f, 0.f};
      for (int jx = 0; jx < ax.n; ++jx) {
        const float wx = aa_weight(ax, jx);
        const int xx = left + ax.lo + jx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint8_t v = NHWC ? src[(((long long)n * H + yy) * W + xx) * 3 + c]
                                 : src[(((long long)n * 3 + c) * H + yy) * W + xx];
          row[c] = fmaf(wx, (float)v, row[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] = fmaf(wy, row[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(((long long)n * 3 + c) * 224 + y) * 224 + x] = acc[c];
  }
}

// ------------------------------------------------------------------------------------------------ tf32 tier
// The inference parity tier keeps activations and filters in fp32 storage, rounded to tf32 (10-bit mantissa, round to
// nearest: the tensor cores would otherwise TRUNCATE their fp32 operands, a biased error that does not average out).
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  pdl_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = round_tf32(src[i]);
}

// fp32 frames / stem operand variant of preprocess_stem_kernel: same gather, fp32 (tf32-rounded) output
template <int FMT>
__global__ void __launch_bounds__(256) preprocess_stem_f32_kernel(const void* __restrict__ obs, float* __restrict__ xs,
                                                                  int N) {
  pdl_sync();
  const long long total = (long long)N * 112 * 112 * 4;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(idx & 3);
    long long t = idx >> 2;
    const int q = (int)(t % 112);
    t /= 112;
    const int i = (int)(t % 112);
    const int n = (int)(t / 112);
    const int col = 2 * (q - 2 + kw);
    float o[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float p0 = 0.f, p1 = 0.f;
        if (col >= 0 && col < 224) {
          load_px2<FMT>(obs, n, c, 2 * i + dy, col, p0, p1);
          // the reference's arithmetic order: (x / 255 - mean) / std  (models_r3m.py:97 + transforms.Normalize)
          p0 = round_tf32((p0 / 255.f - mean[c]) / stdv[c]);
          p1 = round_tf32((p1 / 255.f - mean[c]) / stdv[c]);
        }
        o[(dy * 2 + 0) * 4 + c] = p0;
        o[(dy * 2 + 1) * 4 + c] = p1;
      }
      o[(dy * 2 + 0) * 4 + 3] = 0.f;
      o[(dy * 2 + 1) * 4 + 3] = 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(xs + idx * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  }
}

// MaxPool2d(3, 2, 1) on the (already BatchNorm-ed and ReLU-ed) fp32 stem output: y [N,H,W,C] -> a [N,H/2,W/2,C]
__global__ void __launch_bounds__(256) maxpool_f32_kernel(const float* __restrict__ y, float* __restrict__ a, int N, int H,
                                                          int W, int C) {
  pdl_sync();
  const int C4 = C >> 2, P = H / 2, Q = W / 2;
  const long long total = (long long)N * P * Q * C4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    long long t = idx / C4;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    const float ninf = __int_as_float(0xff800000);
    float4 best = make_float4(ninf, ninf, ninf, ninf);
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * p - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s2 = 0; s2 < 3; ++s2) {
        const int w = 2 * q - 1 + s2;
        if (w < 0 || w >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(y + (((long long)n * H + h) * W + w) * C) + c4);
        best.x = fmaxf(best.x, v.x);
        best.y = fmaxf(best.y, v.y);
        best.z = fmaxf(best.z, v.z);
        best.w = fmaxf(best.w, v.w);
      }
    }
    reinterpret_cast<float4*>(a)[idx] = best;
  }
}

__global__ void avgpool_fwd_f32_kernel(const float* __restrict__ a, float* __restrict__ out, int HW, int C) {
  pdl_sync();
  const int n = blockIdx.x;
  const float inv = 1.0f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    const float* base = a + (long long)n * HW * C + c;
    for (int i = 0; i < HW; ++i) s += base[(long long)i * C];
    out[(long long)n * C + c] = s * inv;
  }
}

// ------------------------------------------------------------------------------------------------ BN apply
template <bool kDual, bool kRes>
__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyArgs a) {
  pdl_sync();
  // wide layers (C >= 1024) are cut into channel slices of 256 (blockIdx.y): a block then builds the coefficient table
  // of its slice only and covers 8 rows per iteration
  const int Cs = gridDim.y > 1 ? 256 : a.C, c_base = blockIdx.y * Cs;
  const int C8 = Cs >> 3;
  const int chunk = threadIdx.x % C8;
  const int rows_per_iter = blockDim.x / C8;
  const int r0 = threadIdx.x / C8;
  constexpr bool dual = kDual;
  // The per-channel coefficients are computed ONCE per block into shared memory (statistics -> scale / shift: a sqrt, a
  // division and, on the engine's path, the conversion of the producers' fixed-point accumulators) instead of once per
  // thread for its 8 channels: with narrow layers 32 threads of a block share a channel chunk.
  extern __shared__ float s_coef[];  // [3][Cs]: scale, shift (+ shift2), scale2
  for (int cl = threadIdx.x; cl < Cs; cl += blockDim.x) {
    const int c = c_base + cl;
    float mean, var, scale, shift, scale2 = 0.f;
    bn_coeffs(a.train, a.sum, a.sq, a.stat_raw, a.M, a.gamma, a.beta, a.running_mean, a.running_var, c, scale, shift,
              mean, var);
    if (dual) {
      float shift2;
      bn_coeffs(a.train, a.sum2, a.sq2, a.stat_raw, a.M, a.gamma2, a.beta2, a.running_mean2, a.running_var2, c, scale2,
                shift2, mean, var);
      shift += shift2;
    }
    s_coef[cl] = scale;
    s_coef[Cs + cl] = shift;
    s_coef[2 * Cs + cl] = scale2;
  }
  __syncthreads();
  float sc[8], sh[8], sc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_coef[chunk * 8 + j];
    sh[j] = s_coef[Cs + chunk * 8 + j];
    sc2[j] = s_coef[2 * Cs + chunk * 8 + j];
  }
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  const bf16* __restrict__ y2 = reinterpret_cast<const bf16*>(a.y2);
  const bf16* __restrict__ res = reinterpret_cast<const bf16*>(a.residual);
  bf16* __restrict__ out = reinterpret_cast<bf16*>(a.a);
R3M_UNROLL(R3M_BN_APPLY_UNROLL)
  for (long long row = (long long)blockIdx.x * rows_per_iter + r0; row < a.M;
       row += (long long)gridDim.x * rows_per_iter) {
    const long long rr = a.reverse ? (long long)a.M - 1 - row : row;
    const long long off = rr * a.C + c_base + chunk * 8;
    // all loads of the row first (no control flow between them)
    F8 f = ld8_last(y + off);
    F8 t, r;
    if (kDual) t = ld8_last(y2 + off);
    if (kRes) r = ld8_last(res + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f.v[j] = fmaf(f.v[j], sc[j], sh[j]);
      if (kDual) f.v[j] = fmaf(t.v[j], sc2[j], f.v[j]);
      if (kRes) f.v[j] += r.v[j];
    }
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f.v[j] = fmaxf(f.v[j], 0.f);
    }
    st8(out + off, f);
    if (a.mask_out) {
      // the mask must describe the STORED (bf16-rounded) activation: a tiny positive value that rounds to +0 is off
      unsigned bits = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) bits |= (__bfloat162float(__float2bfloat16_rn(f.v[j])) > 0.f ? 1u : 0u) << j;
      a.mask_out[rr * (a.C >> 3) + (c_base >> 3) + chunk] = (uint8_t)bits;
    }
  }
  pdl_done();
  if (a.train && blockIdx.x == 0 && blockIdx.y == 0) {
    // every block has already read sum/sq into registers for its own coefficients; running stats are separate
    // buffers, so the in-place update below cannot race with other blocks
    bn_publish(a.C, a.M, a.sum, a.sq, a.stat_raw, a.save_mean, a.save_rstd, a.running_mean, a.running_var, a.update_running);
    if (dual)
      bn_publish(a.C, a.M, a.sum2, a.sq2, a.stat_raw, a.save_mean2, a.save_rstd2, a.running_mean2, a.running_var2,
                 a.update_running);
  }
}

__global__ void bn_fold_kernel(const BnFoldEntry* __restrict__ table) {
  pdl_sync();
  const BnFoldEntry e = table[blockIdx.x];
  for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
    const float sc = e.gamma[c] / sqrtf(e.running_var[c] + kBnEps);
    e.scale[c] = sc;
    e.shift[c] = e.beta[c] - e.running_mean[c] * sc;
  }
}

// ------------------------------------------------------------------------------------------------ stem pool
// argmax code per pooled element: scan-order index 0..8 of the window position that holds the maximum, or
// kPoolDead when the maximum is 0 (every input of the window was clipped by the ReLU: no gradient flows).
constexpr int kPoolDead = 15;

__device__ __forceinline__ __nv_bfloat162 as_bf162(uint32_t v) {
  __nv_bfloat162 r;
  *reinterpret_cast<uint32_t*>(&r) = v;
  return r;
}
__device__ __forceinline__ uint32_t as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

// The first version ran BN + ReLU in fp32 on all nine taps (700 instructions per 8-channel output; instruction bound
// at 1.7 TB/s).  relu(sc*y + sh) is monotone in y (increasing for sc >= 0, decreasing for sc < 0), so the window
// maximum of the activation is the activation of the window maximum of y (minimum for sc < 0: the sign bit of y is
// flipped for those channels).  The search therefore runs on the raw bf16 pairs with packed compare / max (exact),
// tracking the first maximum in scan order like ATen's max_pool2d, and the affine + ReLU is applied once.
__global__ void __launch_bounds__(224) stem_pool_kernel(const StemPoolArgs a) {
  pdl_sync();
  const int C8 = a.C >> 3;
  const int P = a.H / 2, Q = a.W / 2;
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  bf16* __restrict__ out = reinterpret_cast<bf16*>(a.a);
  // One block iteration = one pooled row (n, p); threads stride over (q, chunk) with 32-bit arithmetic only.  blockDim
  // is a multiple of C8, so a thread keeps its 8-channel chunk and the coefficients are hoisted.
  const int chunk = threadIdx.x % C8;
  __shared__ float s_sc[256], s_sh[256];  // C <= 256: coefficients once per block (see bn_apply_kernel)
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float mean, var;
    bn_coeffs(a.train, a.sum, a.sq, a.stat_raw, a.N * a.H * a.W, a.gamma, a.beta, a.running_mean, a.running_var, c, s_sc[c],
              s_sh[c], mean, var);
  }
  __syncthreads();
  float sc[8], sh[8];
  uint32_t flip[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = s_sc[chunk * 8 + j];
    sh[j] = s_sh[chunk * 8 + j];
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) flip[w] = (sc[2 * w] < 0.f ? 0x8000u : 0u) | (sc[2 * w + 1] < 0.f ? 0x80000000u : 0u);
  constexpr uint32_t kNegInf2 = 0xFF80FF80u;
  for (int row = blockIdx.x; row < a.N * P; row += gridDim.x) {
    const int n = row / P, p = row - n * P;
    for (int i = threadIdx.x; i < Q * C8; i += blockDim.x) {
      const int q = i / C8;
      const long long idx = (long long)row * (Q * C8) + i;
      // all nine loads first, from clamped coordinates
      uint4 f[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hc = max(2 * p - 1 + r, 0);  // 2p+1 <= H-1 always (H even)
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int wc = max(2 * q - 1 + t, 0);
          f[r * 3 + t] = __ldg(reinterpret_cast<const uint4*>(y + (((long long)n * a.H + hc) * a.W + wc) * a.C + chunk * 8));
        }
      }
      uint32_t best[4] = {kNegInf2, kNegInf2, kNegInf2, kNegInf2};
      uint32_t code[4] = {0u, 0u, 0u, 0u};  // one 16-bit code per channel, packed like the values
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const bool valid = !((k < 3 && p == 0) || (k % 3 == 0 && q == 0));  // taps at h = -1 / w = -1
        const uint32_t word[4] = {f[k].x, f[k].y, f[k].z, f[k].w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const uint32_t v = valid ? (word[w] ^ flip[w]) : kNegInf2;
          const uint32_t gt = __hgt2_mask(as_bf162(v), as_bf162(best[w]));  // strict: the first maximum wins
          best[w] = as_u32(__hmax2(as_bf162(best[w]), as_bf162(v)));
          code[w] = (code[w] & ~gt) | ((uint32_t)(k | (k << 16)) & gt);
        }
      }
      F8 o;
      int cd[8];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t yb = best[w] ^ flip[w];
        o.v[2 * w] = fmaxf(fmaf(bf16lo(yb), sc[2 * w], sh[2 * w]), 0.f);
        o.v[2 * w + 1] = fmaxf(fmaf(bf16hi(yb), sc[2 * w + 1], sh[2 * w + 1]), 0.f);
        cd[2 * w] = (int)(code[w] & 0xFFu);
        cd[2 * w + 1] = (int)(code[w] >> 16);
      }
      st8(out + idx * 8, o);
      if (a.ymax)
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(a.ymax) + idx * 8) =
            make_uint4(best[0] ^ flip[0], best[1] ^ flip[1], best[2] ^ flip[2], best[3] ^ flip[3]);
      if (a.argmax) {
        // "dead" must describe the STORED (bf16-rounded) activation, like the ReLU masks of bn_apply
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(__bfloat162float(__float2bfloat16_rn(o.v[j])) > 0.f)) cd[j] = kPoolDead;
        uint2 packed;
        packed.x = cd[0] | (cd[1] << 8) | (cd[2] << 16) | (cd[3] << 24);
        packed.y = cd[4] | (cd[5] << 8) | (cd[6] << 16) | (cd[7] << 24);
        *reinterpret_cast<uint2*>(a.argmax + idx * 8) = packed;
      }
    }
  }
  pdl_done();
  if (a.train && blockIdx.x == 0) {
    bn_publish(a.C, a.N * a.H * a.W, a.sum, a.sq, a.stat_raw, a.save_mean, a.save_rstd, a.running_mean, a.running_var,
               a.update_running);
  }
}

// Gradient w.r.t. the stem's BatchNorm output (ReLU mask applied) of one 8-channel chunk at conv-output pixel (h, w):
// the sum over the (up to four) pooling windows that cover the pixel of dA[window] where the window's argmax is (h, w).
// All eight loads (clamped coordinates) are issued before any use.
struct PoolGather {
  const bf16* dA;
  const uint8_t* argmax;
  int H, W, P, Q, C;
};
__device__ __forceinline__ F8 pool_gather(const PoolGather& g, int n, int h, int w, int chunk) {
  const int p0 = h >> 1, p1 = (h + 1) >> 1;  // windows p with 2p-1 <= h <= 2p+1
  const int q0 = w >> 1, q1 = (w + 1) >> 1;
  const int pc[2] = {p0, min(p1, g.P - 1)};
  const int qc[2] = {q0, min(q1, g.Q - 1)};
  const bool pv[2] = {true, p1 != p0 && p1 < g.P};
  const bool qv[2] = {true, q1 != q0 && q1 < g.Q};
  uint2 codes[4];
  F8 gr[4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long o = (((long long)n * g.P + pc[i]) * g.Q + qc[k]) * g.C + chunk * 8;
      codes[i * 2 + k] = __ldg(reinterpret_cast<const uint2*>(g.argmax + o));
      gr[i * 2 + k] = ld8(g.dA + o);
    }
  F8 acc;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      // position of (h, w) inside window (p, q); an invalid window can match no code
      const int my = (pv[i] && qv[k]) ? (h - (2 * pc[i] - 1)) * 3 + (w - (2 * qc[k] - 1)) : 255;
      const uint2 cd = codes[i * 2 + k];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cj = (j < 4 ? (cd.x >> (8 * j)) : (cd.y >> (8 * (j - 4)))) & 0xFF;
        acc.v[j] += (cj == my) ? gr[i * 2 + k].v[j] : 0.f;
      }
    }
  return acc;
}

__global__ void __launch_bounds__(224) maxpool_bwd_kernel(const bf16* __restrict__ dA,
                                                          const uint8_t* __restrict__ argmax, bf16* __restrict__ dz,
                                                          int N, int H, int W, int C) {
  pdl_sync();
  const int C8 = C >> 3;
  const PoolGather g{dA, argmax, H, W, H / 2, W / 2, C};
  const int chunk = threadIdx.x % C8;  // blockDim is a multiple of C8
  for (int row = blockIdx.x; row < N * H; row += gridDim.x)
    for (int i = threadIdx.x; i < W * C8; i += blockDim.x) {
      const int n = row / H, h = row - n * H;
      st8(dz + ((long long)row * (W * C8) + i) * 8, pool_gather(g, n, h, i / C8, chunk));
    }
}

// Stem backward, fused: maxpool backward + ReLU mask are recomputed on the fly from (dA_pool, argmax) inside both
// passes of the BatchNorm backward, so the [N,112,112,64] masked gradient is never written or re-read.
//   pass 1 (kApply = false): sums[c] += dz, sums[C + c] += dz * xhat
//   pass 2 (kApply = true) : dy = gamma*rstd * (dz - mean(dz) - xhat * mean(dz * xhat)); block 0 publishes dgamma, dbeta
// A thread owns the 2x2 input quad (2k+dh, 2m+dw) of one 8-channel chunk.  Exactly the four pooling windows
// (k+i, m+j), i, j in {0, 1}, cover the quad, and the argmax code that selects pixel (dh, dw) from window (i, j) is the
// compile-time constant (dh + 1 - 2i) * 3 + (dw + 1 - 2j): nine (window, pixel) pairs per channel instead of sixteen
// per-pixel gathers, and the window loads are shared by the four pixels (the per-pixel gather was instruction bound:
// 360 instructions per 8 channels, 1.5 TB/s).
// Block-level sums of per-thread 8-channel partials (s1: sum dz, s2: sum dz * (y - mean), scaled by rstd here) without
// shared-memory atomics: every thread parks its 16 values, thread c < 2C adds the copies of its channel in thread order.
// out[0..C) = sum s1, out[C..2C) = sum s2 * rstd.  blockDim.x <= 256 and a multiple of C / 8.
__device__ __forceinline__ void block_channel_sums(const float (&s1)[8], const float (&s2)[8], const float* __restrict__ rstd,
                                                   int chunk, int C, float* tab /* [256][16] */, float* out /* [2C] */) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    tab[threadIdx.x * 16 + j] = s1[j];
    tab[threadIdx.x * 16 + 8 + j] = s2[j] * rstd[chunk * 8 + j];
  }
  __syncthreads();
  const int C8 = C >> 3;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i / C, c = i - which * C;
    const int ch = c >> 3, j = c & 7;
    float acc = 0.f;
    for (int t = ch; t < (int)blockDim.x; t += C8) acc += tab[t * 16 + which * 8 + j];
    out[i] = acc;
  }
  __syncthreads();
}

template <bool kApply>
__global__ void __launch_bounds__(224) stem_bwd_kernel(const StemBwdArgs a) {
  pdl_sync();
  __shared__ float s_red[2 * 256];  // C <= 256
  __shared__ float s_tab[kApply ? 256 : 256 * 16];
  const int C8 = a.C >> 3;
  const int P = a.H / 2, Q = a.W / 2;
  const bf16* __restrict__ dA = reinterpret_cast<const bf16*>(a.dA);
  const uint8_t* __restrict__ am = a.argmax;
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  bf16* __restrict__ dy = reinterpret_cast<bf16*>(a.dy);
  const int chunk = threadIdx.x % C8;  // blockDim is a multiple of C8
  const float inv_m = 1.0f / ((float)a.N * a.H * a.W);
  float c0[8], c1[8], c2[8];  // reduce: mean, s1, s2;  apply: cA, cB, cC
  if (kApply) {  // constants once per block (the sums may be fixed-point accumulators: see bn_bwd_apply_kernel)
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      const float mean = a.mean[c], rstd = a.rstd[c];
      const float mdz = bsum_at(a.sums, a.sums_raw, c) * inv_m, mdzx = bsum_at(a.sums, a.sums_raw, a.C + c) * inv_m;
      const float A = a.gamma[c] * rstd, B = -A * rstd * mdzx;
      s_red[c] = A;
      s_red[a.C + c] = B;
      s_tab[c] = -A * mdz - B * mean;
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = chunk * 8 + j;
    if (kApply) {
      c0[j] = s_red[c];
      c1[j] = s_red[a.C + c];
      c2[j] = s_tab[c];
    } else {
      c0[j] = a.mean[c];
      c1[j] = 0.f;
      c2[j] = 0.f;
    }
  }
  for (int row = blockIdx.x; row < a.N * P; row += gridDim.x) {
    const int n = row / P, k = row - n * P;
    for (int it = threadIdx.x; it < Q * C8; it += blockDim.x) {
      const int m = it / C8;
      // ---- all loads first: 4 input pixels, 4 windows (clamped; out-of-range windows are disabled below)
      const int k1 = min(k + 1, P - 1), m1 = min(m + 1, Q - 1);
      const long long ybase = (((long long)n * a.H + 2 * k) * a.W + 2 * m) * a.C + chunk * 8;
      uint4 yv[4];
      yv[0] = __ldg(reinterpret_cast<const uint4*>(y + ybase));
      yv[1] = __ldg(reinterpret_cast<const uint4*>(y + ybase + a.C));
      yv[2] = __ldg(reinterpret_cast<const uint4*>(y + ybase + (long long)a.W * a.C));
      yv[3] = __ldg(reinterpret_cast<const uint4*>(y + ybase + (long long)a.W * a.C + a.C));
      const long long w00 = (((long long)n * P + k) * Q + m) * a.C + chunk * 8;
      const long long w01 = (((long long)n * P + k) * Q + m1) * a.C + chunk * 8;
      const long long w10 = (((long long)n * P + k1) * Q + m) * a.C + chunk * 8;
      const long long w11 = (((long long)n * P + k1) * Q + m1) * a.C + chunk * 8;
      uint2 cw[4];
      uint4 gw[4];
      cw[0] = __ldg(reinterpret_cast<const uint2*>(am + w00));
      cw[1] = __ldg(reinterpret_cast<const uint2*>(am + w01));
      cw[2] = __ldg(reinterpret_cast<const uint2*>(am + w10));
      cw[3] = __ldg(reinterpret_cast<const uint2*>(am + w11));
      gw[0] = __ldg(reinterpret_cast<const uint4*>(dA + w00));
      gw[1] = __ldg(reinterpret_cast<const uint4*>(dA + w01));
      gw[2] = __ldg(reinterpret_cast<const uint4*>(dA + w10));
      gw[3] = __ldg(reinterpret_cast<const uint4*>(dA + w11));
      // a window beyond the last pooled row / column selects nothing: 0xFF matches no code
      if (m + 1 >= Q) cw[1] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      if (k + 1 >= P) cw[2] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      if (m + 1 >= Q || k + 1 >= P) cw[3] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      // ---- masked gradient of the four pixels: dz[dh * 2 + dw][channel]
      float dz[4][8];
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) {
        const int i = wi >> 1, jw = wi & 1;
        const uint32_t gwords[4] = {gw[wi].x, gw[wi].y, gw[wi].z, gw[wi].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t cj = ((j < 4 ? cw[wi].x : cw[wi].y) >> (8 * (j & 3))) & 0xFFu;
          const float g = (j & 1) ? bf16hi(gwords[j >> 1]) : bf16lo(gwords[j >> 1]);
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const int dh = px >> 1, dw = px & 1;
            const int r = dh + 1 - 2 * i, t = dw + 1 - 2 * jw;  // position of the pixel inside the window
            if (wi == 0) dz[px][j] = 0.f;
            if (r >= 0 && t >= 0) dz[px][j] += (cj == (uint32_t)(r * 3 + t)) ? g : 0.f;
          }
        }
      }
      // ---- consume
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const uint32_t ywords[4] = {yv[px].x, yv[px].y, yv[px].z, yv[px].w};
        if (kApply) {
          F8 o;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float yy = (j & 1) ? bf16hi(ywords[j >> 1]) : bf16lo(ywords[j >> 1]);
            o.v[j] = fmaf(c0[j], dz[px][j], fmaf(c1[j], yy, c2[j]));
          }
          st8(dy + ybase + (long long)(px >> 1) * a.W * a.C + (px & 1) * a.C, o);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float yy = (j & 1) ? bf16hi(ywords[j >> 1]) : bf16lo(ywords[j >> 1]);
            c1[j] += dz[px][j];
            c2[j] = fmaf(dz[px][j], yy - c0[j], c2[j]);
          }
        }
      }
    }
  }
  pdl_done();
  if (!kApply) {
    block_channel_sums(c1, c2, a.rstd, chunk, a.C, s_tab, s_red);
    float* sums = a.sums;
    if (a.sums_raw) {
      unsigned long long* acc = reinterpret_cast<unsigned long long*>(sums);
      for (int i = threadIdx.x; i < 2 * a.C; i += blockDim.x) fx_add(acc + kFxWords * i, s_red[i]);
    } else {
      det_grid_reduce(s_red, 2 * a.C, a.det, [=](int i, float v) { sums[i] = v; });
    }
  } else if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      if (a.dbeta) a.dbeta[c] = bsum_at(a.sums, a.sums_raw, c);
      if (a.dgamma) a.dgamma[c] = bsum_at(a.sums, a.sums_raw, a.C + c);
    }
  }
}

// Reduce pass of the stem backward over the POOLED elements (needs ymax from the forward): every live window routes its
// gradient to exactly one conv-output pixel, whose raw value is ymax, so
//   sum_pixels dz = sum_windows dA * live,   sum_pixels dz * (y - mean) = sum_windows dA * live * (ymax - mean).
// Reads 320 MB instead of the 706 MB (and the 2x2-quad matching) of the per-pixel formulation.
__global__ void __launch_bounds__(256) stem_bwd_reduce_pooled_kernel(const StemBwdArgs a) {
  pdl_sync();
  __shared__ float s_red[2 * 256];
  __shared__ float s_tab[256 * 16];
  const int C8 = a.C >> 3;
  const int chunk = threadIdx.x % C8;  // the grid stride is a multiple of C8
  const bf16* __restrict__ dA = reinterpret_cast<const bf16*>(a.dA);
  const bf16* __restrict__ ym = reinterpret_cast<const bf16*>(a.ymax);
  const size_t total = (size_t)a.N * (a.H / 2) * (a.W / 2) * C8;
  float mean[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = a.mean[chunk * 8 + j];
    s1[j] = 0.f;
    s2[j] = 0.f;
  }
#pragma unroll 2
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const F8 g = ld8(dA + idx * 8);
    const F8 y = ld8(ym + idx * 8);
    const uint2 cd = __ldg(reinterpret_cast<const uint2*>(a.argmax + idx * 8));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t cj = ((j < 4 ? cd.x : cd.y) >> (8 * (j & 3))) & 0xFFu;
      const float gj = (cj != (uint32_t)kPoolDead) ? g.v[j] : 0.f;
      s1[j] += gj;
      s2[j] = fmaf(gj, y.v[j] - mean[j], s2[j]);
    }
  }
  pdl_done();
  block_channel_sums(s1, s2, a.rstd, chunk, a.C, s_tab, s_red);
  float* sums = a.sums;
  if (a.sums_raw) {
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(sums);
    for (int i = threadIdx.x; i < 2 * a.C; i += blockDim.x) fx_add(acc + kFxWords * i, s_red[i]);
  } else {
    det_grid_reduce(s_red, 2 * a.C, a.det, [=](int i, float v) { sums[i] = v; });
  }
}

// ------------------------------------------------------------------------------------------------ avg pool
__global__ void avgpool_fwd_kernel(const bf16* __restrict__ a, float* __restrict__ out, int HW, int C) {
  pdl_sync();
  const int n = blockIdx.x;
  const float inv = 1.0f / (float)HW;
  for (int c2 = threadIdx.x; c2 < C / 2; c2 += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    const uint32_t* base = reinterpret_cast<const uint32_t*>(a + (long long)n * HW * C) + c2;
    for (int i = 0; i < HW; ++i) {
      const uint32_t v = base[(long long)i * (C / 2)];
      s0 += bf16lo(v);
      s1 += bf16hi(v);
    }
    out[(long long)n * C + 2 * c2] = s0 * inv;
    out[(long long)n * C + 2 * c2 + 1] = s1 * inv;
  }
}

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dE, bf16* __restrict__ dA, int N,
                                                          int HW, int C) {
  pdl_sync();
  const int C8 = C >> 3;
  const long long total = (long long)N * HW * C8;
  const float inv = 1.0f / (float)HW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(idx % C8);
    const int n = (int)(idx / ((long long)C8 * HW));
    const float4 g0 = *reinterpret_cast<const float4*>(dE + (long long)n * C + chunk * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(dE + (long long)n * C + chunk * 8 + 4);
    F8 f;
    f.v[0] = g0.x * inv;
    f.v[1] = g0.y * inv;
    f.v[2] = g0.z * inv;
    f.v[3] = g0.w * inv;
    f.v[4] = g1.x * inv;
    f.v[5] = g1.y * inv;
    f.v[6] = g1.z * inv;
    f.v[7] = g1.w * inv;
    st8(dA + idx * 8, f);
  }
}

// ------------------------------------------------------------------------------------------------ BN backward
// ReLU-mask source of the backward kernels (compile-time, so that the row loop has no control flow between its loads:
// a branch between the loads serialises them and the kernels become latency bound — measured 3.8 vs 6.2 TB/s)
enum { kMaskNone = 0, kMaskAct = 1, kMaskBits = 2 };
constexpr int kBwdSlice = 64;  // channels per block of bn_bwd_reduce_kernel

template <int kMask>
__device__ __forceinline__ void mask_gradient(F8& g, const F8& act, unsigned bits) {
  if (kMask == kMaskAct) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g.v[j] = act.v[j] > 0.f ? g.v[j] : 0.f;
  } else if (kMask == kMaskBits) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g.v[j] = ((bits >> j) & 1u) ? g.v[j] : 0.f;
  }
}

// Reduce pass of one block: (bx, by) of a (gx, C / 64) grid.  Leaves the block's partial sums — V = kQ * 64 floats, slice-local
// [0, 64) sum(dz), [64, 128) sum(dz * xhat), [128, 192) sum(dz * xhat2) — in s_part[0 .. V / 4).
template <bool kDual, int kMask>
__device__ __forceinline__ void bn_bwd_reduce_body(const BnBwdArgs& a, float4* s_part, int bx, int by, int gx,
                                                   bool signal_dependents) {
  // s_part: [rows_per_iter][kQ * Cs / 4]: per-thread partials of sum(dz), sum(dz*xhat)[, 2]
  constexpr int kQ = kDual ? 3 : 2;
  // The tensor is cut into channel slices of kBwdSlice = 64 channels (blockIdx.y): a block covers 32 rows per iteration
  // (one fully used 128-byte line per row) and its partial vector is kQ * 64 floats instead of kQ * C, so the number of
  // fixed-point reductions the grid issues (blocks x partial length: the kernel's tail) is C / 64 times smaller.
  const int Cs = kBwdSlice, c_base = by * Cs;
  const int C8 = Cs >> 3, ld8c = a.C >> 3;
  const int chunk = threadIdx.x % C8;
  const int rows_per_iter = blockDim.x / C8;
  const int r0 = threadIdx.x / C8;
  // s2 / s3 accumulate sum(dz * (y - mean)); the 1/std factor is applied once at the end
  float mean[8], mean2[8], s1[8], s2[8], s3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = a.mean[c_base + chunk * 8 + j];
    mean2[j] = kDual ? a.mean2[c_base + chunk * 8 + j] : 0.f;
    s1[j] = 0.f;
    s2[j] = 0.f;
    s3[j] = 0.f;
  }
  const bf16* __restrict__ dA = reinterpret_cast<const bf16*>(a.dA);
  const bf16* __restrict__ act = reinterpret_cast<const bf16*>(a.a);
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  const bf16* __restrict__ y2 = reinterpret_cast<const bf16*>(a.y2);
  const uint8_t* __restrict__ mask = a.mask;
  // Four rows per trip with every load issued before the first use: the small layers give a thread 16-60 rows, and one
  // dependent DRAM round trip per row made them latency bound (33 MB in 16 us, ncu).  Rows are still accumulated in
  // row order (the sums do not depend on the unroll factor).
  const long long stride = (long long)gx * rows_per_iter;
  long long row = (long long)bx * rows_per_iter + r0;
  auto accumulate = [&](const uint4& ug, const uint4& uy, const uint4& um, const uint4& ut, unsigned bits) {
    F8 g = cvt8(ug);
    const F8 yy = cvt8(uy);
    F8 m, t;
    if (kMask == kMaskAct) m = cvt8(um);
    if (kDual) t = cvt8(ut);
    mask_gradient<kMask>(g, m, bits);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] += g.v[j];
      s2[j] = fmaf(g.v[j], yy.v[j] - mean[j], s2[j]);
      if (kDual) s3[j] = fmaf(g.v[j], t.v[j] - mean2[j], s3[j]);
    }
  };
  constexpr int kU = 4;
  for (; !a.single_rows && row + (kU - 1) * stride < a.M; row += kU * stride) {
    uint4 ug[kU], uy[kU], um[kU], ut[kU];
    unsigned bits[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long r = a.rev_reduce ? (long long)a.M - 1 - (row + u * stride) : row + u * stride;
      const long long off = r * a.C + c_base + chunk * 8;
      ug[u] = ld8_raw(dA + off);
      uy[u] = ld8_raw(y + off);
      um[u] = make_uint4(0, 0, 0, 0);
      ut[u] = make_uint4(0, 0, 0, 0);
      bits[u] = 0;
      if (kMask == kMaskAct) um[u] = ld8_raw(act + off);
      if (kMask == kMaskBits) bits[u] = __ldg(mask + r * ld8c + (c_base >> 3) + chunk);
      if (kDual) ut[u] = ld8_raw(y2 + off);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) accumulate(ug[u], uy[u], um[u], ut[u], bits[u]);
  }
  for (; row < a.M; row += stride) {
    const long long r = a.rev_reduce ? (long long)a.M - 1 - row : row;
    const long long off = r * a.C + c_base + chunk * 8;
    const uint4 ug = ld8_raw(dA + off), uy = ld8_raw(y + off);
    uint4 um = make_uint4(0, 0, 0, 0), ut = um;
    unsigned bits = 0;
    if (kMask == kMaskAct) um = ld8_raw(act + off);
    if (kMask == kMaskBits) bits = __ldg(mask + r * ld8c + (c_base >> 3) + chunk);
    if (kDual) ut = ld8_raw(y2 + off);
    accumulate(ug, uy, um, ut, bits);
  }
  if (signal_dependents) pdl_done();
  // Block reduction without atomics: every thread parks its partials, one thread per four outputs adds the
  // rows_per_iter copies in row order, and the grid-wide sum is the ordered two-level reduction above.
  const int q4 = kQ * Cs / 4;  // float4 slots per partial row
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s2[j] *= a.rstd[c_base + chunk * 8 + j];
    if (kDual) s3[j] *= a.rstd2[c_base + chunk * 8 + j];
  }
  float4* mine = s_part + (size_t)r0 * q4 + chunk * 2;
  mine[0] = make_float4(s1[0], s1[1], s1[2], s1[3]);
  mine[1] = make_float4(s1[4], s1[5], s1[6], s1[7]);
  mine[Cs / 4] = make_float4(s2[0], s2[1], s2[2], s2[3]);
  mine[Cs / 4 + 1] = make_float4(s2[4], s2[5], s2[6], s2[7]);
  if (kDual) {
    mine[Cs / 2] = make_float4(s3[0], s3[1], s3[2], s3[3]);
    mine[Cs / 2 + 1] = make_float4(s3[4], s3[5], s3[6], s3[7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < q4; i += blockDim.x) {
    float4 acc = s_part[i];
    for (int r = 1; r < rows_per_iter; ++r) {
      const float4 v = s_part[(size_t)r * q4 + i];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    s_part[i] = acc;  // row 0 of the table becomes the block's partial (slot i is touched by this thread only)
  }
  __syncthreads();
}

// slice-local partial i of channel slice `by` -> the layer's fixed-point accumulators (engine path: added, zero per step)
template <bool kDual>
__device__ __forceinline__ void bn_bwd_emit_raw(const BnBwdArgs& a, const float4* s_part, int by) {
  constexpr int kQ = kDual ? 3 : 2;
  const int Cs = kBwdSlice, c_base = by * Cs, C = a.C, V = kQ * Cs;
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(a.sums);
  unsigned long long* acc2 = reinterpret_cast<unsigned long long*>(a.sums2);
  const float* part = reinterpret_cast<const float*>(s_part);
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const int which = i / Cs, c = c_base + i - which * Cs;
    fx_add(which == 2 ? acc2 + kFxWords * c : acc + kFxWords * (which * C + c), part[i]);
  }
}

template <bool kDual, int kMask>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  pdl_sync();
  extern __shared__ float4 s_part[];
  constexpr int kQ = kDual ? 3 : 2;
  bn_bwd_reduce_body<kDual, kMask>(a, s_part, blockIdx.x, blockIdx.y, gridDim.x, true);
  const int Cs = kBwdSlice, c_base = blockIdx.y * Cs;
  // slice-local floats [0, Cs) -> sums[c], [Cs, 2Cs) -> sums[C + c], [2Cs, 3Cs) -> sums2[c]; written, not accumulated
  float* sums = a.sums;
  float* sums2 = a.sums2;
  const int C = a.C, V = kQ * Cs;
  if (a.sums_raw) {
    // engine path: straight into the layer's own fixed-point accumulators (zeroed per step); bn_bwd_apply converts
    bn_bwd_emit_raw<kDual>(a, s_part, blockIdx.y);
    return;
  }
  DetScratch d;
  d.scratch = a.det.scratch + (size_t)blockIdx.y * V * 2 * kFxWords;  // V accumulators of kFxWords 64-bit words
  d.tickets = a.det.tickets + blockIdx.y;
  det_grid_reduce(reinterpret_cast<const float*>(s_part), V, d, [=](int i, float v) {
    const int which = i / Cs, c = c_base + i - which * Cs;
    if (which == 0)
      sums[c] = v;
    else if (which == 1)
      sums[C + c] = v;
    else
      sums2[c] = v;
  });
}

// Apply pass of one block: (bx, by) of a (gx, ny) grid.
template <bool kDual, int kMask, bool kDz>
__device__ __forceinline__ void bn_bwd_apply_body(const BnBwdArgs& a, float* s_coef, int bx, int by, int gx, int ny,
                                                  bool signal_dependents) {
  const int Cs = ny > 1 ? 256 : a.C, c_base = by * Cs;  // channel slices for wide layers (see bn_apply)
  const int C8 = Cs >> 3;
  const int chunk = threadIdx.x % C8;
  const int rows_per_iter = blockDim.x / C8;
  const int r0 = threadIdx.x / C8;
  const float inv_m = 1.0f / (float)a.M;
  // dy = gamma*rstd * (dz - mean(dz) - xhat * mean(dz*xhat)),  xhat = (y - mean) * rstd
  //    = cA * dz + cB * y + cC   with per-channel constants (three registers per channel instead of five), computed once
  // per block into shared memory (on the engine's path this includes converting the reduce pass' fixed-point sums)
  // s_coef: [kDual ? 6 : 3][Cs]
  for (int cl = threadIdx.x; cl < Cs; cl += blockDim.x) {
    const int c = c_base + cl;
    const float mean = a.mean[c], rstd = a.rstd[c];
    const float mdz = bsum_at(a.sums, a.sums_raw, c) * inv_m, mdzx = bsum_at(a.sums, a.sums_raw, a.C + c) * inv_m;
    const float A = a.gamma[c] * rstd;
    const float B = -A * rstd * mdzx;
    s_coef[cl] = A;
    s_coef[Cs + cl] = B;
    s_coef[2 * Cs + cl] = -A * mdz - B * mean;
    if (kDual) {
      const float mean2 = a.mean2[c], rstd2 = a.rstd2[c];
      const float mdzx2 = bsum_at(a.sums2, a.sums_raw, c) * inv_m;
      const float A2 = a.gamma2[c] * rstd2;
      const float B2 = -A2 * rstd2 * mdzx2;
      s_coef[3 * Cs + cl] = A2;
      s_coef[4 * Cs + cl] = B2;
      s_coef[5 * Cs + cl] = -A2 * mdz - B2 * mean2;
    }
  }
  __syncthreads();
  float cA[8], cB[8], cC[8], cA2[8], cB2[8], cC2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = chunk * 8 + j;
    cA[j] = s_coef[c];
    cB[j] = s_coef[Cs + c];
    cC[j] = s_coef[2 * Cs + c];
    if (kDual) {
      cA2[j] = s_coef[3 * Cs + c];
      cB2[j] = s_coef[4 * Cs + c];
      cC2[j] = s_coef[5 * Cs + c];
    } else {
      cA2[j] = cB2[j] = cC2[j] = 0.f;
    }
  }
  const bf16* __restrict__ dA = reinterpret_cast<const bf16*>(a.dA);
  const bf16* __restrict__ act = reinterpret_cast<const bf16*>(a.a);
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  const bf16* __restrict__ y2 = reinterpret_cast<const bf16*>(a.y2);
  const uint8_t* __restrict__ mask = a.mask;
  bf16* __restrict__ dy = reinterpret_cast<bf16*>(a.dy);
  bf16* __restrict__ dy2 = reinterpret_cast<bf16*>(a.dy2);
  bf16* __restrict__ dzo = reinterpret_cast<bf16*>(a.dz_out);
R3M_UNROLL(R3M_BN_BWD_UNROLL)
  for (long long row = (long long)bx * rows_per_iter + r0; row < a.M; row += (long long)gx * rows_per_iter) {
    const long long rr = a.rev_apply ? (long long)a.M - 1 - row : row;
    const long long off = rr * a.C + c_base + chunk * 8;
    F8 g = ld8_last(dA + off);
    const F8 yy = ld8_last(y + off);
    F8 m, t;
    unsigned bits = 0;
    if (kMask == kMaskAct) m = ld8(act + off);
    if (kMask == kMaskBits) bits = __ldg(mask + rr * (a.C >> 3) + (c_base >> 3) + chunk);
    if (kDual) t = ld8_last(y2 + off);
    mask_gradient<kMask>(g, m, bits);
    if (kDz) st8(dzo + off, g);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = fmaf(cA[j], g.v[j], fmaf(cB[j], yy.v[j], cC[j]));
    st8(dy + off, o);
    if (kDual) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = fmaf(cA2[j], g.v[j], fmaf(cB2[j], t.v[j], cC2[j]));
      st8(dy2 + off, o);
    }
  }
  if (signal_dependents) pdl_done();
  if (bx == 0 && by == 0) {
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      if (a.dbeta) a.dbeta[c] = bsum_at(a.sums, a.sums_raw, c);
      if (a.dgamma) a.dgamma[c] = bsum_at(a.sums, a.sums_raw, a.C + c);
      if (kDual) {
        if (a.dbeta2) a.dbeta2[c] = bsum_at(a.sums, a.sums_raw, c);
        if (a.dgamma2) a.dgamma2[c] = bsum_at(a.sums2, a.sums_raw, c);
      }
    }
  }
}

template <bool kDual, int kMask, bool kDz>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs a) {
  pdl_sync();
  extern __shared__ float s_coef[];
  bn_bwd_apply_body<kDual, kMask, kDz>(a, s_coef, blockIdx.x, blockIdx.y, gridDim.x, gridDim.y, true);
}

// Reduce + grid barrier + apply in ONE launch, for layers whose gradient and raw output (2 x M x C bf16) stay in the
// 126 MB L2 between the two passes: the second pass never goes to DRAM, and the launch, the drain of the reduce grid and
// the refill of the apply grid (about 10 us for a 16-33 MB layer) disappear.  Every block of the grid is resident by
// construction (the host sizes the grid to one wave of THIS kernel and launches it cooperatively, so work on other
// streams cannot keep part of the grid unscheduled while the rest spins); the barrier is a counter in the step's zeroed
// region.  A watchdog turns a lost block into an error flag instead of a hang.
template <int kMask>
__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(const BnBwdArgs a, int slices_r, int slices_a, int* counter,
                                                           int* error_flag) {
  pdl_sync();
  extern __shared__ float4 s_dyn[];
  const int b = blockIdx.x, G = gridDim.x;
  bn_bwd_reduce_body<false, kMask>(a, s_dyn, b / slices_r, b % slices_r, G / slices_r, false);
  bn_bwd_emit_raw<false>(a, s_dyn, b % slices_r);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1);
    const uint64_t t0 = globaltimer_ns();
    while (*reinterpret_cast<volatile int*>(counter) < G) {
      if (globaltimer_ns() - t0 > R3M_WAIT_TIMEOUT_NS) {
        atomicExch(error_flag, 31);
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
  pdl_done();
  bn_bwd_apply_body<false, kMask, false>(a, reinterpret_cast<float*>(s_dyn), b / slices_a, b % slices_a, G / slices_a,
                                         slices_a, false);
}

// ------------------------------------------------------------------------------------------------ Adam / casts
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   bf16* __restrict__ pb, size_t n, float lr, float beta1, float beta2,
                                                   float eps, float bc1, float bc2_sqrt, float grad_scale) {
  pdl_sync();
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - step_size * (mi / denom);
    p[i] = pi;
    if (pb) pb[i] = __float2bfloat16_rn(pi);
  }
}

// four parameters per thread: 16-byte loads / stores of p, g, m, v and one 8-byte store of the bf16 copy
__global__ void __launch_bounds__(256) adam_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                        float4* __restrict__ m, float4* __restrict__ v,
                                                        uint2* __restrict__ pb, size_t n4, float lr, float beta1,
                                                        float beta2, float eps, float bc1, float bc2_sqrt,
                                                        float grad_scale) {
  pdl_sync();
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 g4 = g[i];
    float4 m4 = m[i], v4 = v[i], p4 = p[i];
    const float gs[4] = {g4.x * grad_scale, g4.y * grad_scale, g4.z * grad_scale, g4.w * grad_scale};
    float ms[4] = {m4.x, m4.y, m4.z, m4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // the same operation order as adam_kernel: results are bit-identical
      ms[k] = beta1 * ms[k] + (1.f - beta1) * gs[k];
      vs[k] = beta2 * vs[k] + (1.f - beta2) * gs[k] * gs[k];
      const float denom = sqrtf(vs[k]) / bc2_sqrt + eps;
      ps[k] = ps[k] - step_size * (ms[k] / denom);
    }
    m[i] = make_float4(ms[0], ms[1], ms[2], ms[3]);
    v[i] = make_float4(vs[0], vs[1], vs[2], vs[3]);
    p[i] = make_float4(ps[0], ps[1], ps[2], ps[3]);
    if (pb) pb[i] = make_uint2(pack_bf16x2(ps[0], ps[1]), pack_bf16x2(ps[2], ps[3]));
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                        size_t n) {
  pdl_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// the step's small host-side inputs (permutations, sentence mask / embedding) are pulled from mapped pinned host memory
// by the SMs: a cudaMemcpyAsync would queue behind the NEXT batch's 193 MB frame upload on the H2D copy engine
__global__ void __launch_bounds__(256) pull_host_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                        size_t n16) {
  pdl_sync();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------ filter packers
struct TapList {
  int t[16];
};
__global__ void pack_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int T, int Cin, int nt,
                                  TapList taps) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int src = taps.t[t];
  const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int k = k0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (k < Cout && c < Cin) ? w[((size_t)k * T + src) * Cin + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, k = k0 + threadIdx.x;
    if (c < Cin && k < Cout) out[((size_t)c * nt + t) * Cout + k] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

// all dgrad re-packs of a network in one launch: the block index is mapped to (entry, tile) through block_begin
__global__ void pack_dgrad_multi_kernel(const PackDgradEntry* __restrict__ table, int entries) {
  pdl_sync();
  __shared__ float tile[32][33];
  int lo = 0, hi = entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackDgradEntry& e = table[lo];
  const int local = blockIdx.x - e.block_begin;
  const int gx = (e.Cin + 31) / 32, gy = (e.Cout + 31) / 32;
  const int bx = local % gx, by = (local / gx) % gy, t = local / (gx * gy);
  const int src = e.taps[t];
  const int c0 = bx * 32, k0 = by * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int k = k0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (k < e.Cout && c < e.Cin) ? e.w[((size_t)k * e.T + src) * e.Cin + c] : 0.f;
  }
  __syncthreads();
  bf16* __restrict__ out = reinterpret_cast<bf16*>(e.out);
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, k = k0 + threadIdx.x;
    if (c < e.Cin && k < e.Cout) out[((size_t)c * e.nt + t) * e.Cout + k] = __float2bfloat16_rn(tile[threadIdx.x][j]);
  }
}

__device__ __forceinline__ bool stem_map(int k, int rr, int j, int& oihw) {
  const int kw = j >> 4, sub = (j >> 2) & 3, c = j & 3;
  const int dy = sub >> 1, dx = sub & 1;
  const int r = 2 * rr + dy - 1, s = 2 * kw + dx - 1;
  if (c >= 3 || r < 0 || r >= 7 || s < 0 || s >= 7) return false;
  oihw = ((k * 3 + c) * 7 + r) * 7 + s;
  return true;
}
__global__ void stem_pack_kernel(const float* __restrict__ w, bf16* __restrict__ wp) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 4 * 64) return;
  const int j = idx & 63, rr = (idx >> 6) & 3, k = idx >> 8;
  int o;
  wp[idx] = __float2bfloat16_rn(stem_map(k, rr, j, o) ? w[o] : 0.f);
}
__global__ void stem_pack_f32_kernel(const float* __restrict__ w, float* __restrict__ wp) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 4 * 64) return;
  const int j = idx & 63, rr = (idx >> 6) & 3, k = idx >> 8;
  int o;
  uint32_t r;
  const float v = stem_map(k, rr, j, o) ? w[o] : 0.f;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  wp[idx] = __uint_as_float(r);
}
__global__ void stem_unpack_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 4 * 64) return;
  const int j = idx & 63, rr = (idx >> 6) & 3, k = idx >> 8;
  int o;
  if (stem_map(k, rr, j, o)) dw[o] = dwp[idx];
}

inline int grid_for(long long work_items, int threads, int max_blocks) {
  long long b = (work_items + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)std::min<long long>(b, max_blocks);
}

// Grid-stride kernels run as ONE resident wave: SMs x (CTAs per SM the kernel's registers / shared memory allow).  A
// fixed cap of 8 CTAs per SM gave 72-120-register kernels 2.7-4 unequal waves (measured 4.0 vs 5.5 TB/s).
// R3M_GRID_WAVES (experiments) scales the wave count.
template <auto Kernel>
int resident_blocks(int threads, size_t smem = 0) {
  static int cache[64] = {0};  // per 1 KB bucket of dynamic shared memory (the occupancy depends on it)
  int& blocks = cache[std::min<size_t>(63, (smem + 1023) >> 10)];
  if (blocks == 0) {
    int occ = 0, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, Kernel, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    const char* e = std::getenv("R3M_GRID_WAVES");
    const int waves = e ? std::max(1, atoi(e)) : 1;
    blocks = occ * sms * waves;
  }
  return blocks;
}

}  // namespace

cudaError_t launch_preprocess_stem(const void* obs, int format, void* xs, int N, cudaStream_t s) {
  const long long total = (long long)N * 112 * 112 * 4;
  const int grid = grid_for(total, 256, 148 * 16);
  bf16* dst = reinterpret_cast<bf16*>(xs);
  switch (format) {
    case kObsF32NCHW: launch_kernel(preprocess_stem_kernel<kObsF32NCHW>, grid, 256, 0, s, obs, dst, N); break;
    case kObsU8NCHW: launch_kernel(preprocess_stem_kernel<kObsU8NCHW>, grid, 256, 0, s, obs, dst, N); break;
    case kObsU8NHWC: launch_kernel(preprocess_stem_kernel<kObsU8NHWC>, grid, 256, 0, s, obs, dst, N); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_crop_resize(const uint8_t* src, int nhwc, const int* boxes, float* out, int N, int H, int W,
                               cudaStream_t s) {
  if (N < 1 || H < 1 || W < 1) return cudaErrorInvalidValue;
  const int grid = grid_for((long long)N * 224 * 224, 256, 148 * 16);
  if (nhwc)
    launch_kernel(crop_resize_kernel<true>, grid, 256, 0, s, src, boxes, out, N, H, W);
  else
    launch_kernel(crop_resize_kernel<false>, grid, 256, 0, s, src, boxes, out, N, H, W);
  return cudaGetLastError();
}

cudaError_t launch_bn_fold(const BnFoldEntry* table_dev, int entries, cudaStream_t s) {
  launch_kernel(bn_fold_kernel, entries, 256, 0, s, table_dev);
  return cudaGetLastError();
}

cudaError_t launch_bn_apply(const BnApplyArgs& a, cudaStream_t s) {
  if (a.C % 8 != 0 || a.C > 2048 || 256 % (a.C / 8) != 0) return cudaErrorInvalidValue;
  if (a.y2 && a.residual) return cudaErrorInvalidValue;  // a block tail has either an identity or a downsample branch
  const int slices = 1, Cs = a.C / slices;  // (channel slices measured slower for this kernel: +4-10 us on the wide tails)
  const int rows_per_iter = 256 / (Cs / 8);
  const size_t coef_smem = 3 * (size_t)Cs * sizeof(float);
#define R3M_LAUNCH(D, R)                                                                                               \
  launch_kernel(bn_apply_kernel<D, R>,                                                                                 \
                dim3(std::max(1, grid_for(a.M, rows_per_iter, resident_blocks<bn_apply_kernel<D, R>>(256, coef_smem) / \
                                                                  slices)),                                          \
                     slices),                                                                                        \
                256, coef_smem, s, a)
  if (a.y2)
    R3M_LAUNCH(true, false);
  else if (a.residual)
    R3M_LAUNCH(false, true);
  else
    R3M_LAUNCH(false, false);
#undef R3M_LAUNCH
  return cudaGetLastError();
}

// Stem kernels: 224 threads = 28 pixels x 8 chunks, so a 112-pixel row is exactly 4 (56: 2) block iterations.
constexpr int kStemThreads = 224;

cudaError_t launch_stem_bn_relu_maxpool(const StemPoolArgs& a, cudaStream_t s) {
  if (a.C % 8 != 0 || kStemThreads % (a.C / 8) != 0 || (a.H & 1) || (a.W & 1)) return cudaErrorInvalidValue;
  const int rows = a.N * (a.H / 2);
  launch_kernel(stem_pool_kernel, std::min(rows, resident_blocks<stem_pool_kernel>(kStemThreads)), kStemThreads, 0, s, a);
  return cudaGetLastError();
}

cudaError_t launch_maxpool_bwd(const void* dA, const uint8_t* argmax, void* dz, int N, int H, int W, int C,
                               cudaStream_t s) {
  if (C % 8 != 0 || kStemThreads % (C / 8) != 0 || (H & 1) || (W & 1)) return cudaErrorInvalidValue;
  launch_kernel(maxpool_bwd_kernel, std::min(N * H, resident_blocks<maxpool_bwd_kernel>(kStemThreads)), kStemThreads, 0,
                s, reinterpret_cast<const bf16*>(dA), argmax, reinterpret_cast<bf16*>(dz), N, H, W, C);
  return cudaGetLastError();
}

int bn_stat_mode() {
  static const int mode = (std::getenv("R3M_BN_FP32") && atoi(std::getenv("R3M_BN_FP32"))) ? 2 : 1;
  return mode;
}

DetScratch device_det_scratch() {
  static DetScratch d;
  if (!d.scratch) {
    float* buf = nullptr;
    const size_t bytes = kDetScratchFloats * sizeof(float) + kDetTickets * sizeof(int);
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return d;
    cudaMemset(buf, 0, bytes);
    d.scratch = buf;
    d.tickets = reinterpret_cast<int*>(buf + kDetScratchFloats);
  }
  return d;
}

cudaError_t launch_stem_bwd(const StemBwdArgs& a_in, cudaStream_t s) {
  StemBwdArgs a = a_in;
  if (a.C % 8 != 0 || a.C > 256 || kStemThreads % (a.C / 8) != 0 || (a.H & 1) || (a.W & 1))
    return cudaErrorInvalidValue;
  if (!a.det.scratch) a.det = device_det_scratch();
  if (!a.det.scratch) return cudaErrorMemoryAllocation;
  const int rows = a.N * (a.H / 2);  // one block iteration = one row of 2x2 quads
  if (a.ymax != nullptr && 256 % (a.C / 8) == 0) {
    const long long total = (long long)a.N * (a.H / 2) * (a.W / 2) * (a.C / 8);
    launch_kernel(stem_bwd_reduce_pooled_kernel,
                  std::min(kDetMaxBlocks, grid_for((total + 3) / 4, 256,
                                                   resident_blocks<stem_bwd_reduce_pooled_kernel>(256))),
                  256, 0, s, a);
  } else {
    launch_kernel(stem_bwd_kernel<false>,
                  std::min(kDetMaxBlocks, std::min(rows, resident_blocks<stem_bwd_kernel<false>>(kStemThreads))),
                  kStemThreads, 0, s, a);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  launch_kernel(stem_bwd_kernel<true>, std::min(rows, resident_blocks<stem_bwd_kernel<true>>(kStemThreads)),
                kStemThreads, 0, s, a);
  return cudaGetLastError();
}

cudaError_t launch_avgpool_fwd(const void* a, float* out, int N, int HW, int C, cudaStream_t s) {
  const int threads = std::min(1024, std::max(32, C / 2));
  launch_kernel(avgpool_fwd_kernel, N, threads, 0, s, reinterpret_cast<const bf16*>(a), out, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_avgpool_bwd(const float* dE, void* dA, int N, int HW, int C, cudaStream_t s) {
  const long long total = (long long)N * HW * (C / 8);
  launch_kernel(avgpool_bwd_kernel, grid_for(total, 256, 148 * 16), 256, 0, s, dE, reinterpret_cast<bf16*>(dA), N, HW, C);
  return cudaGetLastError();
}

namespace {
inline int mask_kind(const BnBwdArgs& a) { return a.a ? kMaskAct : (a.mask ? kMaskBits : kMaskNone); }
}  // namespace

cudaError_t launch_bn_bwd_reduce(const BnBwdArgs& a_in, cudaStream_t s) {
  BnBwdArgs a = a_in;
  static const int ab = std::getenv("R3M_AB") ? atoi(std::getenv("R3M_AB")) : 0;  // A/B aid: bit 1 = one row per trip
  a.single_rows = (ab & 2) ? 1 : 0;
  if (a.C % 8 != 0 || a.C > 2048 || 256 % (a.C / 8) != 0) return cudaErrorInvalidValue;
  if (!a.det.scratch) a.det = device_det_scratch();
  if (!a.det.scratch) return cudaErrorMemoryAllocation;
  if (a.C % kBwdSlice != 0) return cudaErrorInvalidValue;
  const int Cs = kBwdSlice, slices = a.C / Cs;
  const int rows_per_iter = 256 / (Cs / 8);
  // several rows per thread so that the block reduction and the ordered grid reduction are amortised; at most one
  // resident wave, at most kDetMaxBlocks blocks in all
  const bool dual = a.y2 != nullptr;
  const size_t smem = (size_t)rows_per_iter * (dual ? 3 : 2) * Cs * sizeof(float);  // = 16 / 24 KB for every C
  if ((reinterpret_cast<uintptr_t>(a.sums) & 15) != 0 || (dual && (reinterpret_cast<uintptr_t>(a.sums2) & 15) != 0))
    return cudaErrorInvalidValue;
#define R3M_LAUNCH(D, K)                                                                      \
  launch_kernel(bn_bwd_reduce_kernel<D, K>,                                                   \
                dim3(std::max(1, std::min(std::min(kDetMaxBlocks, resident_blocks<bn_bwd_reduce_kernel<D, K>>(256, smem)) / \
                                              slices,                                                       \
                                          grid_for((a.M + 15) / 16, rows_per_iter, 1 << 20))),                \
                     slices),                                                                               \
                256, smem, s, a)
  switch (mask_kind(a)) {
    case kMaskAct: if (dual) R3M_LAUNCH(true, kMaskAct); else R3M_LAUNCH(false, kMaskAct); break;
    case kMaskBits: if (dual) R3M_LAUNCH(true, kMaskBits); else R3M_LAUNCH(false, kMaskBits); break;
    default: if (dual) R3M_LAUNCH(true, kMaskNone); else R3M_LAUNCH(false, kMaskNone); break;
  }
#undef R3M_LAUNCH
  return cudaGetLastError();
}

bool bn_bwd_can_fuse(const BnBwdArgs& a) {
  // one BatchNorm, no residual share, raw accumulators (the engine's path), channel slicing of both passes compatible
  return a.y2 == nullptr && a.dz_out == nullptr && a.sums_raw && a.a == nullptr && a.C % kBwdSlice == 0 && a.C <= 2048 &&
         256 % (std::min(a.C, 256) / 8) == 0;
}

cudaError_t launch_bn_bwd_fused(const BnBwdArgs& a, int* counter, int* flag, cudaStream_t s) {
  if (!bn_bwd_can_fuse(a) || counter == nullptr || flag == nullptr) return cudaErrorInvalidValue;
  const int slices_r = a.C / kBwdSlice, slices_a = a.C >= 1024 ? a.C / 256 : 1;  // slices_a divides slices_r
  const int rows_r = 256 / (kBwdSlice / 8);
  const size_t smem = std::max((size_t)rows_r * 2 * kBwdSlice * sizeof(float), 3 * (size_t)(a.C / slices_a) * sizeof(float));
  int grid = 0;
#define R3M_LAUNCH(K)                                                                                   \
  do {                                                                                                  \
    grid = resident_blocks<bn_bwd_fused_kernel<K>>(256, smem);                                          \
    grid -= grid % slices_r;                                                                            \
    if (grid < slices_r) return cudaErrorInvalidValue;                                                  \
    launch_kernel_cooperative(bn_bwd_fused_kernel<K>, dim3(grid), 256, smem, s, a, slices_r, slices_a, counter, flag); \
  } while (0)
  if (a.mask)
    R3M_LAUNCH(kMaskBits);
  else
    R3M_LAUNCH(kMaskNone);
#undef R3M_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t s) {
  if (a.C % 8 != 0 || a.C > 2048 || 256 % (a.C / 8) != 0) return cudaErrorInvalidValue;
  const int slices = a.C >= 1024 ? a.C / 256 : 1, Cs = a.C / slices;
  const int rows_per_iter = 256 / (Cs / 8);
  const bool dual = a.y2 != nullptr, dz = a.dz_out != nullptr;
  const size_t coef_smem = (dual ? 6 : 3) * (size_t)Cs * sizeof(float);
#define R3M_LAUNCH1(D, K, Z)                                                                                          \
  launch_kernel(bn_bwd_apply_kernel<D, K, Z>,                                                                         \
                dim3(std::max(1, grid_for(a.M, rows_per_iter,                                                         \
                                          resident_blocks<bn_bwd_apply_kernel<D, K, Z>>(256, coef_smem) / slices)),   \
                     slices),                                                                                       \
                256, coef_smem, s, a)
#define R3M_LAUNCH(D, K)            \
  do {                              \
    if (dz) R3M_LAUNCH1(D, K, true); \
    else R3M_LAUNCH1(D, K, false);  \
  } while (0)
  switch (mask_kind(a)) {
    case kMaskAct: if (dual) R3M_LAUNCH(true, kMaskAct); else R3M_LAUNCH(false, kMaskAct); break;
    case kMaskBits: if (dual) R3M_LAUNCH(true, kMaskBits); else R3M_LAUNCH(false, kMaskBits); break;
    default: if (dual) R3M_LAUNCH(true, kMaskNone); else R3M_LAUNCH(false, kMaskNone); break;
  }
#undef R3M_LAUNCH
#undef R3M_LAUNCH1
  return cudaGetLastError();
}

cudaError_t launch_adam(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n, float lr, float beta1,
                        float beta2, float eps, int step, float grad_scale, cudaStream_t s) {
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  const uintptr_t align = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                          reinterpret_cast<uintptr_t>(v) | (reinterpret_cast<uintptr_t>(p_bf16) << 1);
  const size_t n4 = (align & 15) == 0 ? n / 4 : 0;  // vector body when every buffer allows 16-byte accesses
  if (n4 > 0)
    launch_kernel(adam_vec4_kernel, grid_for((long long)n4, 256, resident_blocks<adam_vec4_kernel>(256)), 256, 0, s,
                  reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
                  reinterpret_cast<float4*>(v), reinterpret_cast<uint2*>(p_bf16), n4, lr, beta1, beta2, eps, bc1,
                  sqrtf(bc2), grad_scale);
  if (n4 * 4 < n) {  // scalar tail (or everything, for unaligned buffers)
    const size_t o = n4 * 4;
    launch_kernel(adam_kernel, grid_for((long long)(n - o), 256, 148 * 16), 256, 0, s, p + o, g + o, m + o, v + o,
                  p_bf16 ? reinterpret_cast<bf16*>(p_bf16) + o : nullptr, n - o, lr, beta1, beta2, eps, bc1, sqrtf(bc2),
                  grad_scale);
  }
  return cudaGetLastError();
}

cudaError_t launch_pull_host(const void* host_mapped, void* dst, size_t bytes, cudaStream_t s) {
  if (bytes == 0 || (bytes & 15) != 0 ||
      ((reinterpret_cast<uintptr_t>(host_mapped) | reinterpret_cast<uintptr_t>(dst)) & 15) != 0)
    return cudaErrorInvalidValue;
  launch_kernel(pull_host_kernel, grid_for((long long)(bytes / 16), 256, 148), 256, 0, s,
                reinterpret_cast<const uint4*>(host_mapped), reinterpret_cast<uint4*>(dst), bytes / 16);
  return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t s) {
  launch_kernel(cast_bf16_kernel, grid_for((long long)n, 256, 148 * 16), 256, 0, s, src, reinterpret_cast<bf16*>(dst), n);
  return cudaGetLastError();
}

cudaError_t launch_pack_dgrad(const float* w, void* out, int Cout, int T, int Cin, int nt, const int* src_tap,
                              cudaStream_t s) {
  if (nt < 1 || nt > 16) return cudaErrorInvalidValue;
  TapList taps;
  for (int i = 0; i < 16; ++i) taps.t[i] = i < nt ? src_tap[i] : 0;
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, nt);
  launch_kernel(pack_dgrad_kernel, grid, dim3(32, 8), 0, s, w, reinterpret_cast<bf16*>(out), Cout, T, Cin, nt, taps);
  return cudaGetLastError();
}

cudaError_t launch_pack_dgrad_multi(const PackDgradEntry* table_dev, int entries, int total_blocks, cudaStream_t s) {
  if (entries < 1 || total_blocks < 1) return cudaErrorInvalidValue;
  launch_kernel(pack_dgrad_multi_kernel, total_blocks, dim3(32, 8), 0, s, table_dev, entries);
  return cudaGetLastError();
}

cudaError_t launch_stem_pack(const float* w_oihw, void* wp_bf16, cudaStream_t s) {
  launch_kernel(stem_pack_kernel, 64, 256, 0, s, w_oihw, reinterpret_cast<bf16*>(wp_bf16));
  return cudaGetLastError();
}

cudaError_t launch_stem_unpack_grad(const float* dwp, float* dw_oihw, cudaStream_t s) {
  launch_kernel(stem_unpack_grad_kernel, 64, 256, 0, s, dwp, dw_oihw);
  return cudaGetLastError();
}

cudaError_t launch_round_tf32(const float* src, float* dst, size_t n, cudaStream_t s) {
  launch_kernel(round_tf32_kernel, grid_for((long long)n, 256, 148 * 16), 256, 0, s, src, dst, n);
  return cudaGetLastError();
}

cudaError_t launch_preprocess_stem_f32(const void* obs, int format, float* xs, int N, cudaStream_t s) {
  const long long total = (long long)N * 112 * 112 * 4;
  const int grid = grid_for(total, 256, 148 * 16);
  switch (format) {
    case kObsF32NCHW: launch_kernel(preprocess_stem_f32_kernel<kObsF32NCHW>, grid, 256, 0, s, obs, xs, N); break;
    case kObsU8NCHW: launch_kernel(preprocess_stem_f32_kernel<kObsU8NCHW>, grid, 256, 0, s, obs, xs, N); break;
    case kObsU8NHWC: launch_kernel(preprocess_stem_f32_kernel<kObsU8NHWC>, grid, 256, 0, s, obs, xs, N); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_maxpool_f32(const float* y, float* a, int N, int H, int W, int C, cudaStream_t s) {
  if (C % 4 != 0 || (H & 1) || (W & 1)) return cudaErrorInvalidValue;
  launch_kernel(maxpool_f32_kernel, grid_for((long long)N * (H / 2) * (W / 2) * (C / 4), 256, 148 * 16), 256, 0, s, y, a,
                N, H, W, C);
  return cudaGetLastError();
}

cudaError_t launch_avgpool_fwd_f32(const float* a, float* out, int N, int HW, int C, cudaStream_t s) {
  launch_kernel(avgpool_fwd_f32_kernel, N, std::min(1024, std::max(32, C)), 0, s, a, out, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_stem_pack_f32(const float* w_oihw, float* wp, cudaStream_t s) {
  launch_kernel(stem_pack_f32_kernel, 64, 256, 0, s, w_oihw, wp);
  return cudaGetLastError();
}

namespace {
__global__ void __launch_bounds__(256) ordered_sum_kernel(const float* __restrict__ x, size_t n, unsigned long long* acc,
                                                          int* ticket, float* __restrict__ out) {
  pdl_sync();
  __shared__ int s_last;
  // one fx_add per element on purpose: the test wants to see that ANY grouping / order gives the same bits
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) fx_add(acc, x[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    out[0] = fx_to_float(acc);
    fx_clear(acc);
    *ticket = 0;
  }
}
// the BatchNorm statistics law on a flat array: fixed-point sum and sum of squares, then bn_batch_moments
__global__ void __launch_bounds__(256) ordered_moments_kernel(const float* __restrict__ x, int n, unsigned long long* acc,
                                                              int* ticket, float* __restrict__ out) {
  pdl_sync();
  __shared__ int s_last;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = x[i];
    fx_add(acc, v);
    fx_add(acc + kFxWords, v * v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    bn_batch_moments(reinterpret_cast<const float*>(acc), nullptr, 1, 0, n, out[0], out[1]);
    fx_clear(acc);
    fx_clear(acc + kFxWords);
    *ticket = 0;
  }
}
}  // namespace

cudaError_t launch_ordered_moments(const float* x, int n, float* out, int blocks, cudaStream_t s) {
  DetScratch d = device_det_scratch();
  if (!d.scratch) return cudaErrorMemoryAllocation;
  launch_kernel(ordered_moments_kernel, std::max(1, std::min(blocks, 1024)), 256, 0, s, x, n,
                reinterpret_cast<unsigned long long*>(d.scratch), d.tickets, out);
  return cudaGetLastError();
}

cudaError_t launch_ordered_sum(const float* x, size_t n, float* out, int blocks, cudaStream_t s) {
  DetScratch d = device_det_scratch();
  if (!d.scratch) return cudaErrorMemoryAllocation;
  launch_kernel(ordered_sum_kernel, std::max(1, std::min(blocks, 1024)), 256, 0, s, x, n,
                reinterpret_cast<unsigned long long*>(d.scratch), d.tickets, out);
  return cudaGetLastError();
}

}  // namespace r3m
