// HBM-bound kernels of the R3M pretraining step: input normalisation + stem re-layout, BatchNorm apply (train / eval)
// with fused residual + ReLU, stem BN+ReLU+maxpool, global average pool, BatchNorm backward, pooling backward,
// fused Adam and the filter re-packers.  All activations are NHWC bf16, 16-byte vector accesses (8 channels per
// thread), fp32 arithmetic.  Reference semantics: torchvision resnet.py (BasicBlock.forward :89-105,
// Bottleneck.forward :143-163, _forward_impl :266-282) and torch.nn.BatchNorm2d / MaxPool2d / Adam defaults.
#include "elementwise.cuh"

#include <cuda_bf16.h>

#include <algorithm>

#include "ptx.cuh"

namespace r3m {

namespace {

typedef __nv_bfloat16 bf16;

struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const bf16* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));  // every ld8 source is read-only within its kernel
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
__device__ __forceinline__ void st8(bf16* p, const F8& f) {
  uint4 u;
  u.x = pack_bf16x2(f.v[0], f.v[1]);
  u.y = pack_bf16x2(f.v[2], f.v[3]);
  u.z = pack_bf16x2(f.v[4], f.v[5]);
  u.w = pack_bf16x2(f.v[6], f.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ void bn_coeffs(int train, const float* sum, const float* sq, float inv_m, const float* gamma,
                                          const float* beta, const float* rm, const float* rv, int c, float& scale,
                                          float& shift, float& mean, float& var) {
  if (train) {
    mean = sum[c] * inv_m;
    var = fmaxf(sq[c] * inv_m - mean * mean, 0.f);
  } else {
    mean = rm[c];
    var = rv[c];
  }
  const float rstd = 1.0f / sqrtf(var + kBnEps);
  scale = gamma[c] * rstd;
  shift = beta[c] - mean * scale;
}

// block 0 publishes the batch statistics for backward and folds them into the running estimates
__device__ __forceinline__ void bn_publish(int C, int M, const float* sum, const float* sq, float* save_mean,
                                           float* save_rstd, float* rm, float* rv, int update_running) {
  const float inv_m = 1.0f / (float)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = sum[c] * inv_m;
    const float var = fmaxf(sq[c] * inv_m - mean * mean, 0.f);
    if (save_mean) save_mean[c] = mean;
    if (save_rstd) save_rstd[c] = 1.0f / sqrtf(var + kBnEps);
    if (update_running && rm && rv) {
      const float unbiased = (M > 1) ? var * ((float)M / (float)(M - 1)) : var;
      rm[c] = (1.f - kBnMomentum) * rm[c] + kBnMomentum * mean;
      rv[c] = (1.f - kBnMomentum) * rv[c] + kBnMomentum * unbiased;
    }
  }
}

// ------------------------------------------------------------------------------------------------ preprocess
__global__ void __launch_bounds__(256) preprocess_stem_kernel(const float* __restrict__ obs, bf16* __restrict__ xs,
                                                              int N) {
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
    const int col = 2 * (q - 2 + kw);  // even, so a float2 covers dx = 0, 1
    uint32_t w[8];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      float v[2][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float2 px = make_float2(0.f, 0.f);
        const bool inside = (col >= 0 && col < 224);
        if (inside) {
          px = *reinterpret_cast<const float2*>(obs + (((long long)n * 3 + c) * 224 + (2 * i + dy)) * 224 + col);
          px.x = (px.x * (1.f / 255.f) - mean[c]) * istd[c];
          px.y = (px.y * (1.f / 255.f) - mean[c]) * istd[c];
        }
        v[0][c] = px.x;
        v[1][c] = px.y;
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

// ------------------------------------------------------------------------------------------------ BN apply
template <bool kDual, bool kRes>
__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyArgs a) {
  const int C8 = a.C >> 3;
  const int chunk = threadIdx.x % C8;
  const int rows_per_iter = blockDim.x / C8;
  const int r0 = threadIdx.x / C8;
  const float inv_m = 1.0f / (float)a.M;
  constexpr bool dual = kDual;
  float sc[8], sh[8], sc2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float mean, var;
    bn_coeffs(a.train, a.sum, a.sq, inv_m, a.gamma, a.beta, a.running_mean, a.running_var, chunk * 8 + j, sc[j], sh[j],
              mean, var);
    sc2[j] = 0.f;
    if (dual) {
      float shift2;
      bn_coeffs(a.train, a.sum2, a.sq2, inv_m, a.gamma2, a.beta2, a.running_mean2, a.running_var2, chunk * 8 + j,
                sc2[j], shift2, mean, var);
      sh[j] += shift2;
    }
  }
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  const bf16* __restrict__ y2 = reinterpret_cast<const bf16*>(a.y2);
  const bf16* __restrict__ res = reinterpret_cast<const bf16*>(a.residual);
  bf16* __restrict__ out = reinterpret_cast<bf16*>(a.a);
#pragma unroll 2
  for (long long row = (long long)blockIdx.x * rows_per_iter + r0; row < a.M;
       row += (long long)gridDim.x * rows_per_iter) {
    const long long off = row * a.C + chunk * 8;
    // all loads of the row first (no control flow between them)
    F8 f = ld8(y + off);
    F8 t, r;
    if (kDual) t = ld8(y2 + off);
    if (kRes) r = ld8(res + off);
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
      a.mask_out[row * C8 + chunk] = (uint8_t)bits;
    }
  }
  if (a.train && blockIdx.x == 0) {
    // every block has already read sum/sq into registers for its own coefficients; running stats are separate
    // buffers, so the in-place update below cannot race with other blocks
    bn_publish(a.C, a.M, a.sum, a.sq, a.save_mean, a.save_rstd, a.running_mean, a.running_var, a.update_running);
    if (dual)
      bn_publish(a.C, a.M, a.sum2, a.sq2, a.save_mean2, a.save_rstd2, a.running_mean2, a.running_var2,
                 a.update_running);
  }
}

__global__ void bn_fold_kernel(const BnFoldEntry* __restrict__ table) {
  const BnFoldEntry e = table[blockIdx.x];
  for (int c = threadIdx.x; c < e.C; c += blockDim.x) {
    const float sc = e.gamma[c] / sqrtf(e.running_var[c] + kBnEps);
    e.scale[c] = sc;
    e.shift[c] = e.beta[c] - e.running_mean[c] * sc;
  }
}

// ------------------------------------------------------------------------------------------------ stem pool
__global__ void __launch_bounds__(256) stem_pool_kernel(const StemPoolArgs a) {
  const int C8 = a.C >> 3;
  const int P = a.H / 2, Q = a.W / 2;
  const long long total = (long long)a.N * P * Q * C8;
  const float inv_m = 1.0f / ((float)a.N * a.H * a.W);
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  bf16* __restrict__ out = reinterpret_cast<bf16*>(a.a);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(idx % C8);
    long long t = idx / C8;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P);
    const int n = (int)(t / P);
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float mean, var;
      bn_coeffs(a.train, a.sum, a.sq, inv_m, a.gamma, a.beta, a.running_mean, a.running_var, chunk * 8 + j, sc[j],
                sh[j], mean, var);
    }
    float best[8];
    int code[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -1.f;
      code[j] = 0;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * p - 1 + r;
      if (h < 0 || h >= a.H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * q - 1 + s;
        if (w < 0 || w >= a.W) continue;
        const F8 f = ld8(y + (((long long)n * a.H + h) * a.W + w) * a.C + chunk * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = fmaxf(fmaf(f.v[j], sc[j], sh[j]), 0.f);
          if (v > best[j]) {  // strict: the first maximum in scan order wins, as in ATen's max_pool2d
            best[j] = v;
            code[j] = r * 3 + s;
          }
        }
      }
    }
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = best[j];
    st8(out + idx * 8, o);
    if (a.argmax) {
      uint2 packed;
      packed.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
      packed.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
      *reinterpret_cast<uint2*>(a.argmax + idx * 8) = packed;
    }
  }
  if (a.train && blockIdx.x == 0) {
    bn_publish(a.C, a.N * a.H * a.W, a.sum, a.sq, a.save_mean, a.save_rstd, a.running_mean, a.running_var,
               a.update_running);
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const bf16* __restrict__ dA, const bf16* __restrict__ a,
                                                          const uint8_t* __restrict__ argmax, bf16* __restrict__ dz,
                                                          int N, int H, int W, int C) {
  const int C8 = C >> 3;
  const int P = H / 2, Q = W / 2;
  const long long total = (long long)N * H * W * C8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(idx % C8);
    long long t = idx / C8;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    F8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    // windows p with 2p-1 <= h <= 2p+1
    const int p_lo = h >> 1, p_hi = (h + 1) >> 1;
    const int q_lo = w >> 1, q_hi = (w + 1) >> 1;
    for (int p = p_lo; p <= p_hi; ++p) {
      if (p >= P) continue;
      const int r = h - (2 * p - 1);
      for (int q = q_lo; q <= q_hi; ++q) {
        if (q >= Q) continue;
        const int s = w - (2 * q - 1);
        const int my = r * 3 + s;
        const long long o = (((long long)n * P + p) * Q + q) * C + chunk * 8;
        const uint2 codes = *reinterpret_cast<const uint2*>(argmax + o);
        const F8 g = ld8(dA + o);
        const F8 act = ld8(a + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int cj = (j < 4 ? (codes.x >> (8 * j)) : (codes.y >> (8 * (j - 4)))) & 0xFF;
          if (cj == my && act.v[j] > 0.f) acc.v[j] += g.v[j];
        }
      }
    }
    st8(dz + idx * 8, acc);
  }
}

// ------------------------------------------------------------------------------------------------ avg pool
__global__ void avgpool_fwd_kernel(const bf16* __restrict__ a, float* __restrict__ out, int HW, int C) {
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

template <bool kDual, int kMask>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  extern __shared__ float s_red[];  // [3][C]: sum(dz), sum(dz*xhat), sum(dz*xhat2)
  const int C8 = a.C >> 3;
  const int chunk = threadIdx.x % C8;
  const int rows_per_iter = blockDim.x / C8;
  const int r0 = threadIdx.x / C8;
  for (int i = threadIdx.x; i < 3 * a.C; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  float mean[8], rstd[8], mean2[8], rstd2[8], s1[8], s2[8], s3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = a.mean[chunk * 8 + j];
    rstd[j] = a.rstd[chunk * 8 + j];
    mean2[j] = kDual ? a.mean2[chunk * 8 + j] : 0.f;
    rstd2[j] = kDual ? a.rstd2[chunk * 8 + j] : 0.f;
    s1[j] = 0.f;
    s2[j] = 0.f;
    s3[j] = 0.f;
  }
  const bf16* __restrict__ dA = reinterpret_cast<const bf16*>(a.dA);
  const bf16* __restrict__ act = reinterpret_cast<const bf16*>(a.a);
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  const bf16* __restrict__ y2 = reinterpret_cast<const bf16*>(a.y2);
  const uint8_t* __restrict__ mask = a.mask;
  for (long long row = (long long)blockIdx.x * rows_per_iter + r0; row < a.M;
       row += (long long)gridDim.x * rows_per_iter) {
    const long long off = row * a.C + chunk * 8;
    // all loads of the row first
    F8 g = ld8(dA + off);
    const F8 yy = ld8(y + off);
    F8 m, t;
    unsigned bits = 0;
    if (kMask == kMaskAct) m = ld8(act + off);
    if (kMask == kMaskBits) bits = __ldg(mask + row * C8 + chunk);
    if (kDual) t = ld8(y2 + off);
    mask_gradient<kMask>(g, m, bits);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] += g.v[j];
      s2[j] = fmaf(g.v[j], (yy.v[j] - mean[j]) * rstd[j], s2[j]);
      if (kDual) s3[j] = fmaf(g.v[j], (t.v[j] - mean2[j]) * rstd2[j], s3[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_red[chunk * 8 + j], s1[j]);
    atomicAdd(&s_red[a.C + chunk * 8 + j], s2[j]);
    if (kDual) atomicAdd(&s_red[2 * a.C + chunk * 8 + j], s3[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * a.C; i += blockDim.x) atomicAdd(&a.sums[i], s_red[i]);
  if (kDual)
    for (int i = threadIdx.x; i < a.C; i += blockDim.x) atomicAdd(&a.sums2[i], s_red[2 * a.C + i]);
}

template <bool kDual, int kMask, bool kDz>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs a) {
  const int C8 = a.C >> 3;
  const int chunk = threadIdx.x % C8;
  const int rows_per_iter = blockDim.x / C8;
  const int r0 = threadIdx.x / C8;
  const float inv_m = 1.0f / (float)a.M;
  float mean[8], rstd[8], grs[8], mdz[8], mdzx[8], mean2[8], rstd2[8], grs2[8], mdzx2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = chunk * 8 + j;
    mean[j] = a.mean[c];
    rstd[j] = a.rstd[c];
    grs[j] = a.gamma[c] * rstd[j];
    mdz[j] = a.sums[c] * inv_m;
    mdzx[j] = a.sums[a.C + c] * inv_m;
    mean2[j] = kDual ? a.mean2[c] : 0.f;
    rstd2[j] = kDual ? a.rstd2[c] : 0.f;
    grs2[j] = kDual ? a.gamma2[c] * rstd2[j] : 0.f;
    mdzx2[j] = kDual ? a.sums2[c] * inv_m : 0.f;
  }
  const bf16* __restrict__ dA = reinterpret_cast<const bf16*>(a.dA);
  const bf16* __restrict__ act = reinterpret_cast<const bf16*>(a.a);
  const bf16* __restrict__ y = reinterpret_cast<const bf16*>(a.y);
  const bf16* __restrict__ y2 = reinterpret_cast<const bf16*>(a.y2);
  const uint8_t* __restrict__ mask = a.mask;
  bf16* __restrict__ dy = reinterpret_cast<bf16*>(a.dy);
  bf16* __restrict__ dy2 = reinterpret_cast<bf16*>(a.dy2);
  bf16* __restrict__ dzo = reinterpret_cast<bf16*>(a.dz_out);
  for (long long row = (long long)blockIdx.x * rows_per_iter + r0; row < a.M;
       row += (long long)gridDim.x * rows_per_iter) {
    const long long off = row * a.C + chunk * 8;
    F8 g = ld8(dA + off);
    const F8 yy = ld8(y + off);
    F8 m, t;
    unsigned bits = 0;
    if (kMask == kMaskAct) m = ld8(act + off);
    if (kMask == kMaskBits) bits = __ldg(mask + row * C8 + chunk);
    if (kDual) t = ld8(y2 + off);
    mask_gradient<kMask>(g, m, bits);
    if (kDz) st8(dzo + off, g);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xhat = (yy.v[j] - mean[j]) * rstd[j];
      o.v[j] = grs[j] * (g.v[j] - mdz[j] - xhat * mdzx[j]);
    }
    st8(dy + off, o);
    if (kDual) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xhat = (t.v[j] - mean2[j]) * rstd2[j];
        o.v[j] = grs2[j] * (g.v[j] - mdz[j] - xhat * mdzx2[j]);
      }
      st8(dy2 + off, o);
    }
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      if (a.dbeta) a.dbeta[c] = a.sums[c];
      if (a.dgamma) a.dgamma[c] = a.sums[a.C + c];
      if (kDual) {
        if (a.dbeta2) a.dbeta2[c] = a.sums[c];
        if (a.dgamma2) a.dgamma2[c] = a.sums2[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ Adam / casts
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   bf16* __restrict__ pb, size_t n, float lr, float beta1, float beta2,
                                                   float eps, float bc1, float bc2_sqrt, float grad_scale) {
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

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                        size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// ------------------------------------------------------------------------------------------------ filter packers
struct TapList {
  int t[16];
};
__global__ void pack_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int T, int Cin, int nt,
                                  TapList taps) {
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

__device__ __forceinline__ bool stem_map(int k, int rr, int j, int& oihw) {
  const int kw = j >> 4, sub = (j >> 2) & 3, c = j & 3;
  const int dy = sub >> 1, dx = sub & 1;
  const int r = 2 * rr + dy - 1, s = 2 * kw + dx - 1;
  if (c >= 3 || r < 0 || r >= 7 || s < 0 || s >= 7) return false;
  oihw = ((k * 3 + c) * 7 + r) * 7 + s;
  return true;
}
__global__ void stem_pack_kernel(const float* __restrict__ w, bf16* __restrict__ wp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 4 * 64) return;
  const int j = idx & 63, rr = (idx >> 6) & 3, k = idx >> 8;
  int o;
  wp[idx] = __float2bfloat16_rn(stem_map(k, rr, j, o) ? w[o] : 0.f);
}
__global__ void stem_unpack_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw) {
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

}  // namespace

cudaError_t launch_preprocess_stem(const float* obs, void* xs, int N, cudaStream_t s) {
  const long long total = (long long)N * 112 * 112 * 4;
  preprocess_stem_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, s>>>(obs, reinterpret_cast<bf16*>(xs), N);
  return cudaGetLastError();
}

cudaError_t launch_bn_fold(const BnFoldEntry* table_dev, int entries, cudaStream_t s) {
  bn_fold_kernel<<<entries, 256, 0, s>>>(table_dev);
  return cudaGetLastError();
}

cudaError_t launch_bn_apply(const BnApplyArgs& a, cudaStream_t s) {
  if (a.C % 8 != 0 || a.C > 2048 || 256 % (a.C / 8) != 0) return cudaErrorInvalidValue;
  const int rows_per_iter = 256 / (a.C / 8);
  const int blocks = grid_for(a.M, rows_per_iter, 148 * 8);
  if (a.y2 && a.residual) return cudaErrorInvalidValue;  // a block tail has either an identity or a downsample branch
  if (a.y2)
    bn_apply_kernel<true, false><<<blocks, 256, 0, s>>>(a);
  else if (a.residual)
    bn_apply_kernel<false, true><<<blocks, 256, 0, s>>>(a);
  else
    bn_apply_kernel<false, false><<<blocks, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_stem_bn_relu_maxpool(const StemPoolArgs& a, cudaStream_t s) {
  if (a.C % 8 != 0 || (a.H & 1) || (a.W & 1)) return cudaErrorInvalidValue;
  const long long total = (long long)a.N * (a.H / 2) * (a.W / 2) * (a.C / 8);
  stem_pool_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_maxpool_bwd(const void* dA, const void* a, const uint8_t* argmax, void* dz, int N, int H, int W,
                               int C, cudaStream_t s) {
  const long long total = (long long)N * H * W * (C / 8);
  maxpool_bwd_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, s>>>(
      reinterpret_cast<const bf16*>(dA), reinterpret_cast<const bf16*>(a), argmax, reinterpret_cast<bf16*>(dz), N, H, W,
      C);
  return cudaGetLastError();
}

cudaError_t launch_avgpool_fwd(const void* a, float* out, int N, int HW, int C, cudaStream_t s) {
  const int threads = std::min(1024, std::max(32, C / 2));
  avgpool_fwd_kernel<<<N, threads, 0, s>>>(reinterpret_cast<const bf16*>(a), out, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_avgpool_bwd(const float* dE, void* dA, int N, int HW, int C, cudaStream_t s) {
  const long long total = (long long)N * HW * (C / 8);
  avgpool_bwd_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, s>>>(dE, reinterpret_cast<bf16*>(dA), N, HW, C);
  return cudaGetLastError();
}

namespace {
inline int mask_kind(const BnBwdArgs& a) { return a.a ? kMaskAct : (a.mask ? kMaskBits : kMaskNone); }
}  // namespace

cudaError_t launch_bn_bwd_reduce(const BnBwdArgs& a, cudaStream_t s) {
  if (a.C % 8 != 0 || a.C > 2048 || 256 % (a.C / 8) != 0) return cudaErrorInvalidValue;
  const int rows_per_iter = 256 / (a.C / 8);
  // several rows per thread so that the shared/global atomics are amortised
  const int blocks = grid_for((a.M + 15) / 16, rows_per_iter, 148 * 4);
  const size_t smem = 3 * a.C * sizeof(float);
  const bool dual = a.y2 != nullptr;
#define R3M_LAUNCH(D, K) bn_bwd_reduce_kernel<D, K><<<blocks, 256, smem, s>>>(a)
  switch (mask_kind(a)) {
    case kMaskAct: if (dual) R3M_LAUNCH(true, kMaskAct); else R3M_LAUNCH(false, kMaskAct); break;
    case kMaskBits: if (dual) R3M_LAUNCH(true, kMaskBits); else R3M_LAUNCH(false, kMaskBits); break;
    default: if (dual) R3M_LAUNCH(true, kMaskNone); else R3M_LAUNCH(false, kMaskNone); break;
  }
#undef R3M_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t s) {
  if (a.C % 8 != 0 || a.C > 2048 || 256 % (a.C / 8) != 0) return cudaErrorInvalidValue;
  const int rows_per_iter = 256 / (a.C / 8);
  const int blocks = grid_for(a.M, rows_per_iter, 148 * 8);
  const bool dual = a.y2 != nullptr, dz = a.dz_out != nullptr;
#define R3M_LAUNCH(D, K)                                                  \
  do {                                                                    \
    if (dz) bn_bwd_apply_kernel<D, K, true><<<blocks, 256, 0, s>>>(a);    \
    else bn_bwd_apply_kernel<D, K, false><<<blocks, 256, 0, s>>>(a);      \
  } while (0)
  switch (mask_kind(a)) {
    case kMaskAct: if (dual) R3M_LAUNCH(true, kMaskAct); else R3M_LAUNCH(false, kMaskAct); break;
    case kMaskBits: if (dual) R3M_LAUNCH(true, kMaskBits); else R3M_LAUNCH(false, kMaskBits); break;
    default: if (dual) R3M_LAUNCH(true, kMaskNone); else R3M_LAUNCH(false, kMaskNone); break;
  }
#undef R3M_LAUNCH
  return cudaGetLastError();
}

cudaError_t launch_adam(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n, float lr, float beta1,
                        float beta2, float eps, int step, float grad_scale, cudaStream_t s) {
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<grid_for((long long)n, 256, 148 * 16), 256, 0, s>>>(p, g, m, v, reinterpret_cast<bf16*>(p_bf16), n, lr,
                                                                    beta1, beta2, eps, bc1, sqrtf(bc2), grad_scale);
  return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t s) {
  cast_bf16_kernel<<<grid_for((long long)n, 256, 148 * 16), 256, 0, s>>>(src, reinterpret_cast<bf16*>(dst), n);
  return cudaGetLastError();
}

cudaError_t launch_pack_dgrad(const float* w, void* out, int Cout, int T, int Cin, int nt, const int* src_tap,
                              cudaStream_t s) {
  if (nt < 1 || nt > 16) return cudaErrorInvalidValue;
  TapList taps;
  for (int i = 0; i < 16; ++i) taps.t[i] = i < nt ? src_tap[i] : 0;
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, nt);
  pack_dgrad_kernel<<<grid, dim3(32, 8), 0, s>>>(w, reinterpret_cast<bf16*>(out), Cout, T, Cin, nt, taps);
  return cudaGetLastError();
}

cudaError_t launch_stem_pack(const float* w_oihw, void* wp_bf16, cudaStream_t s) {
  stem_pack_kernel<<<64, 256, 0, s>>>(w_oihw, reinterpret_cast<bf16*>(wp_bf16));
  return cudaGetLastError();
}

cudaError_t launch_stem_unpack_grad(const float* dwp, float* dw_oihw, cudaStream_t s) {
  stem_unpack_grad_kernel<<<64, 256, 0, s>>>(dwp, dw_oihw);
  return cudaGetLastError();
}

}  // namespace r3m
