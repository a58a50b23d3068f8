#include "lang.cuh"

#include <algorithm>

#include "loss.cuh"
#include "launch.h"
#include "ptx.cuh"

namespace r3m {

namespace {

// (e0, e_t) row pair of evaluation j for clip b — r3m/trainer.py:72-92:
//   j = 0..2  positives        G(e0, {eg, es1, es2})
//   j = 3..5  in-clip negatives G(e0, {e0, es0, es1})
//   j = 6..14 shuffled negatives, iteration i = (j-6)/3, target t = (j-6)%3, permutation perms[3i+t]:
//             G(e0[perm], {eg, es1, es2}[perm]) — the sentence embedding is NOT permuted.
__device__ __forceinline__ void eval_rows(int j, int b, const int* __restrict__ perms, int B, int& r0, int& r1) {
  if (j < 3) {
    r0 = 5 * b;
    r1 = 5 * b + (j == 0 ? 1 : (j == 1 ? 3 : 4));
  } else if (j < 6) {
    r0 = 5 * b;
    r1 = 5 * b + (j == 3 ? 0 : (j == 4 ? 2 : 3));
  } else {
    const int q = j - 6, t = q % 3;
    const int pb = perms[q * B + b];
    r0 = 5 * pb;
    r1 = 5 * pb + (t == 0 ? 1 : (t == 1 ? 3 : 4));
  }
}

// H1[row] = relu(U[clip of e0] + V[e_t row] + Lc[b] + b1)
__global__ void __launch_bounds__(256) lang_layer1_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                          const float* __restrict__ Lc, const float* __restrict__ b1,
                                                          const int* __restrict__ perms, float* __restrict__ H1,
                                                          LangDims d) {
  pdl_sync();
  const int row = blockIdx.x;
  const int j = row / d.B, b = row - j * d.B;
  int r0, r1;
  eval_rows(j, b, perms, d.B, r0, r1);
  const float4* u = reinterpret_cast<const float4*>(U + (size_t)(r0 / 5) * d.H);
  const float4* v = reinterpret_cast<const float4*>(V + (size_t)r1 * d.H);
  const float4* l = reinterpret_cast<const float4*>(Lc + (size_t)b * d.H);
  const float4* bias = reinterpret_cast<const float4*>(b1);
  float4* out = reinterpret_cast<float4*>(H1 + (size_t)row * d.H);
  for (int i = threadIdx.x; i < d.H / 4; i += blockDim.x) {
    const float4 a = u[i], c = v[i], e = l[i], f = bias[i];
    out[i] = make_float4(fmaxf(a.x + c.x + e.x + f.x, 0.f), fmaxf(a.y + c.y + e.y + f.y, 0.f),
                         fmaxf(a.z + c.z + e.z + f.z, 0.f), fmaxf(a.w + c.w + e.w + f.w, 0.f));
  }
}

// Backward of the gather-add of layer 1, itself as a GATHER (no atomics: deterministic): one block per target row,
//   block < B        : dU[c]      = sum of dpre1 rows whose e0 comes from clip c
//   block < 6B       : dV[r]      = sum of dpre1 rows whose e_t is embedding row r = block - B
//   else             : dLc[b]     = sum over the 15 evaluations of clip b
// Evaluations 0..5 are in-clip (the source row is known); for the shuffled negatives 6..14 the block marks, in shared
// memory, every clip b whose permutation points at the target clip (a scan, so any index map works, not only
// bijections) and every thread then adds the marked rows in (evaluation, clip) order.
__global__ void __launch_bounds__(256) lang_layer1_bwd_kernel(const float* __restrict__ dpre, const int* __restrict__ perms,
                                                              float* __restrict__ dU, float* __restrict__ dV,
                                                              float* __restrict__ dLc, LangDims d) {
  pdl_sync();
  extern __shared__ unsigned char s_hit[];  // [9][B]
  __shared__ int s_src[64];                  // compacted source rows of the shuffled negatives (in (q, b) order)
  __shared__ int s_n;
  const int B = d.B;
  const int blk = blockIdx.x;
  const int kind = blk < B ? 0 : (blk < 6 * B ? 1 : 2);
  const int target = kind == 0 ? blk : (kind == 1 ? blk - B : blk - 6 * B);
  const int c = kind == 1 ? target / 5 : target;  // clip of the target row
  const int f = kind == 1 ? target - 5 * c : 0;   // frame of a dV row
  float* out = (kind == 0 ? dU : (kind == 1 ? dV : dLc)) + (size_t)target * d.H;
  for (int idx = threadIdx.x; idx < 9 * B; idx += blockDim.x) {
    const int q = idx / B, t = q % 3;
    const int ft = t == 0 ? 1 : (t == 1 ? 3 : 4);  // e_t frame of shuffled negative q
    s_hit[idx] = (kind != 2 && perms[idx] == c && (kind == 0 || ft == f)) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    // warp 0 compacts the marks in (q, b) order (ballot + prefix popcount: deterministic); a permutation marks exactly
    // one clip per negative, so the list has <= 9 entries — the rare overflow of a non-bijective map stays in s_hit
    int n = 0;
    for (int base = 0; base < 9 * B; base += 32) {
      const int idx = base + threadIdx.x;
      const bool hit = idx < 9 * B && s_hit[idx];
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = n + __popc(m & ((1u << threadIdx.x) - 1u));
        if (pos < 64) {
          s_src[pos] = (6 * B) + idx;  // row (6 + q) * B + b of dpre
          s_hit[idx] = 0;
        }
      }
      n += __popc(m);
    }
    if (threadIdx.x == 0) s_n = n;
  }
  __syncthreads();
  const int n_list = min(s_n, 64);
  const bool overflow = s_n > 64;
  for (int i = threadIdx.x; i < d.H; i += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < 6; ++j) {
      bool hit = true;
      if (kind == 1) {
        const int fj = j < 3 ? (j == 0 ? 1 : (j == 1 ? 3 : 4)) : (j == 3 ? 0 : (j == 4 ? 2 : 3));
        hit = (fj == f);
      }
      if (hit) acc += dpre[(size_t)(j * B + c) * d.H + i];
    }
    if (kind == 2) {
      for (int j = 6; j < 15; ++j) acc += dpre[(size_t)(j * B + c) * d.H + i];
    } else {
      for (int k = 0; k < n_list; ++k) acc += dpre[(size_t)s_src[k] * d.H + i];
      if (overflow)
        for (int idx = 0; idx < 9 * B; ++idx)
          if (s_hit[idx]) acc += dpre[(size_t)(6 * B + idx) * d.H + i];
    }
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 SIMT GEMM (exact fp32: the reference's nn.Linear runs in true fp32, torch.backends.cuda.matmul.allow_tf32 is
// False).  C[M,N] = A . B with selectable operand storage:
//   kAK: A stored [M][K] (K contiguous) else [K][M];   kBK: B stored [N][K] (K contiguous) else [K][N].
// Epilogue: + bias[col], ReLU, or gating by mask[row][col] > 0 (ReLU backward), optional C += result.
// Tile 64 x 128 x 16, 128 threads, 8 x 8 outputs per thread (two 4-row groups x two 4-column groups, so that every
// fragment read is one LDS.128 and the 64 FMAs per k step are fed by four of them), shared-memory double buffering with
// the next tile's global loads held in registers.  The head's GEMMs are 960 x 1024 x 1024: 120 CTAs, one wave.
// Vector loads need the contiguous dimension of each operand to be a multiple of 4 floats (checked by the launcher).
// ---------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int lda, ldb, ldc;
  const float* bias;
  int relu;
  const float* mask;
  int ldm;
  int accumulate;  // C += result
  int kchunk = 0;  // split-K (gridDim.z > 1): K range per z slice, a multiple of kTK; slice z writes its partial product
                   // to splitk[z][M][N] and splitk_reduce_kernel adds the slices in order (deterministic)
  float* splitk = nullptr;
};

constexpr int kTM = 64, kTN = 128, kTK = 16;
constexpr size_t kSplitKFloats = (size_t)4 << 20;  // split-K scratch of the head's few-tile GEMMs (16 MB)

// one operand tile [kTK][ROWS] (k-major in shared memory) <- global, through registers
template <int ROWS, bool kKContig>
struct TileLoader {
  static constexpr int kVecs = ROWS * kTK / 4 / 128;  // float4 per thread
  float4 v[kVecs];
  __device__ __forceinline__ void load(const float* __restrict__ src, int ld, int row0, int rows, int k0, int K) {
#pragma unroll
    for (int i = 0; i < kVecs; ++i) {
      const int f = threadIdx.x + i * 128;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kKContig) {
        const int r = f >> 2, kq = (f & 3) * 4;  // vector along k
        if (row0 + r < rows && k0 + kq < K) v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * ld + k0 + kq));
      } else {
        const int k = f / (ROWS / 4), rq = (f % (ROWS / 4)) * 4;  // vector along the row index
        if (k0 + k < K && row0 + rq < rows) v[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(k0 + k) * ld + row0 + rq));
      }
    }
  }
  __device__ __forceinline__ void store(float (*dst)[ROWS + 4]) const {
#pragma unroll
    for (int i = 0; i < kVecs; ++i) {
      const int f = threadIdx.x + i * 128;
      if (kKContig) {
        const int r = f >> 2, kq = (f & 3) * 4;
        dst[kq + 0][r] = v[i].x;
        dst[kq + 1][r] = v[i].y;
        dst[kq + 2][r] = v[i].z;
        dst[kq + 3][r] = v[i].w;
      } else {
        const int k = f / (ROWS / 4), rq = (f % (ROWS / 4)) * 4;
        *reinterpret_cast<float4*>(&dst[k][rq]) = v[i];
      }
    }
  }
};

template <bool kAK, bool kBK>
__global__ void __launch_bounds__(128) sgemm_kernel(const GemmArgs g) {
  pdl_sync();
  __shared__ __align__(16) float As[2][kTK][kTM + 4];
  __shared__ __align__(16) float Bs[2][kTK][kTN + 4];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;  // rows {4ty.., 32+4ty..}, columns {4tx.., 64+4tx..}
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool split = gridDim.z > 1;
  const int kbeg = split ? blockIdx.z * g.kchunk : 0;
  const int kend = split ? min(g.K, kbeg + g.kchunk) : g.K;
  TileLoader<kTM, kAK> la;
  TileLoader<kTN, kBK> lb;
  la.load(g.A, g.lda, m0, g.M, kbeg, kend);
  lb.load(g.B, g.ldb, n0, g.N, kbeg, kend);
  la.store(As[0]);
  lb.store(Bs[0]);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += kTK) {
    const bool more = k0 + kTK < kend;
    if (more) {
      la.load(g.A, g.lda, m0, g.M, k0 + kTK, kend);
      lb.load(g.B, g.ldb, n0, g.N, k0 + kTK, kend);
    }
    // fragments of step kk + 1 are fetched while the 64 FMAs of step kk issue (one warp per scheduler: an exposed
    // shared-memory round trip per step cost a third of the FMA rate)
    float4 fa[2][2], fb[2][2];
    fa[0][0] = *reinterpret_cast<const float4*>(&As[buf][0][ty * 4]);
    fa[0][1] = *reinterpret_cast<const float4*>(&As[buf][0][32 + ty * 4]);
    fb[0][0] = *reinterpret_cast<const float4*>(&Bs[buf][0][tx * 4]);
    fb[0][1] = *reinterpret_cast<const float4*>(&Bs[buf][0][64 + tx * 4]);
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const int cur = kk & 1, nxt = cur ^ 1;
      if (kk + 1 < kTK) {
        fa[nxt][0] = *reinterpret_cast<const float4*>(&As[buf][kk + 1][ty * 4]);
        fa[nxt][1] = *reinterpret_cast<const float4*>(&As[buf][kk + 1][32 + ty * 4]);
        fb[nxt][0] = *reinterpret_cast<const float4*>(&Bs[buf][kk + 1][tx * 4]);
        fb[nxt][1] = *reinterpret_cast<const float4*>(&Bs[buf][kk + 1][64 + tx * 4]);
      }
      const float av[8] = {fa[cur][0].x, fa[cur][0].y, fa[cur][0].z, fa[cur][0].w,
                           fa[cur][1].x, fa[cur][1].y, fa[cur][1].z, fa[cur][1].w};
      const float bv[8] = {fb[cur][0].x, fb[cur][0].y, fb[cur][0].z, fb[cur][0].w,
                           fb[cur][1].x, fb[cur][1].y, fb[cur][1].z, fb[cur][1].w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      la.store(As[buf ^ 1]);
      lb.store(Bs[buf ^ 1]);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 32 + ty * 4 + (i - 4));
    if (row >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      if (col >= g.N) continue;  // N is a multiple of 4 whenever it is the contiguous dimension of C
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (col + j >= g.N) {
          v[j] = 0.f;
          continue;
        }
        if (g.bias) v[j] += g.bias[col + j];
        if (g.relu) v[j] = fmaxf(v[j], 0.f);
        if (g.mask) v[j] = g.mask[(size_t)row * g.ldm + col + j] > 0.f ? v[j] : 0.f;
      }
      float* dst = &g.C[(size_t)row * g.ldc + col];
      if (split) {
        float* part = g.splitk + ((size_t)blockIdx.z * g.M + row) * g.N + col;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < g.N) part[j] = v[j];
      } else if (col + 3 < g.N && (g.ldc & 3) == 0) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (g.accumulate) {
          const float4 c = *reinterpret_cast<const float4*>(dst);
          o.x += c.x;
          o.y += c.y;
          o.z += c.z;
          o.w += c.w;
        }
        *reinterpret_cast<float4*>(dst) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (col + j < g.N) dst[j] = g.accumulate ? dst[j] + v[j] : v[j];
      }
    }
  }
}

// C[m][n] (+)= sum over slices (in slice order) of part[z][m][n]
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ C, int M,
                                                            int N, int ldc, int slices, int accumulate) {
  pdl_sync();
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
    float a = accumulate ? C[(size_t)m * ldc + n] : 0.f;
    for (int z = 0; z < slices; ++z) a += part[(size_t)z * total + i];
    C[(size_t)m * ldc + n] = a;
  }
}

template <bool kAK, bool kBK>
cudaError_t run_gemm(const GemmArgs& g, cudaStream_t s) {
  // contiguous dimensions must be whole float4s, and every operand base 16-byte aligned
  if ((kAK ? g.K : g.M) % 4 != 0 || (kBK ? g.K : g.N) % 4 != 0 || g.lda % 4 != 0 || g.ldb % 4 != 0 ||
      ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.B) | reinterpret_cast<uintptr_t>(g.C)) & 15) != 0)
    return cudaErrorInvalidValue;
  dim3 grid((g.N + kTN - 1) / kTN, (g.M + kTM - 1) / kTM);
  GemmArgs a = g;
  // Few-tile GEMMs with a long K (the M = clips rows of the factorised first layer) are split along K so that they
  // fill the machine; the slices land in a.splitk and are added in slice order by a second kernel (no epilogue).
  const int tiles = grid.x * grid.y;
  if (a.kchunk < 0) {
    a.kchunk = 0;
    if (tiles < 100 && !a.bias && !a.relu && !a.mask && a.splitk != nullptr) {
      int splits = std::min((148 + tiles - 1) / tiles, a.K / 64);
      while (splits > 1 && (size_t)splits * a.M * a.N > kSplitKFloats) --splits;
      if (splits > 1) {
        a.kchunk = ((a.K + splits - 1) / splits + kTK - 1) / kTK * kTK;
        grid.z = (a.K + a.kchunk - 1) / a.kchunk;
      }
    }
  }
  launch_kernel(sgemm_kernel<kAK, kBK>, grid, 128, 0, s, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || grid.z <= 1) return e;
  launch_kernel(splitk_reduce_kernel, 148 * 4, 256, 0, s, (const float*)a.splitk, a.C, a.M, a.N, a.ldc, (int)grid.z,
                a.accumulate);
  return cudaGetLastError();
}

// S[row] = H4[row,:] . w5 + b5
__global__ void __launch_bounds__(256) lang_score_kernel(const float* __restrict__ H4, const float* __restrict__ w5,
                                                         const float* __restrict__ b5, float* __restrict__ S, int H) {
  pdl_sync();
  __shared__ float red[32];
  const int row = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < H; i += blockDim.x) acc = fmaf(H4[(size_t)row * H + i], w5[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) S[row] = v + b5[0];
  }
}

// InfoNCE over 1 positive + 4 negatives, 3 targets per clip (trainer.py:93-117).  ONE block: a thread handles clips
// tid, tid + blockDim, ...; the four metric sums go through a fixed shuffle / shared-memory tree (deterministic).
__global__ void __launch_bounds__(256) lang_loss_kernel(const float* __restrict__ S, const float* __restrict__ mask,
                                                        float* __restrict__ dS, int B, float langw,
                                                        float* __restrict__ metrics) {
  pdl_sync();
  __shared__ float red[4 * 32];
  const float invB = 1.0f / (float)B;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // rewloss, rewacc1..3
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float mk = mask[b];
    float loss = 0.f;
    for (int t = 0; t < 3; ++t) {
      const float pos = S[t * B + b];
      float neg[4];
      neg[0] = S[(3 + t) * B + b];
      for (int i = 0; i < 3; ++i) neg[1 + i] = S[(6 + 3 * i + t) * B + b];
      const float ep = expf(pos);
      float en[4], sum = 0.f, mx = neg[0];
      for (int k = 0; k < 4; ++k) {
        en[k] = expf(neg[k]);
        sum += en[k];
        mx = fmaxf(mx, neg[k]);
      }
      const float Dn = kLossEps + ep + sum;
      const float r = ep / Dn;
      loss += -logf(kLossEps + r);
      acc[1 + t] += (mx < pos) ? 1.f : 0.f;
      if (dS) {
        const float f = langw * mk * invB * (1.0f / 3.0f);
        const float gr = -f / (kLossEps + r);
        dS[t * B + b] = gr * (r - r * r);
        dS[(3 + t) * B + b] = gr * (-r * en[0] / Dn);
        for (int i = 0; i < 3; ++i) dS[(6 + 3 * i + t) * B + b] = gr * (-r * en[1 + i] / Dn);
      }
    }
    acc[0] += mk * loss * (1.0f / 3.0f);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0)
    for (int i = 0; i < 4; ++i) red[i * 32 + warp] = acc[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < nwarp; ++w)
      for (int i = 0; i < 4; ++i) tot[i] += red[i * 32 + w];
    metrics[kRewLoss] += tot[0] * invB;
    metrics[kFullLoss] += langw * tot[0] * invB;
    for (int t = 0; t < 3; ++t) metrics[kRewAcc1 + t] += tot[1 + t] * invB;
  }
}

// dH4[row, j] = dS[row] * w5[j] * (H4[row, j] > 0)
__global__ void __launch_bounds__(256) lang_dscore_kernel(const float* __restrict__ dS, const float* __restrict__ w5,
                                                          const float* __restrict__ H4, float* __restrict__ dH4,
                                                          int rows, int H) {
  pdl_sync();
  const size_t total = (size_t)rows * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / H), j = (int)(i - (size_t)row * H);
    dH4[i] = H4[i] > 0.f ? dS[row] * w5[j] : 0.f;
  }
}

// out[j] = sum_row scale[row] * Mtx[row, j]   (scale == null -> 1).  Block = 16 columns x 64 row-slices; the slices are
// added in slice order (no atomics: deterministic).
__global__ void __launch_bounds__(1024) col_sum_kernel(const float* __restrict__ Mtx, const float* __restrict__ scale,
                                                       float* __restrict__ out, int rows, int cols) {
  pdl_sync();
  __shared__ float part[64][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j = blockIdx.x * 16 + tx;
  float acc = 0.f;
  if (j < cols)
    for (int r = ty; r < rows; r += 64) acc = fmaf(scale ? scale[r] : 1.f, Mtx[(size_t)r * cols + j], acc);
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < cols) {
    float v = 0.f;
#pragma unroll 8
    for (int i = 0; i < 64; ++i) v += part[i][tx];
    out[j] = v;
  }
}

__global__ void vec_sum_kernel(const float* __restrict__ v, float* __restrict__ out, int n) {
  pdl_sync();
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float x = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    x = warp_sum(x);
    if (threadIdx.x == 0) out[0] = x;
  }
}

__device__ __forceinline__ float tc_round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// Operand preparation of the tensor-core path, one 32 x 32 tile of src [R][C] per block: optional ReLU gate
// (gate[r][c] > 0), optional dense copy of the gated values (may alias src), the row-major split copy [R][3C] and the
// TRANSPOSED split copy [C][3Rp] (rows >= R zero filled: the padded tail of a reduction axis).  Split order: 0 =
// activation (hi | lo | hi), 1 = weight (hi | hi | lo).
__global__ void __launch_bounds__(256) lang_prep_kernel(const float* src, const float* __restrict__ gate,
                                                        float* dense, float* __restrict__ split,
                                                        float* __restrict__ tsplit, int R, int C, int Rp, int s_weights,
                                                        int t_weights) {
  __shared__ float tile[32][33];
  pdl_sync();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int rl = ty + 8 * j, r = blockIdx.y * 32 + rl;
    float v = 0.f;
    if (r < R) {
      v = src[(size_t)r * C + c];
      if (gate != nullptr && !(gate[(size_t)r * C + c] > 0.f)) v = 0.f;
      if (dense != nullptr) dense[(size_t)r * C + c] = v;
      if (split != nullptr) {
        const float hi = tc_round_tf32(v), lo = tc_round_tf32(v - hi);
        float* o = split + (size_t)r * 3 * C;
        o[c] = hi;
        o[C + c] = s_weights ? hi : lo;
        o[2 * C + c] = s_weights ? lo : hi;
      }
    }
    tile[rl][tx] = v;
  }
  if (tsplit == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cl = ty + 8 * j;
    const float v = tile[tx][cl];  // element (row by*32 + tx, column bx*32 + cl)
    const float hi = tc_round_tf32(v), lo = tc_round_tf32(v - hi);
    float* o = tsplit + (size_t)(blockIdx.x * 32 + cl) * 3 * Rp + blockIdx.y * 32 + tx;
    o[0] = hi;
    o[Rp] = t_weights ? hi : lo;
    o[2 * Rp] = t_weights ? lo : hi;
  }
}

__global__ void lang_fill_kernel(float* p, int n, float v) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

cudaError_t launch_prep(const float* src, const float* gate, float* dense, float* split, float* tsplit, int R, int C,
                        int Rp, int s_weights, int t_weights, cudaStream_t s) {
  launch_kernel(lang_prep_kernel, dim3(C / 32, Rp / 32), dim3(256), 0, s, src, gate, dense, split, tsplit, R, C, Rp,
                s_weights, t_weights);
  return cudaGetLastError();
}

size_t up64(size_t v) { return (v + 63) / 64 * 64; }

}  // namespace

size_t lang_tc_floats(const LangDims& d) {
  const size_t r = (size_t)d.rows(), H = (size_t)d.H, Rp = (r + 31) / 32 * 32;
  return 3 * up64(r * 3 * H) + 3 * up64(H * 3 * Rp) + up64(r * 3 * H) + up64(H * 3 * Rp) + 6 * up64(H * 3 * H) + up64(H);
}

std::string lang_tc_plan(const LangDims& d, const LangParams& p, const LangWorkspace& ws, float* base, LangTc* tc) {
  *tc = LangTc();
  const int rows = d.rows(), H = d.H;
  if (H % 64 != 0 || H > 1024) return std::string();  // plain fp32 SIMT path (tc->enabled stays false)
  const int Rp = (rows + 31) / 32 * 32;
  tc->Rp = Rp;
  float* q = base;
  auto take = [&](size_t n) {
    float* r = q;
    q += up64(n);
    return r;
  };
  for (int i = 0; i < 3; ++i) tc->Hs[i] = take((size_t)rows * 3 * H);
  for (int i = 0; i < 3; ++i) tc->HTs[i] = take((size_t)H * 3 * Rp);
  tc->dHs = take((size_t)rows * 3 * H);
  tc->dHTs = take((size_t)H * 3 * Rp);
  for (int i = 0; i < 3; ++i) tc->Ws[i] = take((size_t)H * 3 * H);
  for (int i = 0; i < 3; ++i) tc->WTs[i] = take((size_t)H * 3 * H);
  tc->ones = take((size_t)H);
  // out[M][H] = act(src[M][K3] . w[H][K3]^T (+ bias)): the conv kernel's tf32 tier as a plain GEMM, 64-wide N tiles
  auto gemm = [&](const float* src, int M, int K3, const float* w, const float* bias, int relu, float* out,
                  ConvPlan* plan) {
    GatherConv g;
    g.src = src;
    g.N = M;
    g.H = g.W = g.P = g.Q = 1;
    g.C = K3;
    g.stride = 1;
    g.ntaps = 1;
    g.wpk = w;
    g.Cout = H;
    g.out = out;
    g.ldo = H;
    if (bias != nullptr) {
      g.ep_scale = tc->ones;
      g.ep_shift = bias;
    }
    g.ep_relu = relu;
    g.ep_exact = 1;
    g.tf32 = 1;
    g.bn = 64;
    return plan_conv(g, plan);
  };
  int cur = 0;
  for (int l = 3; l >= 1; --l) {  // the backward loop's ping-pong: dH_l lives in dH[cur], dH_{l-1} goes to dH[cur ^ 1]
    std::string e = gemm(tc->dHs, rows, 3 * H, tc->WTs[l - 1], nullptr, 0, ws.dH[cur ^ 1], &tc->dx[l - 1]);
    if (e.empty()) e = gemm(tc->dHTs, H, 3 * Rp, tc->HTs[l - 1], nullptr, 0, p.dw[l], &tc->dw[l - 1]);
    if (e.empty()) e = gemm(tc->Hs[l - 1], rows, 3 * H, tc->Ws[l - 1], p.b[l], 1, ws.Hact[l], &tc->fwd[l - 1]);
    if (!e.empty()) return "language head (tensor-core path): " + e;
    cur ^= 1;
  }
  tc->enabled = true;
  return std::string();
}

size_t lang_workspace_floats(const LangDims& d) {
  const size_t r = (size_t)d.rows();
  auto up = [](size_t v) { return (v + 63) / 64 * 64; };
  return 2 * (2 * up((size_t)d.B * d.H) + up((size_t)5 * d.B * d.H)) + 6 * up(r * d.H) + 2 * up(r) + kSplitKFloats;
}

void lang_carve_workspace(float* base, const LangDims& d, LangWorkspace* ws) {
  const size_t r = (size_t)d.rows();
  auto up = [](size_t v) { return (v + 63) / 64 * 64; };
  float* p = base;
  // dU, dV, dLc are contiguous: one memset clears them
  ws->dU = p;
  p += up((size_t)d.B * d.H);
  ws->dV = p;
  p += up((size_t)5 * d.B * d.H);
  ws->dLc = p;
  p += up((size_t)d.B * d.H);
  ws->U = p;
  p += up((size_t)d.B * d.H);
  ws->V = p;
  p += up((size_t)5 * d.B * d.H);
  ws->Lc = p;
  p += up((size_t)d.B * d.H);
  for (int i = 0; i < 4; ++i) {
    ws->Hact[i] = p;
    p += up(r * d.H);
  }
  for (int i = 0; i < 2; ++i) {
    ws->dH[i] = p;
    p += up(r * d.H);
  }
  ws->S = p;
  p += up(r);
  ws->dS = p;
  p += up(r);
  ws->splitk = p;
}

cudaError_t lang_head_run(const LangDims& d, const LangParams& p, const LangWorkspace& ws, const float* E, float* dE,
                          const int* perms, const float* lang_emb, const float* lang_mask, float langw,
                          float* metrics, int* launches, cudaStream_t s, const LangTc* tc) {
  const bool use_tc = tc != nullptr && tc->enabled;
  const int rows = d.rows(), H = d.H, K1 = d.k1();
  int n = 0;
  cudaError_t e;
#define R3M_TRY(expr)            \
  do {                           \
    e = (expr);                  \
    ++n;                         \
    if (e != cudaSuccess) {      \
      if (launches) *launches = n; \
      return e;                  \
    }                            \
  } while (0)

  const int B = d.B, D = d.D;
  // ---- forward layer 1 (factorised): U = E0 . W1a^T, V = E . W1b^T, Lc = L . W1c^T, then gather-add + bias + ReLU
  {
    // few-tile products: split along K (kchunk = -1: automatic), slices reduced in order through ws.splitk
    GemmArgs gu{E, p.w[0], ws.U, B, H, D, 5 * D, K1, H, nullptr, 0, nullptr, 0, 0, -1, ws.splitk};
    R3M_TRY((run_gemm<true, true>(gu, s)));
    GemmArgs gv{E, p.w[0] + D, ws.V, 5 * B, H, D, D, K1, H, nullptr, 0, nullptr, 0, 0, -1, ws.splitk};
    R3M_TRY((run_gemm<true, true>(gv, s)));
    GemmArgs gl{lang_emb, p.w[0] + 2 * D, ws.Lc, B, H, d.L, d.L, K1, H, nullptr, 0, nullptr, 0, 0, -1, ws.splitk};
    R3M_TRY((run_gemm<true, true>(gl, s)));
    launch_kernel(lang_layer1_kernel, rows, 256, 0, s, ws.U, ws.V, ws.Lc, p.b[0], perms, ws.Hact[0], d);
    R3M_TRY(cudaGetLastError());
  }
  // ---- layers 2-4: Linear + ReLU, then Linear(H -> 1)
  if (use_tc) {
    launch_kernel(lang_fill_kernel, dim3((H + 255) / 256), dim3(256), 0, s, tc->ones, H, 1.0f);
    R3M_TRY(cudaGetLastError());
    for (int l = 1; l < 4; ++l) {
      // this step's weights: split W_l (forward operand) and, for the backward pass, split W_l^T (dX operand)
      R3M_TRY(launch_prep(p.w[l], nullptr, nullptr, tc->Ws[l - 1], dE ? tc->WTs[l - 1] : nullptr, H, H, H, 1, 1, s));
      // H_{l-1}: split (forward operand) and, for the backward pass, split H_{l-1}^T (dW operand)
      R3M_TRY(launch_prep(ws.Hact[l - 1], nullptr, nullptr, tc->Hs[l - 1], dE ? tc->HTs[l - 1] : nullptr, rows, H, tc->Rp,
                          0, 1, s));
      R3M_TRY(run_conv(tc->fwd[l - 1], s));
    }
  } else {
  for (int l = 1; l < 4; ++l) {
    GemmArgs g{ws.Hact[l - 1], p.w[l], ws.Hact[l], rows, H, H, H, H, H, p.b[l], 1, nullptr, 0, 0};
    R3M_TRY((run_gemm<true, true>(g, s)));
  }
  }
  launch_kernel(lang_score_kernel, rows, 256, 0, s, ws.Hact[3], p.w[4], p.b[4], ws.S, H);
  R3M_TRY(cudaGetLastError());
  launch_kernel(lang_loss_kernel, 1, 256, 0, s, ws.S, lang_mask, dE ? ws.dS : nullptr, d.B, langw, metrics);
  R3M_TRY(cudaGetLastError());
  if (dE) {
    // ---- backward
    launch_kernel(col_sum_kernel, dim3((H + 15) / 16, 1), 1024, 0, s, ws.Hact[3], ws.dS, p.dw[4], rows, H);  // dw5 = dS^T H4
    R3M_TRY(cudaGetLastError());
    launch_kernel(vec_sum_kernel, 1, 256, 0, s, ws.dS, p.db[4], rows);
    R3M_TRY(cudaGetLastError());
    launch_kernel(lang_dscore_kernel, 148 * 4, 256, 0, s, ws.dS, p.w[4], ws.Hact[3], ws.dH[0], rows, H);
    R3M_TRY(cudaGetLastError());
    int cur = 0;
    if (use_tc) {
      for (int l = 3; l >= 1; --l) {
        float* dHl = ws.dH[cur];
        // dH_l: apply the ReLU gate of layer l to the raw dX product (l = 3 arrives gated from lang_dscore_kernel), keep
        // the gated values (bias gradient, layer-1 backward), split it (dX operand) and split its transpose (dW operand)
        R3M_TRY(launch_prep(dHl, l < 3 ? ws.Hact[l] : nullptr, l < 3 ? dHl : nullptr, tc->dHs, tc->dHTs, rows, H, tc->Rp, 0,
                            0, s));
        launch_kernel(col_sum_kernel, dim3((H + 15) / 16, 1), 1024, 0, s, (const float*)dHl, (const float*)nullptr,
                      p.db[l], rows, H);
        R3M_TRY(cudaGetLastError());
        R3M_TRY(run_conv(tc->dw[l - 1], s));
        R3M_TRY(run_conv(tc->dx[l - 1], s));
        cur ^= 1;
      }
      // the raw dX of layer 1's output still needs its ReLU gate
      R3M_TRY(launch_prep(ws.dH[cur], ws.Hact[0], ws.dH[cur], nullptr, nullptr, rows, H, tc->Rp, 0, 0, s));
    } else
    for (int l = 3; l >= 1; --l) {
      const float* dHl = ws.dH[cur];
      // db_l = column sums of dH_l;  dW_l[n][k] = sum_m dH_l[m][n] * H_{l-1}[m][k]
      launch_kernel(col_sum_kernel, dim3((H + 15) / 16, 1), 1024, 0, s, dHl, nullptr, p.db[l], rows, H);
      R3M_TRY(cudaGetLastError());
      GemmArgs gw{dHl, ws.Hact[l - 1], p.dw[l], H, H, rows, H, H, H, nullptr, 0, nullptr, 0, 0};
      R3M_TRY((run_gemm<false, false>(gw, s)));
      // dH_{l-1}[m][k] = sum_n dH_l[m][n] * W_l[n][k], gated by ReLU of the layer below
      GemmArgs gx{dHl, p.w[l], ws.dH[cur ^ 1], rows, H, H, H, H, H, nullptr, 0, ws.Hact[l - 1], H, 0};
      R3M_TRY((run_gemm<true, false>(gx, s)));
      cur ^= 1;
    }
    // ---- layer 1 backward (factorised)
    const float* dpre = ws.dH[cur];
    launch_kernel(col_sum_kernel, dim3((H + 15) / 16, 1), 1024, 0, s, dpre, nullptr, p.db[0], rows, H);
    R3M_TRY(cudaGetLastError());
    launch_kernel(lang_layer1_bwd_kernel, 7 * B, 256, (size_t)9 * B, s, dpre, perms, ws.dU, ws.dV, ws.dLc, d);
    R3M_TRY(cudaGetLastError());
    // dW1 = [dU^T E0 | dV^T E | dLc^T L]   (column blocks of the [H][2D+768] gradient)
    GemmArgs wa{ws.dU, E, p.dw[0], H, D, B, H, 5 * D, K1, nullptr, 0, nullptr, 0, 0};
    R3M_TRY((run_gemm<false, false>(wa, s)));
    GemmArgs wb{ws.dV, E, p.dw[0] + D, H, D, 5 * B, H, D, K1, nullptr, 0, nullptr, 0, 0};
    R3M_TRY((run_gemm<false, false>(wb, s)));
    GemmArgs wc{ws.dLc, lang_emb, p.dw[0] + 2 * D, H, d.L, B, H, d.L, K1, nullptr, 0, nullptr, 0, 0};
    R3M_TRY((run_gemm<false, false>(wc, s)));
    // dE0 += dU . W1a,  dE += dV . W1b   (the sentence embedding is frozen: no gradient through W1c's input)
    GemmArgs ea{ws.dU, p.w[0], dE, B, D, H, H, K1, 5 * D, nullptr, 0, nullptr, 0, 1, -1, ws.splitk};
    R3M_TRY((run_gemm<true, false>(ea, s)));
    GemmArgs eb{ws.dV, p.w[0] + D, dE, 5 * B, D, H, H, K1, D, nullptr, 0, nullptr, 0, 1, -1, ws.splitk};
    R3M_TRY((run_gemm<true, false>(eb, s)));
  }
#undef R3M_TRY
  if (launches) *launches = n;
  return cudaSuccess;
}

}  // namespace r3m
