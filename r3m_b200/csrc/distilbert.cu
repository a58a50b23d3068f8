// DistilBERT sentence encoder (see distilbert.h).  Arithmetic restated from transformers' DistilBertModel
// (models/distilbert/modeling_distilbert.py: Embeddings, DistilBertSelfAttention + eager_attention_forward, FFN,
// TransformerBlock), which is what r3m/models/models_language.py:23-35 runs under torch.no_grad().
#include "distilbert.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "elementwise.cuh"
#include "launch.h"
#include "ptx.cuh"

namespace r3m {

namespace {

constexpr float kLnEps = 1e-12f;  // nn.LayerNorm(eps=1e-12) everywhere in DistilBERT
constexpr int kHeadDim = 64;
constexpr int kMaxVec = 8;  // float4 per lane: dim <= 1024

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// fp32 on tf32 tensor cores without losing fp32: x = hi + lo with both parts tf32 (round-to-nearest), and
//   a . w  ~=  a_hi . w_hi + a_lo . w_hi + a_hi . w_lo      (the dropped lo . lo term is ~2^-22 relative),
// evaluated as ONE GEMM over a three-times-longer reduction axis: activations stored [hi | lo | hi], weights
// [hi | hi | lo] (fp32 accumulation in the tensor core).
__device__ __forceinline__ void split_tf32(float v, float* hi, float* lo) {
  *hi = round_tf32(v);
  *lo = round_tf32(v - *hi);
}
__device__ __forceinline__ void split4(const float4& v, float4* hi, float4* lo) {
  split_tf32(v.x, &hi->x, &lo->x);
  split_tf32(v.y, &hi->y, &lo->y);
  split_tf32(v.z, &hi->z, &lo->z);
  split_tf32(v.w, &hi->w, &lo->w);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// LayerNorm of one token held by one warp (lane owns float4 i at element lane*4 + 128*i): biased variance around the
// mean (two passes over registers), y = (v - mean) * rsqrt(var + eps) * w + b.  Writes the exact fp32 result (residual
// stream) and its tf32-rounded copy (operand of the next Linear: the tensor core would truncate instead of round).
template <bool EMBED>
__global__ void __launch_bounds__(128) layernorm_kernel(const float* __restrict__ h, const int* __restrict__ ids,
                                                        const float* __restrict__ word, const float* __restrict__ pos,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ x, float* __restrict__ xr, int M, int T,
                                                        int dim, int vocab, int* error_flag) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (m >= M) return;
  const int nv = dim >> 7;
  float4 v[kMaxVec];
  if (EMBED) {
    int id = ids[m];
    if (id < 0 || id >= vocab) {
      if (lane == 0) atomicExch(error_flag, 21);
      id = 0;
    }
    const float4* wr = reinterpret_cast<const float4*>(word + static_cast<size_t>(id) * dim);
    const float4* pr = reinterpret_cast<const float4*>(pos + static_cast<size_t>(m % T) * dim);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        const float4 a = __ldg(wr + lane + 32 * i), c = __ldg(pr + lane + 32 * i);
        v[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
      }
  } else {
    const float4* hr = reinterpret_cast<const float4*>(h + static_cast<size_t>(m) * dim);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) v[i] = hr[lane + 32 * i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / static_cast<float>(dim);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + c * c) + (d * d + e * e);
    }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(dim) + kLnEps);
  float4* xo = reinterpret_cast<float4*>(x + static_cast<size_t>(m) * dim);
  float4* xro = reinterpret_cast<float4*>(xr + static_cast<size_t>(m) * 3 * dim);  // [hi | lo | hi]
  const float4* wv = reinterpret_cast<const float4*>(w);
  const float4* bv = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i)
    if (i < nv) {
      const float4 g = __ldg(wv + lane + 32 * i), o = __ldg(bv + lane + 32 * i);
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + o.x;
      y.y = (v[i].y - mean) * rstd * g.y + o.y;
      y.z = (v[i].z - mean) * rstd * g.z + o.z;
      y.w = (v[i].w - mean) * rstd * g.w + o.w;
      xo[lane + 32 * i] = y;
      float4 hi, lo;
      split4(y, &hi, &lo);
      xro[lane + 32 * i] = hi;
      xro[nv * 32 + lane + 32 * i] = lo;
      xro[2 * nv * 32 + lane + 32 * i] = hi;
    }
}

// softmax(q k^T / 8 + mask) v for one (sentence, head) per block; fp32 throughout (the scores of a 64-wide head are
// not worth a tensor-core tile at sentence lengths of 10-30 tokens).  One warp per query, keys in tiles of 64 staged in
// shared memory (single tile: loaded once), online softmax across tiles.  Padded KEYS (mask == 0) get weight zero —
// transformers adds finfo.min to their scores, which underflows to the same thing whenever a sentence has at least one
// real token (always: [CLS] / [SEP]); padded QUERIES are computed like any other (the reference's mean includes them).
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                        float* __restrict__ ctx, int T, int dim, int heads) {
  __shared__ float sK[64][kHeadDim + 1];
  __shared__ float sV[64][kHeadDim];
  __shared__ float sQ[4][kHeadDim];
  __shared__ float sMask[64];
  pdl_sync();
  const int b = blockIdx.x / heads, hd = blockIdx.x % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * dim;
  const float* base = qkv + static_cast<size_t>(b) * T * ld + hd * kHeadDim;
  const int ntiles = (T + 63) / 64;
  auto load_tile = [&](int t0) {
    for (int i = threadIdx.x; i < 64 * (kHeadDim / 4); i += 128) {
      const int j = i / (kHeadDim / 4), c = i % (kHeadDim / 4);
      float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = k4;
      if (t0 + j < T) {
        const float* row = base + static_cast<size_t>(t0 + j) * ld;
        k4 = *reinterpret_cast<const float4*>(row + dim + 4 * c);
        v4 = *reinterpret_cast<const float4*>(row + 2 * dim + 4 * c);
      }
      sK[j][4 * c] = k4.x;
      sK[j][4 * c + 1] = k4.y;
      sK[j][4 * c + 2] = k4.z;
      sK[j][4 * c + 3] = k4.w;
      *reinterpret_cast<float4*>(&sV[j][4 * c]) = v4;
    }
    if (threadIdx.x < 64) sMask[threadIdx.x] = (t0 + threadIdx.x < T) ? mask[b * T + t0 + threadIdx.x] : 0.f;
  };
  if (ntiles == 1) {
    load_tile(0);
    __syncthreads();
  }
  for (int q0 = 0; q0 < T; q0 += 4) {
    const int qi = q0 + warp;
    const bool active = qi < T;
    if (active) {
      const float* qrow = base + static_cast<size_t>(qi) * ld;
      sQ[warp][lane] = qrow[lane];
      sQ[warp][lane + 32] = qrow[lane + 32];
    }
    __syncwarp();
    float mx = -INFINITY, l = 0.f, acc0 = 0.f, acc1 = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      if (ntiles > 1) {
        __syncthreads();
        load_tile(t * 64);
        __syncthreads();
      }
      if (!active) continue;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
      for (int d = 0; d < kHeadDim; ++d) {
        const float qd = sQ[warp][d];
        s0 = fmaf(qd, sK[lane][d], s0);
        s1 = fmaf(qd, sK[lane + 32][d], s1);
      }
      s0 = (sMask[lane] != 0.f) ? s0 * 0.125f : -INFINITY;
      s1 = (sMask[lane + 32] != 0.f) ? s1 * 0.125f : -INFINITY;
      const float mt = warp_max(fmaxf(s0, s1));
      if (mt == -INFINITY) continue;  // every key of this tile is padding
      const float mn = fmaxf(mx, mt);
      const float corr = (mx == -INFINITY) ? 0.f : expf(mx - mn);
      const float p0 = (s0 == -INFINITY) ? 0.f : expf(s0 - mn);
      const float p1 = (s1 == -INFINITY) ? 0.f : expf(s1 - mn);
      l = l * corr + warp_sum(p0 + p1);
      acc0 *= corr;
      acc1 *= corr;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p0, j);
        acc0 = fmaf(pj, sV[j][lane], acc0);
        acc1 = fmaf(pj, sV[j][lane + 32], acc1);
      }
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p1, j);
        acc0 = fmaf(pj, sV[j + 32][lane], acc0);
        acc1 = fmaf(pj, sV[j + 32][lane + 32], acc1);
      }
      mx = mn;
    }
    if (active) {
      const float inv = (l > 0.f) ? 1.f / l : 0.f;
      float* o = ctx + static_cast<size_t>(b * T + qi) * 3 * dim + hd * kHeadDim;  // [hi | lo | hi]
      float h0, l0, h1, l1;
      split_tf32(acc0 * inv, &h0, &l0);
      split_tf32(acc1 * inv, &h1, &l1);
      o[lane] = h0;
      o[lane + 32] = h1;
      o[dim + lane] = l0;
      o[dim + lane + 32] = l1;
      o[2 * dim + lane] = h0;
      o[2 * dim + lane + 32] = h1;
    }
    __syncwarp();
  }
}

// last_hidden_state.mean(1): every position, padding included (models_language.py:34); summed in position order
__global__ void __launch_bounds__(256) mean_pool_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                        float* __restrict__ hidden, int T, int dim) {
  pdl_sync();
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) {
      const float v = x[(static_cast<size_t>(b) * T + t) * dim + d];
      if (hidden) hidden[(static_cast<size_t>(b) * T + t) * dim + d] = v;
      s += v;
    }
    out[static_cast<size_t>(b) * dim + d] = s / static_cast<float>(T);
  }
}

// rows [R][K] fp32 -> [R][3K]: activations (weights == 0) as [hi | lo | hi], weights (weights == 1) as [hi | hi | lo]
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                     long long rows, int K, int weights) {
  pdl_sync();
  const long long total = rows * (K / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / (K / 4);
    const int c = static_cast<int>(i % (K / 4));
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    float4 hi, lo;
    split4(v, &hi, &lo);
    float4* o = reinterpret_cast<float4*>(dst + r * 3 * K);
    o[c] = hi;
    o[K / 4 + c] = weights ? hi : lo;
    o[2 * (K / 4) + c] = weights ? lo : hi;
  }
}

__global__ void fill_kernel(float* p, int n, float v) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

std::string DistilBert::create(const BertDims& d, DistilBert** out) {
  if (d.dim % 128 != 0 || d.dim > 128 * kMaxVec) return "distilbert: dim must be a multiple of 128, at most 1024";
  if (d.heads < 1 || d.dim != d.heads * kHeadDim) return "distilbert: head size must be 64";
  if (d.dim % 256 != 0 || d.ffn % 256 != 0) return "distilbert: dim and ffn width must be multiples of 256";
  if (d.layers < 1 || d.vocab < 1 || d.max_pos < 1) return "distilbert: bad dimensions";
  DistilBert* m = new DistilBert();
  m->d_ = d;
  size_t off = 0;
  auto add = [&](const std::string& name, int kind, int d0, int d1) {
    TensorInfo ti;
    ti.name = name;
    ti.kind = kind;
    ti.offset = off;
    ti.dims[0] = d0;
    ti.dims[1] = d1;
    ti.ndim = d1 > 0 ? 2 : 1;
    if (d1 <= 0) ti.dims[1] = 1;
    m->tensors_.push_back(ti);
    const size_t at = off;
    off += static_cast<size_t>(d0) * (d1 > 0 ? d1 : 1);
    return at;
  };
  m->word_off_ = add("embeddings.word_embeddings.weight", kLinearW, d.vocab, d.dim);
  m->pos_off_ = add("embeddings.position_embeddings.weight", kLinearW, d.max_pos, d.dim);
  m->eln_w_ = add("embeddings.LayerNorm.weight", kVector, d.dim, 0);
  m->eln_b_ = add("embeddings.LayerNorm.bias", kVector, d.dim, 0);
  m->lin_begin_ = off;
  static const char* lin_names[6] = {"attention.q_lin", "attention.k_lin", "attention.v_lin", "attention.out_lin",
                                     "ffn.lin1", "ffn.lin2"};
  for (int l = 0; l < d.layers; ++l) {
    const std::string pre = "transformer.layer." + std::to_string(l) + ".";
    Layer L;
    for (int i = 0; i < 4; ++i) {
      L.w[i] = add(pre + lin_names[i] + ".weight", kLinearW, d.dim, d.dim);
      L.b[i] = add(pre + lin_names[i] + ".bias", kLinearB, d.dim, 0);
    }
    L.ln1_w = add(pre + "sa_layer_norm.weight", kVector, d.dim, 0);
    L.ln1_b = add(pre + "sa_layer_norm.bias", kVector, d.dim, 0);
    L.w[4] = add(pre + "ffn.lin1.weight", kLinearW, d.ffn, d.dim);
    L.b[4] = add(pre + "ffn.lin1.bias", kLinearB, d.ffn, 0);
    L.w[5] = add(pre + "ffn.lin2.weight", kLinearW, d.dim, d.ffn);
    L.b[5] = add(pre + "ffn.lin2.bias", kLinearB, d.dim, 0);
    L.ln2_w = add(pre + "output_layer_norm.weight", kVector, d.dim, 0);
    L.ln2_b = add(pre + "output_layer_norm.bias", kVector, d.dim, 0);
    for (int i = 0; i < 6; ++i) {
      L.wt[i] = 3 * m->lin_weight_floats_;
      m->lin_weight_floats_ += static_cast<size_t>(i == 4 ? d.ffn : d.dim) * (i == 5 ? d.ffn : d.dim);
    }
    m->layers_.push_back(L);
  }
  m->lin_end_ = off;
  m->nparams_ = off;
  *out = m;
  return std::string();
}

size_t DistilBert::workspace_bytes(int max_tokens) const {
  const size_t M = static_cast<size_t>(max_tokens > 0 ? max_tokens : 0);
  size_t b = align_up(3 * lin_weight_floats_ * 4, 1024);  // split weights [hi | hi | lo]
  b += align_up(M * d_.dim * 4, 1024);                    // x
  b += 2 * align_up(M * 3 * d_.dim * 4, 1024);            // xr, ctx (split)
  b += align_up(M * 3 * d_.dim * 4, 1024);                // qkv
  b += align_up(M * d_.dim * 4, 1024);                    // h
  b += align_up(M * d_.ffn * 4, 1024);                    // ff
  b += align_up(M * 3 * d_.ffn * 4, 1024);                // ffs (split)
  b += align_up(static_cast<size_t>(d_.ffn) * 4, 1024);   // ones
  return b;
}

std::string DistilBert::bind(float* params, void* ws, size_t ws_bytes, int max_tokens) {
  if (!params || !ws) return "distilbert: null buffer";
  if (max_tokens < 1) return "distilbert: max_tokens must be positive";
  if (ws_bytes < workspace_bytes(max_tokens)) return "distilbert: workspace too small";
  if (reinterpret_cast<uintptr_t>(ws) % 1024 != 0 || reinterpret_cast<uintptr_t>(params) % 16 != 0)
    return "distilbert: buffers must be aligned (workspace 1024 bytes, parameters 16 bytes)";
  P_ = params;
  max_tokens_ = max_tokens;
  plans_.clear();
  for (auto& kv : graphs_)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
  uint8_t* p = static_cast<uint8_t*>(ws);
  const size_t M = static_cast<size_t>(max_tokens);
  auto take = [&](size_t bytes) {
    float* r = reinterpret_cast<float*>(p);
    p += align_up(bytes, 1024);
    return r;
  };
  Pt_ = take(3 * lin_weight_floats_ * 4);
  x_ = take(M * d_.dim * 4);
  xr_ = take(M * 3 * d_.dim * 4);
  ctx_ = take(M * 3 * d_.dim * 4);
  qkv_ = take(M * 3 * d_.dim * 4);
  h_ = take(M * d_.dim * 4);
  ff_ = take(M * d_.ffn * 4);
  ffs_ = take(M * 3 * d_.ffn * 4);
  ones_ = take(static_cast<size_t>(d_.ffn) * 4);
  return std::string();
}

std::string DistilBert::sync_weights(cudaStream_t stream) {
  if (!P_) return "distilbert: not bound";
  for (const Layer& L : layers_)
    for (int i = 0; i < 6; ++i) {
      const long long rows = (i == 4) ? d_.ffn : d_.dim;
      const int K = (i == 5) ? d_.ffn : d_.dim;
      const long long work = rows * (K / 4);
      const int grid = static_cast<int>(std::min<long long>((work + 255) / 256, 148 * 8));
      launch_kernel(split3_kernel, dim3(grid), dim3(256), 0, stream, static_cast<const float*>(P_ + L.w[i]),
                    Pt_ + L.wt[i], rows, K, 1);
    }
  const int n = d_.ffn;
  launch_kernel(fill_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, ones_, n, 1.0f);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return std::string("distilbert sync_weights: ") + cudaGetErrorString(e);
  return std::string();
}

std::string DistilBert::plan_for(int M, Plans** out) {
  auto it = plans_.find(M);
  if (it != plans_.end()) {
    *out = &it->second;
    return std::string();
  }
  Plans pl;
  std::string err;
  // y[M][Cout] (row pitch ldo, column offset col0) = act(src[M][K] . W[Cout][K]^T + bias [+ res]); the per-channel
  // affine table of the epilogue holds at most 2048 channels per launch, wider outputs go out as column slices
  auto gemm = [&](const float* src, int K3, size_t wt_off, size_t b_off, int Cout, float* dst, int ldo, const float* res,
                  int act) {
    // few token rows (a c3 step encodes 64 sentences of ~12 tokens: six 128-row tiles): 64-wide N tiles put 4x more
    // CTAs on the long split-tf32 reduction axis than 256-wide ones (2.35 -> measured below, tools/bench_bert.py)
    const int bn = (M <= 1536) ? 64 : 0;
    const int max_cols = bn ? 1024 : 2048;
    const int slices = (Cout + max_cols - 1) / max_cols;
    const int per = Cout / slices;
    if (per * slices != Cout || per % 64 != 0) {
      err = "distilbert: cannot slice a Linear of width " + std::to_string(Cout);
      return;
    }
    for (int s = 0; s < slices; ++s) {
      GatherConv g;
      g.src = src;
      g.N = M;
      g.H = g.W = g.P = g.Q = 1;
      g.C = K3;
      g.stride = 1;
      g.ntaps = 1;
      g.wpk = Pt_ + wt_off + static_cast<size_t>(s) * per * K3;
      g.Cout = per;
      g.out = dst + s * per;
      g.ldo = ldo;
      g.ep_scale = ones_;
      g.ep_shift = P_ + b_off + s * per;
      g.ep_res = res ? res + s * per : nullptr;
      g.ep_relu = act;
      g.ep_exact = 1;  // every consumer re-splits (or is fp32 SIMT): keep the fp32 result
      g.tf32 = 1;
      g.bn = bn;
      ConvPlan cp;
      const std::string e2 = plan_conv(g, &cp);
      if (!e2.empty()) {
        err = e2;
        return;
      }
      pl.gemm.push_back(cp);
    }
  };
  const int D3 = 3 * d_.dim, F3 = 3 * d_.ffn;
  for (const Layer& L : layers_) {
    const size_t before = pl.gemm.size();
    for (int i = 0; i < 3; ++i) gemm(xr_, D3, L.wt[i], L.b[i], d_.dim, qkv_ + i * d_.dim, D3, nullptr, 0);
    gemm(ctx_, D3, L.wt[3], L.b[3], d_.dim, h_, d_.dim, x_, 0);    // out_lin + residual -> sa_layer_norm
    gemm(xr_, D3, L.wt[4], L.b[4], d_.ffn, ff_, d_.ffn, nullptr, 2);  // lin1 + GELU
    gemm(ffs_, F3, L.wt[5], L.b[5], d_.dim, h_, d_.dim, x_, 0);   // lin2 + residual -> output_layer_norm
    if (!err.empty()) return err;
    pl.per_layer = static_cast<int>(pl.gemm.size() - before);
  }
  auto ins = plans_.emplace(M, std::move(pl));
  *out = &ins.first->second;
  return std::string();
}

std::string DistilBert::forward(const int* ids, const float* mask, int B, int T, float* out, float* hidden,
                                cudaStream_t stream) {
  if (!P_) return "distilbert: not bound";
  if (!ids || !mask || !out) return "distilbert: null argument";
  if (B < 1 || T < 1) return "distilbert: empty batch";
  if (T > d_.max_pos) return "distilbert: sequence longer than the position table";
  if (B * T > max_tokens_) return "distilbert: more tokens than the workspace was sized for";
  // ~75 launches of a few microseconds each: the call is launch bound.  Like the engine's training step, the sequence is
  // replayed as ONE CUDA graph from the second call with the same buffers and shape on (R3M_STEP_GRAPH=0: plain launches).
  static const bool graphs = !(std::getenv("R3M_STEP_GRAPH") && std::getenv("R3M_STEP_GRAPH")[0] == '0');
  if (!graphs) return enqueue(ids, mask, B, T, out, hidden, stream);
  const std::vector<uint64_t> key = {(uint64_t)reinterpret_cast<uintptr_t>(ids), (uint64_t)reinterpret_cast<uintptr_t>(mask),
                                     (uint64_t)reinterpret_cast<uintptr_t>(out), (uint64_t)reinterpret_cast<uintptr_t>(hidden),
                                     (uint64_t)B, (uint64_t)T};
  auto it = graphs_.find(key);
  if (it == graphs_.end()) {
    if (graphs_.size() >= 8) {  // evict the least recently used entry (buffers that went away)
      auto victim = graphs_.begin();
      for (auto jt = graphs_.begin(); jt != graphs_.end(); ++jt)
        if (jt->second.last_use < victim->second.last_use) victim = jt;
      if (victim->second.exec) cudaGraphExecDestroy(victim->second.exec);
      graphs_.erase(victim);
    }
    it = graphs_.emplace(key, Graph()).first;
  }
  Graph& g = it->second;
  g.last_use = ++graph_clock_;
  if (g.exec == nullptr && !g.failed && g.seen >= 1) {
    if (!cap_ && cudaStreamCreateWithFlags(&cap_, cudaStreamNonBlocking) != cudaSuccess) cap_ = nullptr;
    cudaGraph_t graph = nullptr;
    if (cap_ && cudaStreamBeginCapture(cap_, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
      const std::string cerr = enqueue(ids, mask, B, T, out, hidden, cap_);
      const cudaError_t ee = cudaStreamEndCapture(cap_, &graph);
      if (!cerr.empty() || ee != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
        g.exec = nullptr;
        g.failed = true;
      }
      g.launches = launches_;
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
    } else {
      g.failed = true;
      (void)cudaGetLastError();
    }
  }
  ++g.seen;
  if (g.exec == nullptr) return enqueue(ids, mask, B, T, out, hidden, stream);
  const cudaError_t e = cudaGraphLaunch(g.exec, stream);
  if (e != cudaSuccess) return std::string("distilbert graph launch: ") + cudaGetErrorString(e);
  launches_ = g.launches;
  return std::string();
}

DistilBert::~DistilBert() {
  for (auto& kv : graphs_)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (cap_) cudaStreamDestroy(cap_);
}

std::string DistilBert::enqueue(const int* ids, const float* mask, int B, int T, float* out, float* hidden,
                                cudaStream_t stream) {
  const int M = B * T;
  Plans* pl = nullptr;
  std::string err = plan_for(M, &pl);
  if (!err.empty()) return err;
  int* flag = device_error_flag();
  if (!flag) return "distilbert: no device error flag";
  launches_ = 0;
  const dim3 ln_grid((M + 3) / 4), ln_block(128);
  launch_kernel(layernorm_kernel<true>, ln_grid, ln_block, 0, stream, static_cast<const float*>(nullptr), ids,
                static_cast<const float*>(P_ + word_off_), static_cast<const float*>(P_ + pos_off_),
                static_cast<const float*>(P_ + eln_w_), static_cast<const float*>(P_ + eln_b_), x_, xr_, M, T, d_.dim,
                d_.vocab, flag);
  ++launches_;
  size_t gi = 0;
  const int ff1_slices = pl->per_layer - 5;
  for (const Layer& L : layers_) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 3 && e == cudaSuccess; ++i, ++launches_) e = run_conv(pl->gemm[gi++], stream);
    if (e != cudaSuccess) return std::string("distilbert qkv: ") + cudaGetErrorString(e);
    launch_kernel(attention_kernel, dim3(B * d_.heads), dim3(128), 0, stream, static_cast<const float*>(qkv_), mask, ctx_,
                  T, d_.dim, d_.heads);
    ++launches_;
    e = run_conv(pl->gemm[gi++], stream);
    ++launches_;
    if (e != cudaSuccess) return std::string("distilbert out_lin: ") + cudaGetErrorString(e);
    launch_kernel(layernorm_kernel<false>, ln_grid, ln_block, 0, stream, static_cast<const float*>(h_),
                  static_cast<const int*>(nullptr), static_cast<const float*>(nullptr),
                  static_cast<const float*>(nullptr), static_cast<const float*>(P_ + L.ln1_w),
                  static_cast<const float*>(P_ + L.ln1_b), x_, xr_, M, T, d_.dim, d_.vocab, flag);
    ++launches_;
    for (int i = 0; i < ff1_slices && e == cudaSuccess; ++i, ++launches_) e = run_conv(pl->gemm[gi++], stream);
    if (e != cudaSuccess) return std::string("distilbert ffn.lin1: ") + cudaGetErrorString(e);
    {
      const long long work = static_cast<long long>(M) * (d_.ffn / 4);
      const int grid = static_cast<int>(std::min<long long>((work + 255) / 256, 148 * 8));
      launch_kernel(split3_kernel, dim3(grid), dim3(256), 0, stream, static_cast<const float*>(ff_), ffs_,
                    static_cast<long long>(M), d_.ffn, 0);
      ++launches_;
    }
    e = run_conv(pl->gemm[gi++], stream);
    ++launches_;
    if (e != cudaSuccess) return std::string("distilbert ffn.lin2: ") + cudaGetErrorString(e);
    launch_kernel(layernorm_kernel<false>, ln_grid, ln_block, 0, stream, static_cast<const float*>(h_),
                  static_cast<const int*>(nullptr), static_cast<const float*>(nullptr),
                  static_cast<const float*>(nullptr), static_cast<const float*>(P_ + L.ln2_w),
                  static_cast<const float*>(P_ + L.ln2_b), x_, xr_, M, T, d_.dim, d_.vocab, flag);
    ++launches_;
  }
  launch_kernel(mean_pool_kernel, dim3(B), dim3(256), 0, stream, static_cast<const float*>(x_), out, hidden, T, d_.dim);
  ++launches_;
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return std::string("distilbert forward: ") + cudaGetErrorString(e);
  return std::string();
}

}  // namespace r3m
