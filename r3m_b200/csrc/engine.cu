#include "engine.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "elementwise.cuh"
#include "launch.h"
#include "loss.cuh"

namespace r3m {

// dA + y of a layer at or below this many bytes: BatchNorm backward runs as one fused launch (see elementwise.cuh)
constexpr double kFuseBnBwdBytes = 72e6;

typedef __nv_bfloat16 bf16;

// Layers with at most this many output rows (N * P * Q) have their BatchNorm reductions finalised by the producers (see
// Conv::fin).  R3M_BN_FIN_ROWS overrides it (0: every layer converts on read).
static long long bn_finalize_max_rows() {
  static const long long rows = std::getenv("R3M_BN_FIN_ROWS") ? atoll(std::getenv("R3M_BN_FIN_ROWS")) : 260000;
  return rows;
}

struct Engine::Conv {
  std::string name, bn;
  int Cin = 0, Cout = 0, R = 1, stride = 1, pad = 0, H = 0, W = 0, P = 0, Q = 0;
  bool stem = false;
  size_t w_off = 0, gamma_off = 0, beta_off = 0;  // flat parameter buffer (elements)
  size_t rm_off = 0, rv_off = 0;                  // BN buffers region (floats)
  // per-step zeroed region (floats): fixed-point accumulators (fx_add, ptx.cuh) of the layer's BatchNorm reductions —
  // [0, 16C) forward statistics (sum, sum of squares: 4 64-bit words each per channel), [16C, 32C) backward sums (2C
  // entries of 4 words), [32C, 40C) the downsample branch's extra backward sum (C entries)
  // Producer-finalised layers (`fin`, see Engine::create) use [40C + 8, 42C + 8) as fp32 batch moments (mean[C], var[C]) and
  // the 64 ints behind them as the per-column-block tickets of the conv kernel's last-CTA conversion.
  size_t zero_off = 0;
  bool fin = false;
  size_t fin_off(int) const { return zero_off + 40 * (size_t)Cout + 8; }
  size_t save_off = 0;                            // saved batch statistics (floats): mean[C] rstd[C]
  size_t wd_off = 0;                              // dgrad-packed filters (bf16 elements)
  size_t y_off = 0, a_off = 0;                    // arena byte offsets of the raw / activated outputs
  size_t mask_off = 0;                            // bit-packed ReLU mask of the activated output (train mode)
  uint8_t* mask = nullptr;
  const bf16* x = nullptr;                        // input activation
  bf16* y = nullptr;
  bf16* a = nullptr;
  size_t out_elems(int N) const { return (size_t)N * P * Q * Cout; }
};

struct Engine::Block {
  std::vector<int> main;  // conv indices of the residual branch, in order
  int ds = -1;            // downsample conv index or -1
  const bf16* x_in = nullptr;
  bf16* a_out = nullptr;
  int Hin = 0, Win = 0, Cin = 0;
  bf16* d_out = nullptr;  // backward: gradient w.r.t. the block's output (read) ...
  bf16* d_in = nullptr;   // ... and w.r.t. its input (written); the two swap from block to block
};

namespace {
constexpr size_t kAlign = 1024;
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr size_t kMaxActPerFrame = 112 * 112 * 64;  // largest activation (stem output == layer1 bottleneck output)
constexpr size_t kMaxPackEntries = 128;  // dgrad re-pack table: one entry per (conv, output-parity class)
constexpr int kGraphMaxFrames = 16;  // eval forwards up to this many frames are launch-latency bound -> CUDA graph
}  // namespace

Engine::~Engine() {
  if (eval_graph_) cudaGraphExecDestroy(eval_graph_);
  for (auto& kv : graphs_)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (cudaEvent_t ev : ext_evs_)
    if (ev) cudaEventDestroy(ev);
  if (cap_) cudaStreamDestroy(cap_);
  for (cudaEvent_t ev : evs_) cudaEventDestroy(ev);
  if (side_) cudaStreamDestroy(side_);
  for (Conv* c : convs_) delete c;
  for (Block* b : blocks_) delete b;
}

void* Engine::region(int which) const {
  if (!bound_) return nullptr;
  switch (which) {
    case 0: return pws_ + off_P_;
    case 1: return pws_ + off_G_;
    case 2: return pws_ + off_M_;
    case 3: return pws_ + off_V_;
    case 4: return pws_ + off_buf_;
    case 5: return ws_ + off_E_;
    case 6: return ws_ + off_dE_;
    case 7: return ws_ + off_metrics_;
    default: return nullptr;
  }
}

std::string Engine::create(int size, int frames, int lang_head, int hidden_dim, Engine** out) {
  if (size != 18 && size != 34 && size != 50) return "size must be 18, 34 or 50 (r3m/models/models_r3m.py:44-52)";
  if (frames < 1) return "frames must be positive";
  if (lang_head && (hidden_dim < 64 || hidden_dim % 64 != 0)) return "hidden_dim must be a positive multiple of 64";
  Engine* e = new Engine();
  e->size_ = size;
  e->N_ = frames;
  e->B_ = (frames % 5 == 0) ? frames / 5 : 0;
  e->lang_ = lang_head;
  e->hidden_ = hidden_dim;
  e->bottleneck_ = (size == 50);
  e->D_ = e->bottleneck_ ? 2048 : 512;

  // ---- architecture (tv resnet.py:166-262; layer counts :705,:731,:763) -------------------------------------
  const int layers18[4] = {2, 2, 2, 2}, layers34[4] = {3, 4, 6, 3};
  const int* layers = (size == 18) ? layers18 : layers34;
  auto add_conv = [&](const std::string& name, const std::string& bn, int cin, int cout, int r, int stride, int pad,
                      int h, int w) {
    Conv* c = new Conv();
    c->name = name;
    c->bn = bn;
    c->Cin = cin;
    c->Cout = cout;
    c->R = r;
    c->stride = stride;
    c->pad = pad;
    c->H = h;
    c->W = w;
    c->P = (h + 2 * pad - r) / stride + 1;
    c->Q = (w + 2 * pad - r) / stride + 1;
    e->convs_.push_back(c);
    return (int)e->convs_.size() - 1;
  };
  const int stem = add_conv("conv1", "bn1", 3, 64, 7, 2, 3, 224, 224);
  e->convs_[stem]->stem = true;
  int inplanes = 64, hw = 56;
  const int expansion = e->bottleneck_ ? 4 : 1;
  const int planes_l[4] = {64, 128, 256, 512};
  for (int li = 0; li < 4; ++li) {
    const int planes = planes_l[li];
    for (int b = 0; b < layers[li]; ++b) {
      const int stride = (li > 0 && b == 0) ? 2 : 1;
      const std::string pre = "layer" + std::to_string(li + 1) + "." + std::to_string(b);
      Block* blk = new Block();
      blk->Hin = hw;
      blk->Win = hw;
      blk->Cin = inplanes;
      const int hout = hw / stride;
      if (!e->bottleneck_) {
        blk->main.push_back(add_conv(pre + ".conv1", pre + ".bn1", inplanes, planes, 3, stride, 1, hw, hw));
        blk->main.push_back(add_conv(pre + ".conv2", pre + ".bn2", planes, planes, 3, 1, 1, hout, hout));
      } else {
        blk->main.push_back(add_conv(pre + ".conv1", pre + ".bn1", inplanes, planes, 1, 1, 0, hw, hw));
        blk->main.push_back(add_conv(pre + ".conv2", pre + ".bn2", planes, planes, 3, stride, 1, hw, hw));
        blk->main.push_back(add_conv(pre + ".conv3", pre + ".bn3", planes, planes * 4, 1, 1, 0, hout, hout));
      }
      if (b == 0 && (stride != 1 || inplanes != planes * expansion))
        blk->ds = add_conv(pre + ".downsample.0", pre + ".downsample.1", inplanes, planes * expansion, 1, stride, 0, hw,
                           hw);
      inplanes = planes * expansion;
      hw = hout;
      e->blocks_.push_back(blk);
    }
  }

  // ---- flat parameter / buffer layout -------------------------------------------------------------------------
  size_t np = 0, nb = 0, nz = 0, ns = 0, nwd = 0;
  auto take = [](size_t& cursor, size_t n, size_t align) {
    cursor = align_up(cursor, align);
    const size_t o = cursor;
    cursor += n;
    return o;
  };
  for (Conv* c : e->convs_) {
    const size_t wn = (size_t)c->Cout * c->R * c->R * c->Cin;
    c->w_off = take(np, wn, 128);
    c->gamma_off = take(np, c->Cout, 128);
    c->beta_off = take(np, c->Cout, 128);
    c->rm_off = take(nb, c->Cout, 32);
    c->rv_off = take(nb, c->Cout, 32);
    c->zero_off = take(nz, 42 * (size_t)c->Cout + 8 + 64, 32);  // + the grid-barrier counter of the fused BatchNorm
                                                                // backward, + fp32 statistics and tickets (Conv::fin)
    // Small-M layers: the producers finalise the reductions (last-CTA conversion of the fixed-point accumulators to
    // fp32: conv statistics per column block, bn_bwd_reduce per channel slice) instead of every block of the consuming
    // BatchNorm kernel converting all C channels in its prologue — with few rows per block that prologue moves more
    // bytes through L2 than the payload (ncu: layer4.bn3 apply 541 MB of L2 traffic for 197 MB of tensors).
    c->fin = !c->stem && (long long)frames * c->P * c->Q <= bn_finalize_max_rows();
    c->save_off = take(ns, 2 * (size_t)c->Cout, 32);
    if (!c->stem) c->wd_off = take(nwd, wn, 128);
    TensorInfo t;
    t.name = "convnet." + c->name + ".weight";
    t.kind = c->stem ? kStemOIHW : kConvKRSC;
    t.offset = c->w_off;
    t.ndim = 4;
    t.dims[0] = c->Cout;
    t.dims[1] = c->Cin;
    t.dims[2] = c->R;
    t.dims[3] = c->R;
    e->tensors_.push_back(t);
    const char* suffix[4] = {".weight", ".bias", ".running_mean", ".running_var"};
    const int kinds[4] = {kVector, kVector, kRunMean, kRunVar};
    const size_t offs[4] = {c->gamma_off, c->beta_off, c->rm_off, c->rv_off};
    for (int i = 0; i < 4; ++i) {
      TensorInfo v;
      v.name = "convnet." + c->bn + suffix[i];
      v.kind = kinds[i];
      v.offset = offs[i];
      v.ndim = 1;
      v.dims[0] = c->Cout;
      e->tensors_.push_back(v);
    }
  }
  if (lang_head) {
    // LanguageReward.pred (models_language.py:43-51): Linear(2D+768,H) ReLU (Linear(H,H) ReLU)x3 Linear(H,1)
    const int dims[6] = {2 * e->D_ + 768, hidden_dim, hidden_dim, hidden_dim, hidden_dim, 1};
    for (int l = 0; l < 5; ++l) {
      e->lang_w_off_[l] = take(np, (size_t)dims[l + 1] * dims[l], 128);
      e->lang_b_off_[l] = take(np, (size_t)dims[l + 1], 128);
      TensorInfo w;
      w.name = "lang_rew.pred." + std::to_string(2 * l) + ".weight";
      w.kind = kLinearW;
      w.offset = e->lang_w_off_[l];
      w.ndim = 2;
      w.dims[0] = dims[l + 1];
      w.dims[1] = dims[l];
      e->tensors_.push_back(w);
      TensorInfo b;
      b.name = "lang_rew.pred." + std::to_string(2 * l) + ".bias";
      b.kind = kLinearB;
      b.offset = e->lang_b_off_[l];
      b.ndim = 1;
      b.dims[0] = dims[l + 1];
      e->tensors_.push_back(b);
    }
    e->lang_dims_.B = e->B_;
    e->lang_dims_.D = e->D_;
    e->lang_dims_.L = 768;
    e->lang_dims_.H = hidden_dim;
  }
  np = align_up(np, 128);
  e->nparams_ = np;
  e->nbuf_ = align_up(nb, 32);
  e->nsaved_ = align_up(ns, 32);
  e->nwd_ = align_up(nwd, 128);

  // ---- two arenas: the parameter block (shared by every engine of one model) and the activation block ----------
  size_t cur = 0;
  auto arena = [&](size_t bytes) { return take(cur, bytes, kAlign); };
  const size_t N = (size_t)frames;
  e->off_P_ = arena(np * 4);
  e->off_G_ = arena(np * 4);
  e->off_M_ = arena(np * 4);
  e->off_V_ = arena(np * 4);
  e->off_Pb_ = arena(np * 2);
  e->off_buf_ = arena(e->nbuf_ * 4);
  e->off_wd_ = arena(e->nwd_ * 2);
  e->off_stem_wp_ = arena(64 * 4 * 64 * 2);
  e->pws_bytes_ = align_up(cur, kAlign);

  cur = 0;
  e->off_saved_ = arena(e->nsaved_ * 4);
  // region cleared at the start of every step: BN statistics, BN-backward sums, metrics, stem filter gradient
  const size_t zero_floats = align_up(nz, 32) + kNumMetrics + 64 * 4 * 64;
  e->off_zero_ = arena(zero_floats * 4);
  e->zero_bytes_ = zero_floats * 4;
  e->off_metrics_ = e->off_zero_ + align_up(nz, 32) * 4;
  e->off_stem_dwp_ = e->off_metrics_ + kNumMetrics * 4;
  e->off_xs_ = arena(N * 112 * 112 * 64 * 2);
  e->off_argmax_ = arena(N * 56 * 56 * 64);
  e->off_ymax_ = arena(N * 56 * 56 * 64 * 2);
  for (Conv* c : e->convs_) {
    c->y_off = arena(c->out_elems(frames) * 2);
    if (c->stem)
      c->a_off = arena(N * 56 * 56 * 64 * 2);  // pooled
    else
      c->a_off = arena(c->out_elems(frames) * 2);
    if (!c->stem) c->mask_off = arena(c->out_elems(frames) / 8);
  }
  {
    // The tf32 inference tier (fwd_eval_tf32_) aliases the activation region [off_xs_, off_E_): an eval forward and a
    // training step's saved activations are mutually exclusive anyway.  Pad the region to what that tier needs:
    // the fp32 stem operand, five ping-pong activation buffers, the tf32-rounded copy of the parameters, the fp32 stem
    // filter and its own embedding buffer.
    size_t need = 0;
    auto sub = [&](size_t bytes) {
      const size_t o = align_up(need, kAlign);
      need = o + bytes;
      return o;
    };
    e->t32_xs_ = sub(N * 112 * 112 * 64 * 4);
    for (int i = 0; i < 5; ++i) e->t32_buf_[i] = sub(N * kMaxActPerFrame * 4);
    e->t32_params_ = sub(np * 4);
    e->t32_stem_w_ = sub(64 * 4 * 64 * 4);
    e->t32_E_ = sub(N * e->D_ * 4);
    const size_t have = align_up(cur, kAlign) - e->off_xs_;
    if (need > have) arena(need - have);
  }
  e->off_E_ = arena(N * e->D_ * 4);
  e->off_dE_ = arena(N * e->D_ * 4);
  for (int i = 0; i < 7; ++i) e->off_g_[i] = arena(N * kMaxActPerFrame * 2);
  // a model with a language head still embeds any number of frames (R3M.forward); only update() needs 5 * clips
  if (lang_head && e->B_ > 0) e->off_lang_ws_ = arena(lang_workspace_floats(e->lang_dims_) * 4);
  if (lang_head && e->B_ > 0) e->off_lang_tc_ = arena(lang_tc_floats(e->lang_dims_) * 4);
  e->off_fold_ = arena(2 * e->convs_.size() * sizeof(BnFoldEntry));  // bf16 tier (stem excluded) | tf32 tier (all)
  e->off_pack_ = arena(kMaxPackEntries * sizeof(PackDgradEntry));
  if (frames <= kGraphMaxFrames) e->off_obs_stage_ = arena(N * 3 * 224 * 224 * 4);
  // scratch of the deterministic reductions (zeroed once at bind: the tickets reset themselves):
  //   conv statistics [kStatScratchFloats floats | kStatTickets ints], BatchNorm-backward / stem ordered reduce
  //   [kDetScratchFloats floats | kDetTickets ints], loss-head partials, two wgrad split-K buffers (side / main stream)
  e->off_det_ = arena(kStatScratchFloats * 4 + kStatTickets * 4);
  e->off_det_bn_ = arena(kDetScratchFloats * 4 + kDetTickets * 4);
  e->off_det_loss_ = arena(((size_t)N * 4 + (size_t)(e->B_ + 1) * 64) * 4);
  e->det_small_bytes_ = align_up(cur, kAlign) - e->off_det_;
  e->off_wgrad_scratch_[0] = arena(kWgradScratchBytes);
  e->off_wgrad_scratch_[1] = arena(kWgradScratchBytes);
  e->ws_bytes_ = align_up(cur, kAlign);
  *out = e;
  return std::string();
}

std::string Engine::bind(void* params, size_t param_bytes, void* ws, size_t bytes, cudaStream_t stream) {
  if (param_bytes < pws_bytes_) return "parameter block too small";
  if (bytes < ws_bytes_) return "workspace too small";
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0 || (reinterpret_cast<uintptr_t>(params) & 1023) != 0)
    return "parameter block and workspace must be 1024-byte aligned";
  pws_ = reinterpret_cast<uint8_t*>(params);
  ws_ = reinterpret_cast<uint8_t*>(ws);
  cudaError_t e = cudaMemsetAsync(ws_ + off_saved_, 0, nsaved_ * 4 + 0, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws_ + off_zero_, 0, zero_bytes_, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws_ + off_det_, 0, det_small_bytes_, stream);
  if (e != cudaSuccess) return std::string("memset: ") + cudaGetErrorString(e);
  for (Conv* c : convs_) {
    c->y = reinterpret_cast<bf16*>(ws_ + c->y_off);
    c->a = reinterpret_cast<bf16*>(ws_ + c->a_off);
    c->mask = c->stem ? nullptr : ws_ + c->mask_off;
  }
  bound_ = true;
  return plan_all();
}

void Engine::add_bn_apply(std::vector<Op>& ops, const Conv& c, const void* residual, void* dst, int relu, int train,
                          const Conv* second) {
  float* P = reinterpret_cast<float*>(pws_ + off_P_);
  float* buf = reinterpret_cast<float*>(pws_ + off_buf_);
  float* zero = reinterpret_cast<float*>(ws_ + off_zero_);
  float* saved = reinterpret_cast<float*>(ws_ + off_saved_);
  BnApplyArgs a;
  a.y = c.y;
  a.a = dst;
  a.residual = residual;
  a.M = N_ * c.P * c.Q;
  a.C = c.Cout;
  a.relu = relu;
  a.train = train;
  const bool fin = c.fin && (!second || second->fin);
  a.sum = fin ? zero + c.fin_off(0) : zero + c.zero_off;
  a.sq = fin ? a.sum + c.Cout : a.sum;
  a.stat_raw = fin ? 3 : bn_stat_mode();
  a.gamma = P + c.gamma_off;
  a.beta = P + c.beta_off;
  a.running_mean = buf + c.rm_off;
  a.running_var = buf + c.rv_off;
  a.save_mean = saved + c.save_off;
  a.save_rstd = saved + c.save_off + c.Cout;
  if (train && relu) a.mask_out = c.mask;
  // the conv that has just written y walked the rows first-to-last, and the conv that reads `dst` next does too
  // (R3M_L2_ORDER_MIN_MB restricts it to tensors above a size; the per-launch profile suggests the small wide layers lose,
  // the step says otherwise: with every tensor reversed it is 0.1-0.3 ms shorter than with 96 / 200 / 400 MB thresholds)
  a.reverse = (train && l2_order_ && (double)a.M * a.C * 2 >= l2_order_min_bytes_) ? 1 : 0;
  if (second) {
    // downsample branch folded in: a = relu(bn(y) + bn_ds(y_ds)), bn_ds(y_ds) is never materialised
    const Conv& d = *second;
    a.y2 = d.y;
    a.sum2 = fin ? zero + d.fin_off(0) : zero + d.zero_off;
    a.sq2 = fin ? a.sum2 + d.Cout : a.sum2;
    a.gamma2 = P + d.gamma_off;
    a.beta2 = P + d.beta_off;
    a.running_mean2 = buf + d.rm_off;
    a.running_var2 = buf + d.rv_off;
    a.save_mean2 = saved + d.save_off;
    a.save_rstd2 = saved + d.save_off + d.Cout;
  }
  const double mc = (double)a.M * a.C * 2;
  ops.push_back(Op([a](cudaStream_t s) { return launch_bn_apply(a, s); }, kFamNorm, 0.0,
                   mc * (2 + (residual ? 1 : 0) + (second ? 1 : 0)) + (a.mask_out ? mc / 16 : 0.0)));
  ops.back().label = "bn_apply " + c.bn + (residual ? " +res" : "") + (second ? " +" + second->bn : "");
}

std::string Engine::plan_all() {
  {
    const char* env = std::getenv("R3M_L2_ORDER");
    l2_order_ = !(env && env[0] == '0');
    env = std::getenv("R3M_L2_ORDER_MIN_MB");
    l2_order_min_bytes_ = (env ? atof(env) : 0.0) * 1e6;  // measured: 0 (every tensor) is best; 96 / 200 / 400 MB cost 0.1-0.3 ms
    // off by default: measured +0.1 ms per ResNet-50 step on B200 (the isolated kernels gain 0.2 ms, but the cooperative
    // launch cannot overlap its neighbours' prologues and the filter gradients lose their slot between the two passes)
    env = std::getenv("R3M_LANG_TC");
    lang_tc_ = LangTc();
    if (lang_ && B_ > 0 && !(env && env[0] == '0')) {
      // tensor-core path of the language head's hidden layers (R3M_LANG_TC=0: fp32 SIMT everywhere)
      float* Pp = reinterpret_cast<float*>(pws_ + off_P_);
      float* Gp = reinterpret_cast<float*>(pws_ + off_G_);
      LangParams lp;
      for (int l = 0; l < 5; ++l) {
        lp.w[l] = Pp + lang_w_off_[l];
        lp.b[l] = Pp + lang_b_off_[l];
        lp.dw[l] = Gp + lang_w_off_[l];
        lp.db[l] = Gp + lang_b_off_[l];
      }
      LangWorkspace lw;
      lang_carve_workspace(reinterpret_cast<float*>(ws_ + off_lang_ws_), lang_dims_, &lw);
      const std::string terr = lang_tc_plan(lang_dims_, lp, lw, reinterpret_cast<float*>(ws_ + off_lang_tc_), &lang_tc_);
      if (!terr.empty()) return terr;
    }
    env = std::getenv("R3M_FUSE_BN_BWD");
    fuse_bn_bwd_ = (env && env[0] == '1') && std::getenv("R3M_GRID_WAVES") == nullptr;
  }
  float* P = reinterpret_cast<float*>(pws_ + off_P_);
  float* G = reinterpret_cast<float*>(pws_ + off_G_);
  bf16* Pb = reinterpret_cast<bf16*>(pws_ + off_Pb_);
  float* buf = reinterpret_cast<float*>(pws_ + off_buf_);
  float* zero = reinterpret_cast<float*>(ws_ + off_zero_);
  float* saved = reinterpret_cast<float*>(ws_ + off_saved_);
  bf16* wd = reinterpret_cast<bf16*>(pws_ + off_wd_);
  bf16* stem_wp = reinterpret_cast<bf16*>(pws_ + off_stem_wp_);
  float* stem_dwp = reinterpret_cast<float*>(ws_ + off_stem_dwp_);
  bf16* xs = reinterpret_cast<bf16*>(ws_ + off_xs_);
  uint8_t* argmax = ws_ + off_argmax_;
  bf16* ymax = reinterpret_cast<bf16*>(ws_ + off_ymax_);
  float* E = reinterpret_cast<float*>(ws_ + off_E_);
  bf16* g[7];
  for (int i = 0; i < 7; ++i) g[i] = reinterpret_cast<bf16*>(ws_ + off_g_[i]);
  const int N = N_;
  std::string err;

  auto push_conv = [&](std::vector<Op>& ops, const GatherConv& gc, const std::string& label) {
    ConvPlan plan;
    std::string e2 = plan_conv(gc, &plan);
    if (!e2.empty()) {
      err = e2;
      return;
    }
    // algorithmic cost: the stem is charged for its real 7x7x3 taps, not the zero-padded 4x64 operand
    const double kdim = (gc.src == xs) ? 147.0 : (double)gc.ntaps * gc.C;
    const double m = (double)gc.N * gc.P * gc.Q;
    ops.push_back(Op([plan](cudaStream_t s) { return run_conv(plan, s); }, kFamConv, 2.0 * m * gc.Cout * kdim,
                     2.0 * ((double)gc.N * gc.H * gc.W * gc.C + m * gc.Cout * (gc.accumulate ? 2 : 1) + gc.Cout * kdim)));
    ops.back().label = label;
  };
  auto fwd_geom = [&](const Conv& c, int train) {
    GatherConv gc;
    gc.N = N;
    gc.out = c.y;
    gc.Cout = c.Cout;
    gc.ldo = c.Cout;
    if (c.stem) {
      // 7x7 s2 p3 over 3 channels == 4 vertical taps over the 64-wide space-to-depth operand (elementwise.cuh)
      gc.src = xs;
      gc.H = 112;
      gc.W = 112;
      gc.C = 64;
      gc.P = 112;
      gc.Q = 112;
      gc.stride = 1;
      gc.base_h = -2;
      gc.base_w = 0;
      gc.ntaps = 4;
      for (int t = 0; t < 4; ++t) {
        gc.tap_h[t] = t;
        gc.tap_w[t] = 0;
      }
      gc.wpk = stem_wp;
    } else {
      gc.src = c.x;
      gc.H = c.H;
      gc.W = c.W;
      gc.C = c.Cin;
      fill_fwd_geometry(&gc, c.R, c.R, c.stride, c.pad);
      gc.wpk = Pb + c.w_off;
    }
    if (train && c.fin) {
      gc.stat_sum = zero + c.fin_off(0);  // (mean, variance) published by the last CTA of every column block
      gc.stat_sq = gc.stat_sum + c.Cout;
      gc.stat_rows = N * c.P * c.Q;
      gc.stat_scratch = zero + c.zero_off;
      gc.stat_ticket = reinterpret_cast<int*>(zero + c.fin_off(0) + 2 * (size_t)c.Cout);
      gc.stat_raw = 0;
    } else if (train) {
      gc.stat_sum = zero + c.zero_off;  // the layer's raw accumulators: the BatchNorm kernels convert on read
      gc.stat_sq = gc.stat_sum;
      gc.stat_raw = 1;
    }
    return gc;
  };

  // ------------------------------------------------------------------------------------------------ forward
  for (int train = 0; train < 2; ++train) {
    std::vector<Op>& ops = train ? fwd_train_ : fwd_eval_;
    ops.clear();
    Conv& st = *convs_[0];
    push_conv(ops, fwd_geom(st, train), "fwd conv1 (stem 7x7 as 4x64)");
    {
      StemPoolArgs a;
      a.y = st.y;
      a.a = st.a;
      a.argmax = train ? argmax : nullptr;
      a.ymax = train ? ymax : nullptr;
      a.N = N;
      a.train = train;
      a.sum = zero + st.zero_off;
      a.sq = a.sum;
      a.stat_raw = bn_stat_mode();
      a.gamma = P + st.gamma_off;
      a.beta = P + st.beta_off;
      a.running_mean = buf + st.rm_off;
      a.running_var = buf + st.rv_off;
      a.save_mean = saved + st.save_off;
      a.save_rstd = saved + st.save_off + 64;
      ops.push_back(Op([a](cudaStream_t s) { return launch_stem_bn_relu_maxpool(a, s); }, kFamPool, 0.0,
                       (double)N * 64 * (112.0 * 112 * 2 + 56.0 * 56 * (train ? 5 : 2))));
    }
    if (!train) {
      // inference: every BatchNorm except the stem's is folded into its conv's epilogue (scale/shift live in the
      // `saved` slots, which only training uses otherwise); one launch refreshes all of them from the running stats
      std::vector<BnFoldEntry> table;
      for (size_t i = 1; i < convs_.size(); ++i) {
        const Conv& c = *convs_[i];
        BnFoldEntry fe;
        fe.gamma = P + c.gamma_off;
        fe.beta = P + c.beta_off;
        fe.running_mean = buf + c.rm_off;
        fe.running_var = buf + c.rv_off;
        fe.scale = saved + c.save_off;
        fe.shift = saved + c.save_off + c.Cout;
        fe.C = c.Cout;
        table.push_back(fe);
      }
      BnFoldEntry* table_dev = reinterpret_cast<BnFoldEntry*>(ws_ + off_fold_);
      cudaError_t ce = cudaMemcpy(table_dev, table.data(), table.size() * sizeof(BnFoldEntry), cudaMemcpyHostToDevice);
      if (ce != cudaSuccess) return std::string("fold table upload: ") + cudaGetErrorString(ce);
      const int entries = (int)table.size();
      ops.push_back(Op([table_dev, entries](cudaStream_t s) { return launch_bn_fold(table_dev, entries, s); }, kFamNorm));
      ops.back().label = "bn_fold (all layers)";
    }
    const bf16* x = st.a;
    for (Block* blk : blocks_) {
      blk->x_in = x;
      const bf16* cur = x;
      for (size_t i = 0; i < blk->main.size(); ++i) {
        Conv& c = *convs_[blk->main[i]];
        c.x = cur;
        if (train) {
          push_conv(ops, fwd_geom(c, train), "fwd " + c.name);
          if (i + 1 < blk->main.size()) {
            add_bn_apply(ops, c, nullptr, c.a, 1, train, nullptr);
            cur = c.a;
          }
        } else if (i + 1 < blk->main.size()) {
          GatherConv gc = fwd_geom(c, 0);
          gc.out = c.a;
          gc.ep_scale = saved + c.save_off;
          gc.ep_shift = saved + c.save_off + c.Cout;
          gc.ep_relu = 1;
          push_conv(ops, gc, "fwd+bn+relu " + c.name);
          cur = c.a;
        }
      }
      Conv& last = *convs_[blk->main.back()];
      if (train) {
        if (blk->ds >= 0) {
          Conv& d = *convs_[blk->ds];
          d.x = blk->x_in;
          push_conv(ops, fwd_geom(d, train), "fwd " + d.name);
          add_bn_apply(ops, last, nullptr, last.a, 1, train, &d);
        } else {
          add_bn_apply(ops, last, blk->x_in, last.a, 1, train, nullptr);
        }
      } else {
        const void* residual = blk->x_in;
        if (blk->ds >= 0) {
          Conv& d = *convs_[blk->ds];
          d.x = blk->x_in;
          GatherConv gd = fwd_geom(d, 0);
          gd.out = d.a;
          gd.ep_scale = saved + d.save_off;
          gd.ep_shift = saved + d.save_off + d.Cout;
          gd.ep_relu = 0;
          push_conv(ops, gd, "fwd+bn " + d.name);
          residual = d.a;
        }
        GatherConv gl = fwd_geom(last, 0);
        gl.out = last.a;
        gl.ep_scale = saved + last.save_off;
        gl.ep_shift = saved + last.save_off + last.Cout;
        gl.ep_res = residual;
        gl.ep_relu = 1;
        push_conv(ops, gl, "fwd+bn+res+relu " + last.name);
      }
      blk->a_out = last.a;
      x = last.a;
    }
    {
      const bf16* a_last = x;
      const int HW = 49, C = D_;
      ops.push_back(Op([a_last, E, N, HW, C](cudaStream_t s) { return launch_avgpool_fwd(a_last, E, N, HW, C, s); },
                       kFamPool, 0.0, (double)N * C * (HW * 2 + 4)));
    }
    if (!err.empty()) return err;
  }

  // ------------------------------------------------------------------------------------------------ tf32 inference tier
  {
    std::vector<Op>& ops = fwd_eval_tf32_;
    ops.clear();
    uint8_t* base = ws_ + off_xs_;
    float* xs32 = reinterpret_cast<float*>(base + t32_xs_);
    float* pool[5];
    for (int i = 0; i < 5; ++i) pool[i] = reinterpret_cast<float*>(base + t32_buf_[i]);
    float* Pt = reinterpret_cast<float*>(base + t32_params_);
    float* stem_w32 = reinterpret_cast<float*>(base + t32_stem_w_);
    float* E32 = reinterpret_cast<float*>(base + t32_E_);
    const size_t np = nparams_;
    // per-call refresh of the derived operands (no dirty tracking: 0.2 GB of traffic, ~0.03 ms)
    ops.push_back(Op([P, Pt, np](cudaStream_t s) { return launch_round_tf32(P, Pt, np, s); }, kFamOptim, 0.0, 8.0 * np));
    ops.back().label = "round_tf32 (all parameters)";
    {
      const float* w = P + convs_[0]->w_off;
      ops.push_back(Op([w, stem_w32](cudaStream_t s) { return launch_stem_pack_f32(w, stem_w32, s); }, kFamOptim));
    }
    {
      std::vector<BnFoldEntry> table;
      for (Conv* cp : convs_) {
        const Conv& c = *cp;
        BnFoldEntry fe;
        fe.gamma = P + c.gamma_off;
        fe.beta = P + c.beta_off;
        fe.running_mean = buf + c.rm_off;
        fe.running_var = buf + c.rv_off;
        fe.scale = saved + c.save_off;
        fe.shift = saved + c.save_off + c.Cout;
        fe.C = c.Cout;
        table.push_back(fe);
      }
      BnFoldEntry* table_dev = reinterpret_cast<BnFoldEntry*>(ws_ + off_fold_) + convs_.size();
      cudaError_t ce = cudaMemcpy(table_dev, table.data(), table.size() * sizeof(BnFoldEntry), cudaMemcpyHostToDevice);
      if (ce != cudaSuccess) return std::string("fold table upload: ") + cudaGetErrorString(ce);
      const int entries = (int)table.size();
      ops.push_back(Op([table_dev, entries](cudaStream_t s) { return launch_bn_fold(table_dev, entries, s); }, kFamNorm));
      ops.back().label = "bn_fold (all layers, tf32 tier)";
    }
    auto conv32 = [&](const Conv& c, const float* src, float* dst, const float* residual, int relu,
                      const std::string& label) {
      GatherConv gc;
      gc.N = N;
      gc.out = dst;
      gc.Cout = c.Cout;
      gc.ldo = c.Cout;
      gc.src = src;
      if (c.stem) {
        gc.H = 112;
        gc.W = 112;
        gc.C = 64;
        gc.P = 112;
        gc.Q = 112;
        gc.stride = 1;
        gc.base_h = -2;
        gc.base_w = 0;
        gc.ntaps = 4;
        for (int t = 0; t < 4; ++t) {
          gc.tap_h[t] = t;
          gc.tap_w[t] = 0;
        }
        gc.wpk = stem_w32;
      } else {
        gc.H = c.H;
        gc.W = c.W;
        gc.C = c.Cin;
        fill_fwd_geometry(&gc, c.R, c.R, c.stride, c.pad);
        gc.wpk = Pt + c.w_off;
      }
      gc.ep_scale = saved + c.save_off;
      gc.ep_shift = saved + c.save_off + c.Cout;
      gc.ep_res = residual;
      gc.ep_relu = relu;
      gc.tf32 = 1;
      ConvPlan plan;
      std::string e2 = plan_conv(gc, &plan);
      if (!e2.empty()) {
        err = e2;
        return;
      }
      const double kdim = c.stem ? 147.0 : (double)gc.ntaps * gc.C;
      const double m = (double)gc.N * gc.P * gc.Q;
      ops.push_back(Op([plan](cudaStream_t s) { return run_conv(plan, s); }, kFamConv, 2.0 * m * gc.Cout * kdim,
                       4.0 * ((double)gc.N * gc.H * gc.W * gc.C + m * gc.Cout + gc.Cout * kdim)));
      ops.back().label = label;
    };
    Conv& st = *convs_[0];
    conv32(st, xs32, pool[0], nullptr, 1, "tf32 fwd+bn+relu conv1 (stem)");
    {
      const float* y = pool[0];
      float* a = pool[1];
      ops.push_back(Op([y, a, N](cudaStream_t s) { return launch_maxpool_f32(y, a, N, 112, 112, 64, s); }, kFamPool, 0.0,
                       (double)N * 64 * 4 * (112.0 * 112 + 56.0 * 56)));
    }
    int cur = 1;
    for (Block* blk : blocks_) {
      int free_i[4], nf = 0;
      for (int i = 0; i < 5; ++i)
        if (i != cur) free_i[nf++] = i;
      const float* x = pool[cur];
      const float* in = x;
      for (size_t i = 0; i + 1 < blk->main.size(); ++i) {
        const Conv& c = *convs_[blk->main[i]];
        conv32(c, in, pool[free_i[i]], nullptr, 1, "tf32 fwd+bn+relu " + c.name);
        in = pool[free_i[i]];
      }
      const float* residual = x;
      if (blk->ds >= 0) {
        const Conv& d = *convs_[blk->ds];
        conv32(d, x, pool[free_i[2]], nullptr, 0, "tf32 fwd+bn " + d.name);
        residual = pool[free_i[2]];
      }
      const Conv& last = *convs_[blk->main.back()];
      conv32(last, in, pool[free_i[3]], residual, 1, "tf32 fwd+bn+res+relu " + last.name);
      cur = free_i[3];
      if (!err.empty()) return err;
    }
    {
      const float* a_last = pool[cur];
      const int C = D_;
      ops.push_back(Op([a_last, E32, N, C](cudaStream_t s) { return launch_avgpool_fwd_f32(a_last, E32, N, 49, C, s); },
                       kFamPool, 0.0, (double)N * C * (49 * 4 + 4)));
    }
  }

  // ------------------------------------------------------------------------------------------------ backward
  bwd_.clear();
  chunks_.clear();
  // Cross-stream schedule.  Every wgrad goes to the side stream and is released by the NEXT BatchNorm-backward reduce
  // pass of the main chain:   dgrad(c) | bn_bwd_reduce(below) | wgrad(c) on the side stream || bn_bwd_apply(below) | ...
  // A wgrad CTA takes a whole SM's shared memory, so it can only share the SM with kernels that use none: the apply
  // pass (HBM bound) co-runs with it, the reduce pass (16-24 KB of shared memory) and the dgrad cannot.  The dy buffer
  // a wgrad reads rotates over three buffers; before the main stream overwrites one, it waits for its last reader.
  int n_events = 0;
  int cur_block = -1;                          // residual block the ops being pushed belong to (debug_run_block_backward)
  std::vector<Op> held_wgrads;                 // created, not yet placed: released by the next reduce pass
  std::vector<int> deferred;                   // waits to attach to the next main-stream op
  std::map<const void*, int> pending_reader;   // dy buffer -> event of the side-stream wgrad that reads it
  auto attach_deferred = [&](size_t first_new_op) {
    if (deferred.empty() || first_new_op >= bwd_.size()) return;
    for (int ev : deferred) bwd_[first_new_op].wait.push_back(ev);
    deferred.clear();
  };
  auto guard_write = [&](const void* buf) {  // the next main op may overwrite `buf`
    auto it = pending_reader.find(buf);
    if (it != pending_reader.end()) {
      deferred.push_back(it->second);
      pending_reader.erase(it);
    }
  };
  auto release_wgrads = [&]() {  // place the held filter gradients behind the main-stream op pushed last
    if (held_wgrads.empty()) return;
    if (bwd_.back().record < 0) bwd_.back().record = n_events++;
    const int after_main = bwd_.back().record;
    for (Op& w : held_wgrads) {
      w.wait.push_back(after_main);
      bwd_.push_back(w);
    }
    held_wgrads.clear();
  };
  // Traversal directions (L2 reuse between consecutive passes over a tensor): `cur` is the direction (0 first-to-last,
  // 1 last-to-first) in which the gradient the next BatchNorm backward consumes was written.  The reduce pass walks it
  // the opposite way, the apply pass opposite to the reduce pass (it re-reads the same two tensors), and the data
  // gradient that consumes the apply pass' output opposite to that — so the direction flips once per layer.
  int cur = 0;
  auto push_bn_bwd = [&](const Conv& c, const bf16* dA, const uint8_t* mask, bf16* dy, bf16* dz_out,
                         const Conv* second, bf16* dy2) {
    BnBwdArgs a;
    const bool big = (double)N * c.P * c.Q * c.Cout * 2 >= l2_order_min_bytes_;
    if (l2_order_ && big) {
      a.rev_reduce = !cur;
      a.rev_apply = cur;  // writes dy in direction `cur`
    } else {
      cur = 0;  // both passes first-to-last: dy is written first-to-last
    }
    a.dA = dA;
    a.mask = mask;
    a.y = c.y;
    a.M = N * c.P * c.Q;
    a.C = c.Cout;
    a.mean = saved + c.save_off;
    a.rstd = saved + c.save_off + c.Cout;
    a.gamma = P + c.gamma_off;
    a.sums = zero + c.zero_off + 16 * c.Cout;
    a.sums_raw = c.fin ? 0 : 1;  // 0: fp32 totals (sums[2C], sums2[C]) written by the reduce pass's last block per slice
    a.dy = dy;
    a.dz_out = dz_out;
    a.dgamma = G + c.gamma_off;
    a.dbeta = G + c.beta_off;
    a.det.scratch = reinterpret_cast<float*>(ws_ + off_det_bn_);
    a.det.tickets = reinterpret_cast<int*>(ws_ + off_det_bn_ + kDetScratchFloats * 4);
    if (second) {
      const Conv& d = *second;
      a.y2 = d.y;
      a.mean2 = saved + d.save_off;
      a.rstd2 = saved + d.save_off + d.Cout;
      a.gamma2 = P + d.gamma_off;
      a.sums2 = zero + c.zero_off + 32 * c.Cout;
      a.dy2 = dy2;
      a.dgamma2 = G + d.gamma_off;
      a.dbeta2 = G + d.beta_off;
    }
    const double mc = (double)a.M * a.C * 2;
    const double rd = 2 + (mask ? 1.0 / 16 : 0.0) + (second ? 1 : 0);
    guard_write(dy);
    if (second && dy2) guard_write(dy2);  // identity blocks pass the buffer but do not write it
    const size_t first = bwd_.size();
    // small layers (gradient + raw output resident in L2 between the passes): one launch with a grid barrier
    if (fuse_bn_bwd_ && bn_bwd_can_fuse(a) && 2.0 * mc <= kFuseBnBwdBytes) {
      int* counter = reinterpret_cast<int*>(zero + c.zero_off + 40 * (size_t)c.Cout);
      int* wd_flag = device_error_flag();
      bwd_.push_back(Op([a, counter, wd_flag](cudaStream_t s) { return launch_bn_bwd_fused(a, counter, wd_flag, s); },
                        kFamNorm, 0.0,
                        mc * (rd + 1)));  // algorithmic bytes: the second pass re-reads from L2
      bwd_.back().label = "bn_bwd_fused " + c.bn;
      bwd_.back().block = cur_block;
      attach_deferred(first);
      release_wgrads();
      return;
    }
    bwd_.push_back(Op([a](cudaStream_t s) { return launch_bn_bwd_reduce(a, s); }, kFamNorm, 0.0, mc * rd));
    bwd_.back().label = "bn_bwd_reduce " + c.bn;
    bwd_.back().block = cur_block;
    attach_deferred(first);
    release_wgrads();
    bwd_.push_back(Op([a](cudaStream_t s) { return launch_bn_bwd_apply(a, s); }, kFamNorm, 0.0,
                      mc * (rd + 1 + (dz_out ? 1 : 0) + (second ? 1 : 0))));
    bwd_.back().label = "bn_bwd_apply " + c.bn;
    bwd_.back().block = cur_block;
  };
  auto push_wgrad = [&](const Conv& c, const bf16* dy) {
    WgradDesc d;
    d.dy = dy;
    d.x = c.x;
    d.N = N;
    d.H = c.H;
    d.W = c.W;
    d.C = c.Cin;
    fill_fwd_geometry(&d, c.R, c.R, c.stride, c.pad);
    d.Cout = c.Cout;
    d.dw = G + c.w_off;
    d.scratch = reinterpret_cast<float*>(ws_ + off_wgrad_scratch_[0]);  // side stream (all of them when single-stream)
    WgradPlan plan;
    std::string e2 = plan_wgrad(d, &plan);
    if (!e2.empty()) {
      err = e2;
      return;
    }
    const double m = (double)N * c.P * c.Q;
    const double kdim = (double)c.R * c.R * c.Cin;
    Op w([plan](cudaStream_t s) { return run_wgrad(plan, s); }, kFamWgrad, 2.0 * m * c.Cout * kdim,
         2.0 * (m * c.Cout + (double)N * c.H * c.W * c.Cin) + 4.0 * c.Cout * kdim);
    w.label = "wgrad " + c.name;
    w.block = cur_block;
    w.side = true;
    w.record = n_events++;
    pending_reader[dy] = w.record;
    held_wgrads.push_back(w);
  };
  auto push_dgrad = [&](const Conv& c, const bf16* dy, bf16* dx, int accumulate) {
    // reads dy (written in direction `cur` by the apply pass) the other way round and leaves dx in that direction
    const bool big_dy = (double)N * c.P * c.Q * c.Cout * 2 >= l2_order_min_bytes_;
    const int dgrad_rev = (l2_order_ && big_dy) ? !cur : 0;
    cur = dgrad_rev;
    std::vector<DgradClass> cls = dgrad_classes(c.H, c.W, c.R, c.R, c.stride, c.pad);
    size_t off = 0;
    for (const DgradClass& k : cls) {
      if (k.ntaps == 0) {
        if (!accumulate) err = "dgrad with an empty parity class needs accumulate mode (" + c.name + ")";
        continue;
      }
      GatherConv gc;
      gc.src = dy;
      gc.N = N;
      gc.H = c.P;
      gc.W = c.Q;
      gc.C = c.Cout;
      gc.P = k.Pc;
      gc.Q = k.Qc;
      gc.stride = 1;
      gc.base_h = k.base_h;
      gc.base_w = k.base_w;
      gc.ntaps = k.ntaps;
      for (int t = 0; t < k.ntaps; ++t) {
        gc.tap_h[t] = k.tap_h[t];
        gc.tap_w[t] = k.tap_w[t];
      }
      gc.wpk = wd + c.wd_off + off;
      gc.Cout = c.Cin;
      gc.out = dx;
      gc.ldo = c.Cin;
      if (c.stride > 1) {
        gc.out_mode = 1;
        gc.oH = c.H;
        gc.oW = c.W;
        gc.o_stride = c.stride;
        gc.o_h0 = k.ph;
        gc.o_w0 = k.pw;
      }
      gc.accumulate = accumulate;
      gc.rev_m = dgrad_rev;
      push_conv(bwd_, gc, "dgrad " + c.name + (accumulate ? " (+=)" : "") + (c.stride > 1 ? " class " + std::to_string(k.ph) + std::to_string(k.pw) : ""));
      if (err.empty()) bwd_.back().block = cur_block;
      off += (size_t)c.Cin * k.ntaps * c.Cout;
    }
  };

  // R3M_TEST_MUTATION=drop_skip_add deliberately breaks the schedule (the skip path's gradient is overwritten): the
  // parity tests assert that they catch it (tests/test_block_backward_gpu.py::test_mutation_is_caught)
  const char* mut_env = std::getenv("R3M_TEST_MUTATION");
  const bool mutate_drop_skip_add = mut_env && std::string(mut_env) == "drop_skip_add";
  bf16* d_out = g[0];
  bf16* d_in = g[1];
  bf16 *s2 = g[3], *s3 = g[4];
  bf16* dy_ring[3] = {g[2], g[5], g[6]};
  int dy_slot = 0;
  auto next_dy = [&]() {
    dy_slot = (dy_slot + 1) % 3;
    return dy_ring[dy_slot];
  };
  {
    const float* dE = reinterpret_cast<const float*>(ws_ + off_dE_);
    bf16* dst = d_out;
    const int C = D_;
    bwd_.push_back(Op([dE, dst, N, C](cudaStream_t s) { return launch_avgpool_bwd(dE, dst, N, 49, C, s); }, kFamPool, 0.0,
                      (double)N * C * (4 + 49 * 2)));
  }
  for (int bi = (int)blocks_.size() - 1; bi >= 0; --bi) {
    Block* blk = blocks_[bi];
    cur_block = bi;
    blk->d_out = d_out;
    blk->d_in = d_in;
    const int n = (int)blk->main.size();
    Conv& last = *convs_[blk->main[n - 1]];
    const bool has_ds = blk->ds >= 0;
    // last BN of the residual branch: masked by the block output's ReLU.  Identity blocks: the masked gradient is
    // also the skip path's share and seeds d_in.  Downsample blocks: the same masked gradient drives the downsample
    // BatchNorm's backward in the same two passes (dy_ds -> s3).
    bf16* dy = next_dy();
    push_bn_bwd(last, d_out, last.mask, dy, has_ds ? nullptr : d_in, has_ds ? convs_[blk->ds] : nullptr, s3);
    for (int i = n - 1; i >= 0; --i) {
      Conv& c = *convs_[blk->main[i]];
      // data gradient first (main stream), then the filter gradient of the same layer on the side stream
      if (i > 0) {
        Conv& prev = *convs_[blk->main[i - 1]];
        push_dgrad(c, dy, s2, 0);
        push_wgrad(c, dy);
        bf16* dy_prev = next_dy();
        push_bn_bwd(prev, s2, prev.mask, dy_prev, nullptr, nullptr, nullptr);
        dy = dy_prev;
      } else {
        // identity blocks: the masked gradient of the block output (written to d_in by the BatchNorm backward above)
        // is the skip path's share, the first conv's data gradient is added on top
        push_dgrad(c, dy, d_in, (has_ds || mutate_drop_skip_add) ? 0 : 1);
        push_wgrad(c, dy);
      }
    }
    if (has_ds) {
      Conv& d = *convs_[blk->ds];
      push_dgrad(d, s3, d_in, 1);
      push_wgrad(d, s3);
    }
    std::swap(d_out, d_in);
    if (!err.empty()) return err;
    // Gradient chunks for the overlapped all-reduce: when the FIRST block of layer 2 / 3 / 4 has been scheduled, every
    // gradient from that layer's first parameter to the end of the flat buffer is final once (a) the main stream has
    // passed this point (BatchNorm gradients) and (b) the side stream has finished this block's filter gradients.
    if (bi > 0 && blk->ds >= 0) {  // the first blocks of layers 2-4 are the ones with a downsample branch (bi > 0)
      GradChunk gc;
      gc.begin = convs_[blk->main[0]]->w_off;
      if (bwd_.back().record < 0) bwd_.back().record = n_events++;
      gc.main_event = bwd_.back().record;
      gc.side_event = held_wgrads.empty() ? -1 : held_wgrads.back().record;
      chunks_.push_back(gc);
    }
  }
  {
    // stem: maxpool backward + ReLU mask + BN backward fused in two passes -> filter gradient (no data gradient)
    release_wgrads();  // layer1.0's filter gradients: behind its last dgrad
    cur_block = -1;
    Conv& st = *convs_[0];
    StemBwdArgs sb;
    sb.dA = d_out;
    sb.argmax = argmax;
    sb.ymax = ymax;
    sb.y = st.y;
    sb.N = N;
    sb.mean = saved + st.save_off;
    sb.rstd = saved + st.save_off + 64;
    sb.gamma = P + st.gamma_off;
    sb.sums = zero + st.zero_off + 16 * 64;
    sb.sums_raw = 1;
    sb.dy = s2;
    sb.dgamma = G + st.gamma_off;
    sb.dbeta = G + st.beta_off;
    sb.det.scratch = reinterpret_cast<float*>(ws_ + off_det_bn_);
    sb.det.tickets = reinterpret_cast<int*>(ws_ + off_det_bn_ + kDetScratchFloats * 4);
    // algorithmic bytes: reduce pass over the pooled elements (dA, ymax, codes); apply pass y + pooled gradient + codes
    // read, dy written
    bwd_.push_back(Op([sb](cudaStream_t s) { return launch_stem_bwd(sb, s); }, kFamNorm, 0.0,
                      (double)N * 64 * (112.0 * 112 * 4 + 56.0 * 56 * 8)));
    bwd_.back().label = "stem_bwd (maxpool + relu + bn1 backward, 2 launches)";
    bwd_.back().nlaunch = 2;
    WgradDesc d;
    d.dy = s2;
    d.x = xs;
    d.N = N;
    d.H = 112;
    d.W = 112;
    d.C = 64;
    d.P = 112;
    d.Q = 112;
    d.stride = 1;
    d.base_h = -2;
    d.base_w = 0;
    d.ntaps = 4;
    for (int t = 0; t < 4; ++t) {
      d.tap_h[t] = t;
      d.tap_w[t] = 0;
    }
    d.Cout = 64;
    d.dw = stem_dwp;
    d.scratch = reinterpret_cast<float*>(ws_ + off_wgrad_scratch_[1]);  // main stream: may overlap the side stream's
    WgradPlan plan;
    err = plan_wgrad(d, &plan);
    if (!err.empty()) return err;
    bwd_.push_back(Op([plan](cudaStream_t s) { return run_wgrad(plan, s); }, kFamWgrad,
                      2.0 * N * 112.0 * 112 * 64 * 147, 2.0 * N * 112.0 * 112 * 128));
    float* dst = G + st.w_off;
    bwd_.push_back(Op([stem_dwp, dst](cudaStream_t s) { return launch_stem_unpack_grad(stem_dwp, dst, s); }, kFamOptim));
    // join: the step's last backward op waits for every filter gradient still running on the side stream
    for (auto& kv : pending_reader) bwd_.back().wait.push_back(kv.second);
    pending_reader.clear();
    GradChunk tail;  // the stem and layer 1: final when the whole backward pass is
    tail.begin = 0;
    bwd_.back().record = n_events++;
    tail.main_event = bwd_.back().record;
    tail.side_event = -1;
    chunks_.push_back(tail);
  }

  {
    std::vector<char> recorded(n_events, 0);
    for (const Op& op : bwd_) {
      for (int ev : op.wait)
        if (!recorded[ev]) return "backward schedule: op '" + op.label + "' waits on an event recorded later";
      if (op.record >= 0) recorded[op.record] = 1;
    }
  }
  while ((int)evs_.size() < n_events) {
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return "cudaEventCreate failed";
    evs_.push_back(ev);
  }
  if (!side_ && cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking) != cudaSuccess) return "side stream creation failed";
  for (auto& kv : graphs_)  // a re-plan (re-bind) invalidates every captured pointer
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
  ext_evs_.resize(evs_.size(), nullptr);
  for (const GradChunk& c : chunks_)
    for (int id : {c.main_event, c.side_event})
      if (id >= 0 && !ext_evs_[id] && cudaEventCreateWithFlags(&ext_evs_[id], cudaEventDisableTiming) != cudaSuccess)
        return "cudaEventCreate failed";
  {
    const char* env = std::getenv("R3M_STEP_GRAPH");
    use_graph_ = !(env && env[0] == '0');
  }
  {
    const char* env = std::getenv("R3M_WGRAD_STREAM");
    use_side_ = !(env && env[0] == '0');
  }

  // ------------------------------------------------------------------------------------------------ re-packs
  repack_.clear();
  std::vector<PackDgradEntry> packs;
  int pack_blocks = 0;
  for (Conv* cp : convs_) {
    const Conv& c = *cp;
    if (c.stem) {
      const float* w = P + c.w_off;
      repack_.push_back(Op([w, stem_wp](cudaStream_t s) { return launch_stem_pack(w, stem_wp, s); }, kFamOptim));
      continue;
    }
    std::vector<DgradClass> cls = dgrad_classes(c.H, c.W, c.R, c.R, c.stride, c.pad);
    size_t off = 0;
    for (const DgradClass& k : cls) {
      if (k.ntaps == 0) continue;
      PackDgradEntry pe;
      pe.w = P + c.w_off;
      pe.out = wd + c.wd_off + off;
      pe.Cout = c.Cout;
      pe.T = c.R * c.R;
      pe.Cin = c.Cin;
      pe.nt = k.ntaps;
      for (int t = 0; t < 16; ++t) pe.taps[t] = t < k.ntaps ? k.src_r[t] * c.R + k.src_s[t] : 0;
      pe.block_begin = pack_blocks;
      pack_blocks += ((c.Cin + 31) / 32) * ((c.Cout + 31) / 32) * k.ntaps;
      packs.push_back(pe);
      off += (size_t)c.Cin * k.ntaps * c.Cout;
    }
  }
  if (!packs.empty()) {
    if (packs.size() > kMaxPackEntries) return "too many dgrad re-pack entries";
    PackDgradEntry* table_dev = reinterpret_cast<PackDgradEntry*>(ws_ + off_pack_);
    cudaError_t ce = cudaMemcpy(table_dev, packs.data(), packs.size() * sizeof(PackDgradEntry), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) return std::string("pack table upload: ") + cudaGetErrorString(ce);
    const int entries = (int)packs.size(), blocks = pack_blocks;
    repack_.push_back(Op([table_dev, entries, blocks](cudaStream_t s) {
      return launch_pack_dgrad_multi(table_dev, entries, blocks, s);
    }, kFamOptim, 0.0, 6.0 * (double)nwd_));
    repack_.back().label = "pack_dgrad (all filters)";
  }
  return err;
}

std::string Engine::run(const std::vector<Op>& ops, cudaStream_t stream) {
  // profiling (an event pair around every launch) keeps everything on one stream so that per-family times add up
  const bool two_streams = use_side_ && !profiling_ && side_ != nullptr;
  for (const Op& op : ops) {
    cudaStream_t s = (two_streams && op.side) ? side_ : stream;
    if (two_streams)
      for (int ev : op.wait) {
        cudaError_t we = cudaStreamWaitEvent(s, evs_[ev], 0);
        if (we != cudaSuccess) return std::string("stream wait failed: ") + cudaGetErrorString(we);
      }
    // a kernel node with several incoming edges (a cross-stream join) cannot carry a programmatic dependency
    static const bool join_pdl = std::getenv("R3M_GRAPH_JOIN_PDL") != nullptr;
    const bool suppress = capturing_ && two_streams && !op.wait.empty() && !join_pdl;
    if (suppress) ++g_pdl_suppress;
    cudaError_t e = launch(op, s);
    if (suppress) --g_pdl_suppress;
    if (e != cudaSuccess) return std::string("kernel launch failed: ") + cudaGetErrorString(e);
    if (op.record >= 0) {  // also in single-stream runs: the gradient-chunk markers are read by wait_grad_chunk
      cudaError_t re = cudaEventRecord(evs_[op.record], s);
      if (re == cudaSuccess && capturing_ && ext_evs_[op.record] != nullptr)
        re = cudaEventRecordWithFlags(ext_evs_[op.record], s, cudaEventRecordExternal);
      if (re != cudaSuccess) return std::string("event record failed: ") + cudaGetErrorString(re);
    }
  }
  return std::string();
}

std::string Engine::run_cached(const std::vector<uint64_t>& key, cudaStream_t stream,
                               const std::function<std::string(cudaStream_t)>& body, bool* graphed) {
  *graphed = false;
  constexpr size_t kMaxGraphs = 8;  // distinct input-pointer sets worth keeping (double-buffered feeders need two)
  if (!use_graph_ || profiling_ || capturing_) return body(stream);
  auto it = graphs_.find(key);
  if (it == graphs_.end()) {
    if (graphs_.size() >= kMaxGraphs) {
      // evict the least recently used entry (input buffers that went away); a key is only captured on its second
      // sighting, so a stream of never-repeating pointers costs map entries, not captures
      auto victim = graphs_.begin();
      for (auto jt = graphs_.begin(); jt != graphs_.end(); ++jt)
        if (jt->second.last_use < victim->second.last_use) victim = jt;
      if (victim->second.exec) cudaGraphExecDestroy(victim->second.exec);
      graphs_.erase(victim);
    }
    it = graphs_.emplace(key, GraphEntry()).first;
  }
  GraphEntry& ge = it->second;
  ge.last_use = ++graph_clock_;
  if (ge.exec == nullptr && !ge.failed && ge.seen >= 1) {
    // captured on a private stream (the caller's may be the legacy default stream, which cannot be captured)
    if (!cap_ && cudaStreamCreateWithFlags(&cap_, cudaStreamNonBlocking) != cudaSuccess) cap_ = nullptr;
    cudaGraph_t graph = nullptr;
    const int before = launches_;
    if (cap_ && cudaStreamBeginCapture(cap_, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
      capturing_ = true;
      const std::string cerr = body(cap_);
      capturing_ = false;
      const cudaError_t ee = cudaStreamEndCapture(cap_, &graph);
      if (cerr.empty() && ee == cudaSuccess && graph != nullptr &&
          cudaGraphInstantiate(&ge.exec, graph, 0) == cudaSuccess) {
        ge.launches = launches_ - before;
      } else {
        ge.exec = nullptr;
        ge.failed = true;
        if (std::getenv("R3M_GRAPH_DEBUG"))
          fprintf(stderr, "r3m_b200: step-graph capture failed (kind %llu): body '%s', end-capture %s, last error %s\n",
                  (unsigned long long)key[0], cerr.c_str(), cudaGetErrorString(ee), cudaGetErrorString(cudaPeekAtLastError()));
      }
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
    } else {
      ge.failed = true;
      (void)cudaGetLastError();
    }
    launches_ = before;
  }
  ++ge.seen;
  if (ge.exec == nullptr) return body(stream);
  const cudaError_t e = cudaGraphLaunch(ge.exec, stream);
  if (e != cudaSuccess) return std::string("graph launch: ") + cudaGetErrorString(e);
  launches_ += ge.launches;
  ++graph_replays_;
  *graphed = true;
  return std::string();
}

cudaError_t Engine::launch(const Op& op, cudaStream_t stream) {
  launches_ += op.nlaunch;
  if (!profiling_) return op.fn(stream);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, stream);
  cudaError_t e = op.fn(stream);
  cudaEventRecord(b, stream);
  prof_events_.push_back(a);
  prof_events_.push_back(b);
  prof_ops_family_.push_back(op.family);
  prof_flops_.push_back(op.flops);
  prof_bytes_.push_back(op.bytes);
  prof_labels_run_.push_back(op.label);
  return e;
}

std::string Engine::profile_update(const void* obs, const int* perms, const float* lang_emb, const float* lang_mask,
                                   const Hyper& h, float lr, int step, double* out, cudaStream_t stream) {
  profiling_ = true;
  prof_events_.clear();
  prof_ops_family_.clear();
  prof_flops_.clear();
  prof_bytes_.clear();
  prof_labels_run_.clear();
  std::string err = update_grads(obs, perms, lang_emb, lang_mask, h, 0, stream);
  if (err.empty()) err = adam_step(lr, 1.0f, step, stream);
  profiling_ = false;
  cudaError_t e = cudaStreamSynchronize(stream);
  if (err.empty() && e != cudaSuccess) err = std::string("profile sync: ") + cudaGetErrorString(e);
  for (int i = 0; i < kNumFamilies * 4; ++i) out[i] = 0.0;
  prof_last_.clear();
  prof_labels_ = prof_labels_run_;
  for (size_t i = 0; i < prof_ops_family_.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, prof_events_[2 * i], prof_events_[2 * i + 1]);
    const int f = prof_ops_family_[i];
    prof_last_.push_back(f);
    prof_last_.push_back(ms);
    prof_last_.push_back(prof_flops_[i]);
    prof_last_.push_back(prof_bytes_[i]);
    out[f * 4 + 0] += ms;
    out[f * 4 + 1] += prof_flops_[i];
    out[f * 4 + 2] += prof_bytes_[i];
    out[f * 4 + 3] += 1.0;
  }
  for (cudaEvent_t ev : prof_events_) cudaEventDestroy(ev);
  prof_events_.clear();
  return err;
}

std::string Engine::sync_weights(cudaStream_t stream) {
  if (!bound_) return "engine has no workspace bound";
  launches_ = 0;
  const float* P = reinterpret_cast<const float*>(pws_ + off_P_);
  void* Pb = pws_ + off_Pb_;
  const size_t n = nparams_;
  cudaError_t e = launch(Op([=](cudaStream_t s) { return launch_cast_bf16(P, Pb, n, s); }, kFamOptim, 0.0, (double)n * 6),
                         stream);
  if (e != cudaSuccess) return std::string("cast: ") + cudaGetErrorString(e);
  return run(repack_, stream);
}

std::string Engine::set_obs_format(int format) {
  if (format != kObsF32NCHW && format != kObsU8NCHW && format != kObsU8NHWC) return "unknown observation format";
  if (format != obs_format_ && eval_graph_) {  // the captured eval graph holds the preprocess kernel of the old format
    cudaGraphExecDestroy(eval_graph_);
    eval_graph_ = nullptr;
    eval_calls_ = 0;
  }
  obs_format_ = format;
  return std::string();
}

std::string Engine::forward(const void* obs, int train, float* out, cudaStream_t stream) {
  if (!bound_) return "engine has no workspace bound";
  launches_ = 0;
  cudaError_t e;
  std::string err;
  // an eval forward overwrites the activations and the saved-statistics slots (BN fold coefficients) that a pending
  // update_grads(obs = NULL) / backward() would read
  if (!train) fwd_train_valid_ = false;
  if (!train && precision_ == 1) {
    // parity tier: fp32 storage, kind::tf32 tensor cores, BatchNorm folded into the conv epilogues
    float* xs32 = reinterpret_cast<float*>(ws_ + off_xs_ + t32_xs_);
    const int N = N_, fmt = obs_format_;
    e = launch(Op([obs, xs32, N, fmt](cudaStream_t s) { return launch_preprocess_stem_f32(obs, fmt, xs32, N, s); },
                  kFamNorm),
               stream);
    if (e != cudaSuccess) return std::string("preprocess (tf32): ") + cudaGetErrorString(e);
    err = run(fwd_eval_tf32_, stream);
    if (!err.empty()) return err;
    if (out) {
      e = cudaMemcpyAsync(out, ws_ + off_xs_ + t32_E_, (size_t)N_ * D_ * 4, cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return std::string("copy out: ") + cudaGetErrorString(e);
      const int* flag = device_error_flag();
      const size_t n = (size_t)N_ * D_;
      e = launch(Op([flag, out, n](cudaStream_t s) { return launch_poison_on_flag(flag, out, n, s); }, kFamLoss), stream);
      if (e != cudaSuccess) return std::string("poison_on_flag: ") + cudaGetErrorString(e);
    }
    return std::string();
  }
  if (!train && N_ <= kGraphMaxFrames && !profiling_) {
    // Launch-latency-bound regime (load_r3m users, r3m/example.py: batch 1-4): ~25 kernels of a few microseconds.
    // The frames are copied to a fixed staging buffer and the whole eval forward is replayed as one CUDA graph
    // (captured on the second call, once every kernel's attributes have been configured by a plain first call).
    void* stage = ws_ + off_obs_stage_;
    const int fmt = obs_format_;
    e = cudaMemcpyAsync(stage, obs, (size_t)N_ * 3 * 224 * 224 * (fmt == kObsF32NCHW ? 4 : 1), cudaMemcpyDeviceToDevice,
                        stream);
    if (e != cudaSuccess) return std::string("stage frames: ") + cudaGetErrorString(e);
    ++eval_calls_;
    if (eval_graph_ == nullptr && eval_calls_ >= 2) {
      // captured on a private stream (the caller's may be the legacy default stream, which cannot be captured)
      cudaGraph_t graph = nullptr;
      cudaStream_t cap = nullptr;
      if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) == cudaSuccess &&
          cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
        cudaError_t ce = launch_preprocess_stem(stage, fmt, ws_ + off_xs_, N_, cap);
        std::string cerr = run(fwd_eval_, cap);
        cudaError_t ee = cudaStreamEndCapture(cap, &graph);
        if (ce == cudaSuccess && cerr.empty() && ee == cudaSuccess && graph != nullptr) {
          if (cudaGraphInstantiate(&eval_graph_, graph, 0) != cudaSuccess) eval_graph_ = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
      }
      if (cap) cudaStreamDestroy(cap);
      launches_ = 0;
    }
    if (eval_graph_ != nullptr) {
      e = cudaGraphLaunch(eval_graph_, stream);
      if (e != cudaSuccess) return std::string("graph launch: ") + cudaGetErrorString(e);
      launches_ = (int)fwd_eval_.size() + 1;
    } else {
      e = launch_preprocess_stem(stage, fmt, ws_ + off_xs_, N_, stream);
      if (e != cudaSuccess) return std::string("preprocess: ") + cudaGetErrorString(e);
      ++launches_;
      err = run(fwd_eval_, stream);
      if (!err.empty()) return err;
    }
  } else {
    auto body = [this, obs, train](cudaStream_t st) -> std::string {
      if (train) {
        cudaError_t me = cudaMemsetAsync(ws_ + off_zero_, 0, zero_bytes_, st);
        if (me != cudaSuccess) return std::string("memset: ") + cudaGetErrorString(me);
      }
      void* xs = ws_ + off_xs_;
      const int N = N_, fmt = obs_format_;
      cudaError_t pe =
          launch(Op([obs, xs, N, fmt](cudaStream_t s) { return launch_preprocess_stem(obs, fmt, xs, N, s); }, kFamNorm,
                    0.0, (double)N * (3.0 * 224 * 224 * (fmt == kObsF32NCHW ? 4 : 1) + 112.0 * 112 * 64 * 2)),
                 st);
      if (pe != cudaSuccess) return std::string("preprocess: ") + cudaGetErrorString(pe);
      return run(train ? fwd_train_ : fwd_eval_, st);
    };
    bool graphed = false;
    if (train)
      err = run_cached({1u, (uint64_t)reinterpret_cast<uintptr_t>(obs), (uint64_t)obs_format_}, stream, body, &graphed);
    else
      err = body(stream);
    if (!err.empty()) return err;
    fwd_train_valid_ = train != 0;
    fwd_launches_ = launches_;
  }
  if (out) {
    e = cudaMemcpyAsync(out, ws_ + off_E_, (size_t)N_ * D_ * 4, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return std::string("copy out: ") + cudaGetErrorString(e);
    const int* flag = device_error_flag();
    const size_t n = (size_t)N_ * D_;
    e = launch(Op([flag, out, n](cudaStream_t s) { return launch_poison_on_flag(flag, out, n, s); }, kFamLoss), stream);
    if (e != cudaSuccess) return std::string("poison_on_flag: ") + cudaGetErrorString(e);
  }
  return std::string();
}

std::string Engine::update_grads(const void* obs, const int* perms, const float* lang_emb, const float* lang_mask,
                                 const Hyper& h, int eval, cudaStream_t stream) {
  if (!bound_) return "engine has no workspace bound";
  if (B_ == 0) return "update needs frames == 5 * clips (r3m/trainer.py:39-40)";
  if (h.langweight > 0.f && !lang_) return "engine was created without the language head";
  if (h.langweight > 0.f && (!lang_emb || !lang_mask)) return "language head needs lang_emb and lang_mask";
  launches_ = 0;
  cudaError_t e;
  std::string err;
  if (obs != nullptr) {
    e = cudaMemsetAsync(ws_ + off_zero_, 0, zero_bytes_, stream);
    if (e != cudaSuccess) return std::string("memset: ") + cudaGetErrorString(e);
    {
      void* xs = ws_ + off_xs_;
      const int N = N_, fmt = obs_format_;
      e = launch(Op([obs, xs, N, fmt](cudaStream_t s) { return launch_preprocess_stem(obs, fmt, xs, N, s); }, kFamNorm,
                    0.0, (double)N * (3.0 * 224 * 224 * (fmt == kObsF32NCHW ? 4 : 1) + 112.0 * 112 * 64 * 2)),
                 stream);
    }
    if (e != cudaSuccess) return std::string("preprocess: ") + cudaGetErrorString(e);
    err = run(eval ? fwd_eval_ : fwd_train_, stream);
    if (!err.empty()) return err;
  } else {
    // obs == NULL: the caller has already enqueued forward(obs, train) on this stream (the activations, batch
    // statistics and embeddings of that call are in place), so that its host-side preparation of perms / lang inputs
    // overlaps the forward pass instead of delaying the step's first kernel
    if (eval || !fwd_train_valid_) return "update_grads(obs = NULL) needs a preceding train-mode forward() on this engine";
    launches_ = fwd_launches_;
  }
  fwd_train_valid_ = false;
  // everything behind the forward pass: gradient clear, loss heads, language head, backward, watchdog flag
  auto tail = [this, perms, lang_emb, lang_mask, h, eval](cudaStream_t st) -> std::string {
    cudaError_t e;
    if (!eval) {
      e = cudaMemsetAsync(pws_ + off_G_, 0, nparams_ * 4, st);
      if (e != cudaSuccess) return std::string("memset: ") + cudaGetErrorString(e);
    }
    const float* E = reinterpret_cast<const float*>(ws_ + off_E_);
    float* dE = eval ? nullptr : reinterpret_cast<float*>(ws_ + off_dE_);
    float* metrics = reinterpret_cast<float*>(ws_ + off_metrics_);
    {
      const int N = N_, D = D_, B = B_;
      const float l2w = h.l2weight, l1w = h.l1weight, tcnw = h.tcnweight;
      float* lp_scratch = reinterpret_cast<float*>(ws_ + off_det_loss_);
      float* tcn_scratch = lp_scratch + (size_t)N * 4;
      Op lp([=](cudaStream_t s) { return launch_loss_lp(E, dE, N, D, l2w, l1w, metrics, s, lp_scratch); }, kFamLoss, 0.0,
            (double)N * D * 8);
      lp.nlaunch = 2;
      e = launch(lp, st);
      if (e != cudaSuccess) return std::string("loss_lp: ") + cudaGetErrorString(e);
      if (tcnw > 0.f) {
        const int l2dist = l2dist_ ? 1 : 0;
        Op tcn([=](cudaStream_t s) { return launch_loss_tcn(E, dE, perms, B, D, tcnw, l2dist, metrics, s, tcn_scratch); },
               kFamLoss, 0.0, (double)B * 18 * D * 4 * 2);
        tcn.nlaunch = dE ? 3 : 2;
        e = launch(tcn, st);
        if (e != cudaSuccess) return std::string("loss_tcn: ") + cudaGetErrorString(e);
      }
    }
    if (h.langweight > 0.f) {
      float* P = reinterpret_cast<float*>(pws_ + off_P_);
      float* G = reinterpret_cast<float*>(pws_ + off_G_);
      LangParams lp;
      for (int l = 0; l < 5; ++l) {
        lp.w[l] = P + lang_w_off_[l];
        lp.b[l] = P + lang_b_off_[l];
        lp.dw[l] = G + lang_w_off_[l];
        lp.db[l] = G + lang_b_off_[l];
      }
      LangWorkspace lw;
      lang_carve_workspace(reinterpret_cast<float*>(ws_ + off_lang_ws_), lang_dims_, &lw);
      const LangDims ld = lang_dims_;
      const LangTc* tc = &lang_tc_;
      const float langw = h.langweight;
      int n_lang = 0;
      int* n_ptr = &n_lang;
      // the head is ~25 launches; it is charged to the "lang" family as one profiled unit
      // algorithmic MACs of the factorised head: layer 1 once per distinct row (B e0 rows, 5B e_t rows, B sentences),
      // layers 2-4 over the 15B evaluations
      const double rows = ld.rows(), H = ld.H, Bc = ld.B, Dd = ld.D, Ll = ld.L;
      const double l1_mac = (6.0 * Bc * Dd + Bc * Ll) * H;
      const double fwd_mac = l1_mac + rows * (3 * H * H + H);
      const double bwd_mac = 2.0 * rows * 3 * H * H + (l1_mac + 6.0 * Bc * Dd * H);
      e = launch(Op([=](cudaStream_t s) {
                   return lang_head_run(ld, lp, lw, E, dE, perms, lang_emb, lang_mask, langw, metrics, n_ptr, s, tc);
                 },
                 kFamLang, 2.0 * (fwd_mac + (eval ? 0.0 : bwd_mac)), 0.0),
                 st);
      if (e != cudaSuccess) return std::string("lang head: ") + cudaGetErrorString(e);
      launches_ += n_lang - 1;
    }
    if (!eval) {
      const std::string berr = run(bwd_, st);
      if (!berr.empty()) return berr;
    }
    const int* flag = device_error_flag();
    e = launch(Op([flag, metrics](cudaStream_t s) { return launch_publish_flag(flag, metrics, s); }, kFamLoss), st);
    if (e != cudaSuccess) return std::string("publish_flag: ") + cudaGetErrorString(e);
    return std::string();
  };
  bool graphed = false;
  if (!eval) {
    auto bits = [](float v) {
      uint32_t u;
      memcpy(&u, &v, 4);
      return (uint64_t)u;
    };
    err = run_cached({2u, (uint64_t)reinterpret_cast<uintptr_t>(perms), (uint64_t)reinterpret_cast<uintptr_t>(lang_emb),
                      (uint64_t)reinterpret_cast<uintptr_t>(lang_mask), bits(h.l2weight), bits(h.l1weight),
                      bits(h.langweight), bits(h.tcnweight), (uint64_t)(l2dist_ ? 1 : 0), (uint64_t)(use_side_ ? 1 : 0)},
                     stream, tail, &graphed);
    last_bwd_graphed_ = graphed;
  } else {
    err = tail(stream);
  }
  return err;
}

std::string Engine::backward(const float* dE, cudaStream_t stream) {
  if (!bound_) return "engine has no workspace bound";
  if (!fwd_train_valid_) return "backward() needs a preceding train-mode forward() on this engine (its activations are gone)";
  launches_ = 0;
  fwd_train_valid_ = false;
  last_bwd_graphed_ = false;
  cudaError_t e = cudaMemcpyAsync(ws_ + off_dE_, dE, (size_t)N_ * D_ * 4, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return std::string("copy dE: ") + cudaGetErrorString(e);
  std::string err = run(bwd_, stream);
  if (!err.empty()) return err;
  const int* flag = device_error_flag();
  float* metrics = reinterpret_cast<float*>(ws_ + off_metrics_);
  e = launch(Op([flag, metrics](cudaStream_t s) { return launch_publish_flag(flag, metrics, s); }, kFamLoss), stream);
  if (e != cudaSuccess) return std::string("publish_flag: ") + cudaGetErrorString(e);
  return std::string();
}

int Engine::num_grad_chunks() const { return (int)chunks_.size(); }

std::string Engine::grad_chunk(int k, size_t* begin, size_t* end) const {
  if (k < 0 || k >= (int)chunks_.size()) return "gradient chunk index out of range";
  *begin = chunks_[k].begin;
  *end = (k == 0) ? nparams_ : chunks_[k - 1].begin;
  return std::string();
}

std::string Engine::wait_grad_chunk(int k, cudaStream_t stream) {
  if (k < 0 || k >= (int)chunks_.size()) return "gradient chunk index out of range";
  const GradChunk& c = chunks_[k];
  // a replayed graph records the chunk markers through their external twins (see run())
  const std::vector<cudaEvent_t>& evs = last_bwd_graphed_ ? ext_evs_ : evs_;
  cudaError_t e = cudaStreamWaitEvent(stream, evs[c.main_event], 0);
  if (e == cudaSuccess && c.side_event >= 0) e = cudaStreamWaitEvent(stream, evs[c.side_event], 0);
  if (e != cudaSuccess) return std::string("stream wait failed: ") + cudaGetErrorString(e);
  return std::string();
}

int Engine::num_blocks() const { return (int)blocks_.size(); }

std::string Engine::debug_block(int block, int what, void** ptr, size_t* count) const {
  if (!bound_) return "engine has no workspace bound";
  if (block < 0 || block >= (int)blocks_.size()) return "block index out of range";
  const Block& b = *blocks_[block];
  const Conv& last = *convs_[b.main.back()];
  const size_t in_elems = (size_t)N_ * b.Hin * b.Win * b.Cin, out_elems = last.out_elems(N_);
  switch (what) {
    case 0: *ptr = const_cast<bf16*>(b.x_in); *count = in_elems; break;
    case 1: *ptr = b.a_out; *count = out_elems; break;
    case 2: *ptr = b.d_out; *count = out_elems; break;
    case 3: *ptr = b.d_in; *count = in_elems; break;
    default: return "unknown block buffer";
  }
  return std::string();
}

std::string Engine::debug_run_block_backward(int block, cudaStream_t stream) {
  if (!bound_) return "engine has no workspace bound";
  if (block < 0 || block >= (int)blocks_.size()) return "block index out of range";
  launches_ = 0;
  // the BatchNorm-backward sums of the block's layers are accumulated into: clear them (a full step clears the whole
  // zeroed region once, before its forward pass)
  const Block& b = *blocks_[block];
  std::vector<int> cs(b.main);
  if (b.ds >= 0) cs.push_back(b.ds);
  for (int ci : cs) {
    const Conv& c = *convs_[ci];
    cudaError_t e = cudaMemsetAsync(reinterpret_cast<float*>(ws_ + off_zero_) + c.zero_off + 16 * c.Cout, 0,
                                    (24 * (size_t)c.Cout + 8) * 4, stream);
    if (e != cudaSuccess) return std::string("memset: ") + cudaGetErrorString(e);
  }
  const bool saved = profiling_;
  profiling_ = false;
  const bool side = use_side_;
  use_side_ = false;  // one stream: the ops of a block are in dependency order inside bwd_
  std::vector<Op> ops;
  for (const Op& op : bwd_)
    if (op.block == block) ops.push_back(op);
  std::string err = run(ops, stream);
  use_side_ = side;
  profiling_ = saved;
  return err;
}

std::string Engine::adam_step(float lr, float grad_scale, int step, cudaStream_t stream) {
  if (!bound_) return "engine has no workspace bound";
  if (step < 1) return "Adam step count starts at 1";
  launches_ = 0;
  float* P = reinterpret_cast<float*>(pws_ + off_P_);
  const float* G = reinterpret_cast<const float*>(pws_ + off_G_);
  float* M = reinterpret_cast<float*>(pws_ + off_M_);
  float* V = reinterpret_cast<float*>(pws_ + off_V_);
  void* Pb = pws_ + off_Pb_;
  const size_t n = nparams_;
  cudaError_t e = launch(Op([=](cudaStream_t s) {
                              return launch_adam(P, G, M, V, Pb, n, lr, 0.9f, 0.999f, 1e-8f, step, grad_scale, s);
                            },
                            kFamOptim, 0.0, (double)n * (16 + 14)),
                         stream);
  if (e != cudaSuccess) return std::string("adam: ") + cudaGetErrorString(e);
  return run(repack_, stream);
}

}  // namespace r3m
