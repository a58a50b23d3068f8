// extern "C" boundary (include/r3m_b200.h).  Plain pointers and sizes only; no torch types cross this line.
#include "../../include/r3m_b200.h"

#include <cstdio>
#include <string>
#include <vector>

#include "api_util.h"
#include "convops.h"
#include "elementwise.cuh"
#include "engine.h"
#include "loss.cuh"

namespace r3m {
thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return R3M_B200_ERR_CUDA;
}
}  // namespace r3m

using namespace r3m;

extern "C" {

const char* r3m_b200_last_error(void) { return g_last_error.c_str(); }
int r3m_b200_abi_version(void) { return 1; }

int r3m_b200_check_device_flag(void) {
  int* flag = device_error_flag();
  if (!flag) return fail(R3M_B200_ERR_CUDA, "no device error flag (is a CUDA device present?)");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceSynchronize");
  int h = 0;
  e = cudaMemcpy(&h, flag, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpy(error flag)");
  if (h != 0) {
    cudaMemset(flag, 0, sizeof(int));
    return fail(R3M_B200_ERR_KERNEL, "device pipeline watchdog fired, code " + std::to_string(h));
  }
  return R3M_B200_OK;
}

int r3m_b200_conv_fwd(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int R, int S,
                      int stride, int pad, float* stat_sum, float* stat_sq, void* stream) {
  GatherConv g;
  g.src = x;
  g.N = N;
  g.H = H;
  g.W = W;
  g.C = Cin;
  fill_fwd_geometry(&g, R, S, stride, pad);
  g.wpk = w;
  g.Cout = Cout;
  g.out = y;
  g.ldo = Cout;
  g.stat_sum = stat_sum;
  g.stat_sq = stat_sq;
  ConvPlan plan;
  std::string err = plan_conv(g, &plan);
  if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
  cudaError_t e = run_conv(plan, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "conv_fwd launch");
  return R3M_B200_OK;
}

int r3m_b200_conv_fwd_affine(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int R, int S,
                             int stride, int pad, const float* scale, const float* shift, const void* residual,
                             int relu, void* stream) {
  if (!scale || !shift) return fail(R3M_B200_ERR_INVALID, "conv_fwd_affine: scale and shift are required");
  GatherConv g;
  g.src = x;
  g.N = N;
  g.H = H;
  g.W = W;
  g.C = Cin;
  fill_fwd_geometry(&g, R, S, stride, pad);
  g.wpk = w;
  g.Cout = Cout;
  g.out = y;
  g.ldo = Cout;
  g.ep_scale = scale;
  g.ep_shift = shift;
  g.ep_res = residual;
  g.ep_relu = relu;
  ConvPlan plan;
  std::string err = plan_conv(g, &plan);
  if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
  cudaError_t e = run_conv(plan, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "conv_fwd_affine launch");
  return R3M_B200_OK;
}

int r3m_b200_pack_dgrad_filter(const float* w, void* w_dgrad, int Cout, int R, int S, int Cin, int stride, int pad,
                               void* stream) {
  // H, W only set the class extents, which the packing does not depend on
  std::vector<DgradClass> cls = dgrad_classes(2 * stride, 2 * stride, R, S, stride, pad);
  size_t off = 0;
  for (const DgradClass& c : cls) {
    if (c.ntaps == 0) continue;
    int src[kMaxTaps];
    for (int t = 0; t < c.ntaps; ++t) src[t] = c.src_r[t] * S + c.src_s[t];
    cudaError_t e = launch_pack_dgrad(w, reinterpret_cast<uint16_t*>(w_dgrad) + off, Cout, R * S, Cin, c.ntaps, src,
                                      (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "pack_dgrad launch");
    off += (size_t)Cin * c.ntaps * Cout;
  }
  return R3M_B200_OK;
}

int r3m_b200_conv_dgrad(const void* dy, const void* w_dgrad, void* dx, int N, int H, int W, int Cin, int Cout, int R,
                        int S, int stride, int pad, int accumulate, void* stream) {
  const int P = (H + 2 * pad - R) / stride + 1;
  const int Q = (W + 2 * pad - S) / stride + 1;
  std::vector<DgradClass> cls = dgrad_classes(H, W, R, S, stride, pad);
  bool has_empty = false;
  for (const DgradClass& c : cls) has_empty |= (c.ntaps == 0);
  if (has_empty && !accumulate) {
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)N * H * W * Cin * 2, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "dgrad memset");
  }
  size_t off = 0;
  for (const DgradClass& c : cls) {
    if (c.ntaps == 0) continue;
    GatherConv g;
    g.src = dy;
    g.N = N;
    g.H = P;
    g.W = Q;
    g.C = Cout;
    g.P = c.Pc;
    g.Q = c.Qc;
    g.stride = 1;
    g.base_h = c.base_h;
    g.base_w = c.base_w;
    g.ntaps = c.ntaps;
    for (int t = 0; t < c.ntaps; ++t) {
      g.tap_h[t] = c.tap_h[t];
      g.tap_w[t] = c.tap_w[t];
    }
    g.wpk = reinterpret_cast<const uint16_t*>(w_dgrad) + off;
    g.Cout = Cin;
    g.out = dx;
    g.ldo = Cin;
    if (stride > 1) {
      g.out_mode = 1;
      g.oH = H;
      g.oW = W;
      g.o_stride = stride;
      g.o_h0 = c.ph;
      g.o_w0 = c.pw;
    }
    g.accumulate = accumulate;
    ConvPlan plan;
    std::string err = plan_conv(g, &plan);
    if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
    cudaError_t e = run_conv(plan, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(e, "conv_dgrad launch");
    off += (size_t)Cin * c.ntaps * Cout;
  }
  return R3M_B200_OK;
}

int r3m_b200_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad, void* stream) {
  WgradDesc d;
  d.dy = dy;
  d.x = x;
  d.N = N;
  d.H = H;
  d.W = W;
  d.C = Cin;
  fill_fwd_geometry(&d, R, S, stride, pad);
  d.Cout = Cout;
  d.dw = dw;
  WgradPlan plan;
  std::string err = plan_wgrad(d, &plan);
  if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
  cudaError_t e = run_wgrad(plan, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "conv_wgrad launch");
  return R3M_B200_OK;
}


#define CUDA_OR_FAIL(expr, what)                         \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return fail_cuda(_e, what);   \
    return R3M_B200_OK;                                  \
  } while (0)

int r3m_b200_preprocess_stem(const float* obs, void* xs, int N, void* stream) {
  if (!obs || !xs || N < 1) return fail(R3M_B200_ERR_INVALID, "preprocess_stem: bad arguments");
  CUDA_OR_FAIL(launch_preprocess_stem(obs, kObsF32NCHW, xs, N, (cudaStream_t)stream), "preprocess_stem");
}

int r3m_b200_preprocess_stem_format(const void* obs, int format, void* xs, int N, void* stream) {
  if (!obs || !xs || N < 1) return fail(R3M_B200_ERR_INVALID, "preprocess_stem: null pointer or N < 1");
  CUDA_OR_FAIL(launch_preprocess_stem(obs, format, xs, N, (cudaStream_t)stream), "preprocess_stem");
  return R3M_B200_OK;
}

int r3m_b200_random_resized_crop(const uint8_t* src, int nhwc, int N, int H, int W, const int* boxes, float* out,
                                 void* stream) {
  if (!src || !boxes || !out) return fail(R3M_B200_ERR_INVALID, "random_resized_crop: null pointer");
  CUDA_OR_FAIL(launch_crop_resize(src, nhwc, boxes, out, N, H, W, (cudaStream_t)stream), "crop_resize");
  return R3M_B200_OK;
}

int r3m_b200_ordered_sum(const float* x, size_t n, float* out, int blocks, void* stream) {
  if (!x || !out) return fail(R3M_B200_ERR_INVALID, "ordered_sum: null pointer");
  CUDA_OR_FAIL(launch_ordered_sum(x, n, out, blocks, (cudaStream_t)stream), "ordered_sum");
  return R3M_B200_OK;
}

int r3m_b200_ordered_moments(const float* x, int n, float* out, int blocks, void* stream) {
  if (!x || !out || n < 1) return fail(R3M_B200_ERR_INVALID, "ordered_moments: null pointer or empty input");
  CUDA_OR_FAIL(launch_ordered_moments(x, n, out, blocks, (cudaStream_t)stream), "ordered_moments");
  return R3M_B200_OK;
}

int r3m_b200_bn_apply(const void* y, void* a, const void* residual, int M, int C, int relu, int train, const float* sum,
                      const float* sq, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      float* save_mean, float* save_rstd, uint8_t* mask_out, const void* y2, const float* sum2,
                      const float* sq2, const float* gamma2, const float* beta2, float* running_mean2,
                      float* running_var2, float* save_mean2, float* save_rstd2, void* stream) {
  BnApplyArgs g;
  g.y = y;
  g.a = a;
  g.residual = residual;
  g.M = M;
  g.C = C;
  g.relu = relu;
  g.train = train;
  g.sum = sum;
  g.sq = sq;
  g.gamma = gamma;
  g.beta = beta;
  g.running_mean = running_mean;
  g.running_var = running_var;
  g.save_mean = save_mean;
  g.save_rstd = save_rstd;
  g.mask_out = mask_out;
  g.y2 = y2;
  g.sum2 = sum2;
  g.sq2 = sq2;
  g.gamma2 = gamma2;
  g.beta2 = beta2;
  g.running_mean2 = running_mean2;
  g.running_var2 = running_var2;
  g.save_mean2 = save_mean2;
  g.save_rstd2 = save_rstd2;
  CUDA_OR_FAIL(launch_bn_apply(g, (cudaStream_t)stream), "bn_apply");
}

int r3m_b200_bn_backward(const void* dA, const void* a, const uint8_t* mask, const void* y, int M, int C,
                         const float* mean, const float* rstd, const float* gamma, float* sums, void* dy, void* dz,
                         float* dgamma, float* dbeta, const void* y2, const float* mean2, const float* rstd2,
                         const float* gamma2, float* sums2, void* dy2, float* dgamma2, float* dbeta2, void* stream) {
  if (a && mask) return fail(R3M_B200_ERR_INVALID, "bn_backward: give the ReLU mask as `a` or as `mask`, not both");
  BnBwdArgs g;
  g.dA = dA;
  g.a = a;
  g.mask = mask;
  g.y = y;
  g.M = M;
  g.C = C;
  g.mean = mean;
  g.rstd = rstd;
  g.gamma = gamma;
  g.sums = sums;
  g.dy = dy;
  g.dz_out = dz;
  g.dgamma = dgamma;
  g.dbeta = dbeta;
  g.y2 = y2;
  g.mean2 = mean2;
  g.rstd2 = rstd2;
  g.gamma2 = gamma2;
  g.sums2 = sums2;
  g.dy2 = dy2;
  g.dgamma2 = dgamma2;
  g.dbeta2 = dbeta2;
  cudaError_t e = launch_bn_bwd_reduce(g, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(e, "bn_bwd_reduce");
  CUDA_OR_FAIL(launch_bn_bwd_apply(g, (cudaStream_t)stream), "bn_bwd_apply");
}

int r3m_b200_stem_bn_relu_maxpool(const void* y, void* a, uint8_t* argmax, void* ymax, int N, int H, int W, int C, int train,
                                  const float* sum, const float* sq, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float* save_mean, float* save_rstd,
                                  void* stream) {
  StemPoolArgs g;
  g.y = y;
  g.a = a;
  g.argmax = argmax;
  g.ymax = ymax;
  g.N = N;
  g.H = H;
  g.W = W;
  g.C = C;
  g.train = train;
  g.sum = sum;
  g.sq = sq;
  g.gamma = gamma;
  g.beta = beta;
  g.running_mean = running_mean;
  g.running_var = running_var;
  g.save_mean = save_mean;
  g.save_rstd = save_rstd;
  CUDA_OR_FAIL(launch_stem_bn_relu_maxpool(g, (cudaStream_t)stream), "stem_bn_relu_maxpool");
}

int r3m_b200_maxpool_backward(const void* dA, const void* a, const uint8_t* argmax, void* dz, int N, int H, int W, int C,
                              void* stream) {
  (void)a;  // the argmax codes mark ReLU-clipped maxima themselves (code 15); kept in the signature for ABI stability
  CUDA_OR_FAIL(launch_maxpool_bwd(dA, argmax, dz, N, H, W, C, (cudaStream_t)stream), "maxpool_backward");
}

int r3m_b200_stem_backward(const void* dA, const uint8_t* argmax, const void* ymax, const void* y, int N, int H, int W, int C,
                           const float* mean, const float* rstd, const float* gamma, float* sums, void* dy,
                           float* dgamma, float* dbeta, void* stream) {
  if (!dA || !argmax || !y || !mean || !rstd || !gamma || !sums || !dy)
    return fail(R3M_B200_ERR_INVALID, "stem_backward: null argument");
  StemBwdArgs g;
  g.dA = dA;
  g.argmax = argmax;
  g.ymax = ymax;
  g.y = y;
  g.N = N;
  g.H = H;
  g.W = W;
  g.C = C;
  g.mean = mean;
  g.rstd = rstd;
  g.gamma = gamma;
  g.sums = sums;
  g.dy = dy;
  g.dgamma = dgamma;
  g.dbeta = dbeta;
  CUDA_OR_FAIL(launch_stem_bwd(g, (cudaStream_t)stream), "stem_backward");
}

int r3m_b200_pull_host(const void* host_pinned, void* dst, size_t bytes, void* stream) {
  if (!host_pinned || !dst) return fail(R3M_B200_ERR_INVALID, "pull_host: null argument");
  CUDA_OR_FAIL(launch_pull_host(host_pinned, dst, bytes, (cudaStream_t)stream), "pull_host");
}

int r3m_b200_avgpool_forward(const void* a, float* out, int N, int HW, int C, void* stream) {
  CUDA_OR_FAIL(launch_avgpool_fwd(a, out, N, HW, C, (cudaStream_t)stream), "avgpool_forward");
}
int r3m_b200_avgpool_backward(const float* dE, void* dA, int N, int HW, int C, void* stream) {
  CUDA_OR_FAIL(launch_avgpool_bwd(dE, dA, N, HW, C, (cudaStream_t)stream), "avgpool_backward");
}

int r3m_b200_loss_lp(const float* E, float* dE, int rows, int D, float l2weight, float l1weight, float* metrics,
                     void* stream) {
  CUDA_OR_FAIL(launch_loss_lp(E, dE, rows, D, l2weight, l1weight, metrics, (cudaStream_t)stream), "loss_lp");
}
int r3m_b200_loss_tcn(const float* E, float* dE, const int* perms, int B, int D, float tcnweight, float* metrics,
                      void* stream) {
  CUDA_OR_FAIL(launch_loss_tcn(E, dE, perms, B, D, tcnweight, 1, metrics, (cudaStream_t)stream), "loss_tcn");
}
int r3m_b200_loss_tcn_sim(const float* E, float* dE, const int* perms, int B, int D, float tcnweight, int l2dist,
                          float* metrics, void* stream) {
  CUDA_OR_FAIL(launch_loss_tcn(E, dE, perms, B, D, tcnweight, l2dist, metrics, (cudaStream_t)stream), "loss_tcn");
}

int r3m_b200_adam(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n, float lr, int step,
                  float grad_scale, void* stream) {
  if (step < 1) return fail(R3M_B200_ERR_INVALID, "Adam step count starts at 1");
  CUDA_OR_FAIL(launch_adam(p, g, m, v, p_bf16, n, lr, 0.9f, 0.999f, 1e-8f, step, grad_scale, (cudaStream_t)stream),
               "adam");
}

// ----------------------------------------------------------------------------------------------------------------
// engine
// ----------------------------------------------------------------------------------------------------------------
#define ENGINE_OR_FAIL(h)                                                        \
  Engine* eng = reinterpret_cast<Engine*>(h);                                    \
  if (!eng) return fail(R3M_B200_ERR_INVALID, "null engine handle")
#define RETURN_STR(expr)                                                         \
  do {                                                                           \
    std::string _e = (expr);                                                     \
    if (!_e.empty()) return fail(R3M_B200_ERR_STATE, _e);                        \
    return R3M_B200_OK;                                                          \
  } while (0)

int r3m_b200_engine_create(int size, int frames, int lang_head, int hidden_dim, void** handle) {
  if (!handle) return fail(R3M_B200_ERR_INVALID, "null handle pointer");
  Engine* e = nullptr;
  std::string err = Engine::create(size, frames, lang_head, hidden_dim, &e);
  if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
  *handle = e;
  return R3M_B200_OK;
}
int r3m_b200_engine_destroy(void* handle) {
  delete reinterpret_cast<Engine*>(handle);
  return R3M_B200_OK;
}
int r3m_b200_engine_workspace_bytes(void* handle, size_t* bytes) {
  ENGINE_OR_FAIL(handle);
  *bytes = eng->workspace_bytes();
  return R3M_B200_OK;
}
int r3m_b200_engine_param_block_bytes(void* handle, size_t* bytes) {
  ENGINE_OR_FAIL(handle);
  *bytes = eng->param_block_bytes();
  return R3M_B200_OK;
}
int r3m_b200_engine_bind(void* handle, void* param_block, size_t param_bytes, void* workspace, size_t bytes,
                         void* stream) {
  ENGINE_OR_FAIL(handle);
  RETURN_STR(eng->bind(param_block, param_bytes, workspace, bytes, (cudaStream_t)stream));
}
int r3m_b200_engine_num_tensors(void* handle, int* count) {
  ENGINE_OR_FAIL(handle);
  *count = (int)eng->tensors().size();
  return R3M_B200_OK;
}
int r3m_b200_engine_tensor_info(void* handle, int index, char* name, int name_capacity, int* kind, long long* offset,
                                int* ndim, int* dims4) {
  ENGINE_OR_FAIL(handle);
  if (index < 0 || index >= (int)eng->tensors().size()) return fail(R3M_B200_ERR_INVALID, "tensor index out of range");
  const TensorInfo& t = eng->tensors()[index];
  if ((int)t.name.size() + 1 > name_capacity) return fail(R3M_B200_ERR_INVALID, "name buffer too small");
  std::snprintf(name, name_capacity, "%s", t.name.c_str());
  *kind = t.kind;
  *offset = (long long)t.offset;
  *ndim = t.ndim;
  for (int i = 0; i < 4; ++i) dims4[i] = t.dims[i];
  return R3M_B200_OK;
}
int r3m_b200_engine_region(void* handle, int which, void** ptr, size_t* count) {
  ENGINE_OR_FAIL(handle);
  void* p = eng->region(which);
  if (!p) return fail(R3M_B200_ERR_STATE, "unknown region, or no workspace bound");
  *ptr = p;
  if (which <= 3) *count = eng->num_params();
  else if (which == 4) *count = eng->num_buffer_floats();
  else if (which == 5 || which == 6) *count = (size_t)eng->frames() * eng->embed_dim();
  else *count = 16;
  return R3M_B200_OK;
}
int r3m_b200_engine_get_int(void* handle, int what, int* value) {
  ENGINE_OR_FAIL(handle);
  switch (what) {
    case 0: *value = eng->embed_dim(); break;
    case 1: *value = eng->frames(); break;
    case 2: *value = eng->launches_last_call(); break;
    case 3: *value = eng->graph_replays(); break;
    default: return fail(R3M_B200_ERR_INVALID, "unknown query");
  }
  return R3M_B200_OK;
}
int r3m_b200_engine_set_int(void* handle, int what, int value) {
  ENGINE_OR_FAIL(handle);
  switch (what) {
    case 0: eng->set_l2dist(value != 0); break;
    case 1: {
      std::string err = eng->set_obs_format(value);
      if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
      break;
    }
    case 2: {
      std::string err = eng->set_precision(value);
      if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
      break;
    }
    default: return fail(R3M_B200_ERR_INVALID, "unknown setting");
  }
  return R3M_B200_OK;
}
int r3m_b200_engine_param_block_layout(void* handle, size_t* offsets5, size_t* num_params, size_t* num_buffer_floats) {
  ENGINE_OR_FAIL(handle);
  eng->param_block_layout(offsets5);
  *num_params = eng->num_params();
  *num_buffer_floats = eng->num_buffer_floats();
  return R3M_B200_OK;
}
int r3m_b200_engine_sync_weights(void* handle, void* stream) {
  ENGINE_OR_FAIL(handle);
  RETURN_STR(eng->sync_weights((cudaStream_t)stream));
}
int r3m_b200_engine_forward(void* handle, const void* obs, int train, float* out, void* stream) {
  ENGINE_OR_FAIL(handle);
  if (!obs) return fail(R3M_B200_ERR_INVALID, "null observation pointer");
  RETURN_STR(eng->forward(obs, train, out, (cudaStream_t)stream));
}
int r3m_b200_engine_update_grads(void* handle, const void* obs, const int* perms, const float* lang_emb,
                                 const float* lang_mask, float l2weight, float l1weight, float langweight,
                                 float tcnweight, int eval, void* stream) {
  ENGINE_OR_FAIL(handle);
  if (!perms) return fail(R3M_B200_ERR_INVALID, "null permutation pointer");
  Hyper h;
  h.l2weight = l2weight;
  h.l1weight = l1weight;
  h.langweight = langweight;
  h.tcnweight = tcnweight;
  RETURN_STR(eng->update_grads(obs, perms, lang_emb, lang_mask, h, eval, (cudaStream_t)stream));
}
int r3m_b200_engine_profile_update(void* handle, const void* obs, const int* perms, const float* lang_emb,
                                   const float* lang_mask, float l2weight, float l1weight, float langweight,
                                   float tcnweight, float lr, int step, double* out32, void* stream) {
  ENGINE_OR_FAIL(handle);
  if (!obs || !perms || !out32) return fail(R3M_B200_ERR_INVALID, "null pointer");
  Hyper h;
  h.l2weight = l2weight;
  h.l1weight = l1weight;
  h.langweight = langweight;
  h.tcnweight = tcnweight;
  RETURN_STR(eng->profile_update(obs, perms, lang_emb, lang_mask, h, lr, step, out32, (cudaStream_t)stream));
}
int r3m_b200_engine_profile_ops(void* handle, double* out, int capacity_ops, int* num_ops) {
  ENGINE_OR_FAIL(handle);
  const std::vector<double>& v = eng->last_profile_ops();
  const int n = (int)(v.size() / 4);
  *num_ops = n;
  for (int i = 0; i < n && i < capacity_ops; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = v[4 * i + j];
  return R3M_B200_OK;
}
int r3m_b200_engine_profile_label(void* handle, int index, char* out, int capacity) {
  ENGINE_OR_FAIL(handle);
  const std::vector<std::string>& v = eng->last_profile_labels();
  if (index < 0 || index >= (int)v.size()) return fail(R3M_B200_ERR_INVALID, "profile index out of range");
  std::snprintf(out, capacity, "%s", v[index].c_str());
  return R3M_B200_OK;
}
int r3m_b200_engine_backward(void* handle, const float* dE, void* stream) {
  ENGINE_OR_FAIL(handle);
  if (!dE) return fail(R3M_B200_ERR_INVALID, "null gradient pointer");
  RETURN_STR(eng->backward(dE, (cudaStream_t)stream));
}
int r3m_b200_engine_num_grad_chunks(void* handle, int* count) {
  ENGINE_OR_FAIL(handle);
  *count = eng->num_grad_chunks();
  return R3M_B200_OK;
}
int r3m_b200_engine_grad_chunk(void* handle, int k, size_t* begin, size_t* end) {
  ENGINE_OR_FAIL(handle);
  if (!begin || !end) return fail(R3M_B200_ERR_INVALID, "null pointer");
  RETURN_STR(eng->grad_chunk(k, begin, end));
}
int r3m_b200_engine_wait_grad_chunk(void* handle, int k, void* stream) {
  ENGINE_OR_FAIL(handle);
  RETURN_STR(eng->wait_grad_chunk(k, (cudaStream_t)stream));
}
int r3m_b200_engine_num_blocks(void* handle, int* count) {
  ENGINE_OR_FAIL(handle);
  *count = eng->num_blocks();
  return R3M_B200_OK;
}
int r3m_b200_engine_debug_block(void* handle, int block, int what, void** ptr, size_t* count) {
  ENGINE_OR_FAIL(handle);
  if (!ptr || !count) return fail(R3M_B200_ERR_INVALID, "null pointer");
  RETURN_STR(eng->debug_block(block, what, ptr, count));
}
int r3m_b200_engine_debug_run_block_backward(void* handle, int block, void* stream) {
  ENGINE_OR_FAIL(handle);
  RETURN_STR(eng->debug_run_block_backward(block, (cudaStream_t)stream));
}
int r3m_b200_engine_adam_step(void* handle, float lr, float grad_scale, int step, void* stream) {
  ENGINE_OR_FAIL(handle);
  RETURN_STR(eng->adam_step(lr, grad_scale, step, (cudaStream_t)stream));
}

}  // extern "C"
