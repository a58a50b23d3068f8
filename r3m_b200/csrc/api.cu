// extern "C" boundary (include/r3m_b200.h).  Plain pointers and sizes only; no torch types cross this line.
#include "../../include/r3m_b200.h"

#include <cstdio>
#include <string>
#include <vector>

#include "api_util.h"
#include "convops.h"
#include "elementwise.cuh"

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

}  // extern "C"
