// Host-side planning of the tcgen05 implicit-GEMM launches: forward conv, dgrad (as gather convs over dY) and wgrad.
#pragma once
#include <string>
#include <vector>

#include "conv_igemm.cuh"

namespace r3m {

// A "gather conv": out[m, k] = sum_{t, c} src[pixel(m) + tap_t, c] * wpk[k, t, c], m = (n, p, q) flattened, where
// pixel(m) = (base_h + p*stride, base_w + q*stride) and taps are non-negative offsets.  Out-of-range source pixels
// read as zero.  Forward convolutions and every dgrad parity class are instances of this.
struct GatherConv {
  const void* src = nullptr;  // bf16 NHWC, C % 64 == 0
  int N = 0, H = 0, W = 0, C = 0;
  int P = 0, Q = 0, stride = 1, base_h = 0, base_w = 0;
  int ntaps = 0;
  int tap_h[kMaxTaps] = {0}, tap_w[kMaxTaps] = {0};
  const void* wpk = nullptr;  // bf16 [Cout][ntaps * C]
  int Cout = 0;
  void* out = nullptr;  // bf16
  int out_mode = 0;     // 0: dense rows [M][ldo];  1: scattered to an (oH, oW) grid at (p*o_stride+o_h0, q*o_stride+o_w0)
  int ldo = 0;
  int oH = 0, oW = 0, o_stride = 1, o_h0 = 0, o_w0 = 0;
  int accumulate = 0;  // out += result (read-modify-write)
  int rev_m = 0;       // walk the M tiles last-to-first (see ConvKernelParams)
  int bn = 0;          // N tile override (64 / 128 / 256; 0: widest that divides Cout): few-row GEMMs want more, narrower tiles
  float* stat_sum = nullptr;  // optional per-channel sum / sum of squares of the stored (bf16) output (written)
  float* stat_sq = nullptr;
  // deterministic two-level reduction of the statistics: [kStatScratchFloats] floats + [kStatTickets] zeroed ints; null:
  // a process-wide buffer (launches that share it must be stream-ordered)
  float* stat_scratch = nullptr;
  int* stat_ticket = nullptr;
  int stat_raw = 0;  // 1: stat_sum is the layer's raw fixed-point accumulator block (see ConvKernelParams)
  int stat_rows = 0;  // > 0 (ticket mode): publish (mean, variance) over stat_rows values instead of (sum, sum of squares)
  // fused inference epilogue (see ConvKernelParams): per-channel affine (+ residual) (+ ReLU) on the accumulators
  const float* ep_scale = nullptr;
  const float* ep_shift = nullptr;
  const void* ep_res = nullptr;
  int ep_relu = 0;   // 0 none, 1 ReLU, 2 exact GELU (tf32 tier only)
  int ep_exact = 0;  // tf32 tier: keep the fp32 result unrounded (see ConvKernelParams)
  // tf32 tier (inference parity): src / wpk / out / ep_res are fp32 (tf32-rounded values), C % 32 == 0; dense
  // non-accumulating outputs without statistics only
  int tf32 = 0;
};

struct ConvPlan {
  CUtensorMap tmA, tmB, tmC;  // im2col activations, filter, dense output (unused for scatter outputs)
  ConvKernelParams p;
  int bn = 0;
  int grid = 0;
  // 3x3 / stride 1 / 64 -> 64 channels: the tap-reuse kernel (halo3x3.cu); tmA = rank-4 tiled halo load, tmB = filter K
  // blocks, tmC = rank-4 tiled store
  int halo = 0;
  HaloKernelParams hp;
};

struct WgradDesc {
  const void* dy = nullptr;  // bf16 [M][Cout], M = N*P*Q
  const void* x = nullptr;   // bf16 NHWC source of the forward conv
  int N = 0, H = 0, W = 0, C = 0;
  int P = 0, Q = 0, stride = 1, base_h = 0, base_w = 0;
  int ntaps = 0;
  int tap_h[kMaxTaps] = {0}, tap_w[kMaxTaps] = {0};
  int Cout = 0;
  float* dw = nullptr;  // fp32 [Cout][ntaps * C], accumulated into (caller zeroes)
  // split-K scratch (kWgradScratchBytes); null: a process-wide buffer (launches that share it must be stream-ordered)
  float* scratch = nullptr;
};

struct WgradPlan {
  CUtensorMap tmDy, tmX;
  WgradKernelParams p;
  int splits = 0, groups = 0, ktiles = 0;
};

constexpr size_t kStatScratchFloats = 160 * 512;  // grid (<= SM count) x 2 x BN (<= 256)
constexpr size_t kStatTickets = 64;
constexpr size_t kWgradScratchBytes = 40u << 20;  // >= splits * |dW| * 4 for every launch (one wave of <= 164 KB slabs)
float* device_wgrad_scratch();
int device_sm_count();
float* device_stat_scratch();  // lazily allocated process-wide scratch: kStatScratchFloats floats + kStatTickets ints
int* device_error_flag();  // lazily allocated device int, zero-initialised

std::string plan_conv(const GatherConv& g, ConvPlan* plan);
cudaError_t run_conv(const ConvPlan& plan, cudaStream_t stream);

std::string plan_wgrad(const WgradDesc& d, WgradPlan* plan);
cudaError_t run_wgrad(const WgradPlan& plan, cudaStream_t stream);

// Geometry helpers ------------------------------------------------------------------------------------------------
// Forward conv of an R x S filter, stride, pad over an H x W input.
void fill_fwd_geometry(GatherConv* g, int R, int S, int stride, int pad);
void fill_fwd_geometry(WgradDesc* d, int R, int S, int stride, int pad);

// dgrad of a forward conv (R, S, stride in {1,2}, pad) whose INPUT is H x W and output P x Q: decomposed into
// stride*stride parity classes.  Class (ph, pw) computes dX[n, stride*i+ph, stride*j+pw, :] as a gather conv over dY.
struct DgradClass {
  int ph = 0, pw = 0;
  int Pc = 0, Qc = 0;          // base grid of the class (rows/cols of dX with that parity)
  int base_h = 0, base_w = 0;  // lower corner in dY coordinates
  int ntaps = 0;
  int tap_h[kMaxTaps], tap_w[kMaxTaps];  // offsets in dY
  int src_r[kMaxTaps], src_s[kMaxTaps];  // forward filter tap feeding each dgrad tap
};
std::vector<DgradClass> dgrad_classes(int H, int W, int R, int S, int stride, int pad);

}  // namespace r3m
