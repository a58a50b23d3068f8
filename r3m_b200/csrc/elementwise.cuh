// Launch interface of the HBM-bound kernels around the convolutions (NHWC bf16 activations, fp32 statistics).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace r3m {

constexpr float kBnEps = 1e-5f;       // torch.nn.BatchNorm2d default (tv resnet.py: norm_layer = nn.BatchNorm2d)
constexpr float kBnMomentum = 0.1f;

// obs fp32 NCHW [N,3,224,224] in [0,255]  ->  stem operand bf16 [N,112,112,64]:
//   channel j = kw*16 + (dy*2+dx)*4 + c  holds  normalise(obs[n, c, 2*i+dy, 2*(q-2+kw)+dx])   (0 outside / c == 3)
// ref: r3m/models/models_r3m.py:97-98 (x/255, Normalize(mean,std)) fused with the space-to-depth re-layout.
// The frames may arrive in three formats (obs is [N,3,224,224] or, for kObsU8NHWC, [N,224,224,3]).
enum ObsFormat { kObsF32NCHW = 0, kObsU8NCHW = 1, kObsU8NHWC = 2 };
cudaError_t launch_preprocess_stem(const void* obs, int format, void* xs, int N, cudaStream_t s);

// RandomResizedCrop's arithmetic (r3m/utils/data_loaders.py:47-50): per-frame crop box (top, left, h, w) of a uint8
// frame [N,3,H,W] (nhwc: [N,H,W,3]) -> fp32 NCHW [N,3,224,224] in [0,255] by antialiased bilinear interpolation
// (ATen upsample_bilinear2d_aa semantics, borders clamped to the crop).  boxes: device int32 [N][4].
cudaError_t launch_crop_resize(const uint8_t* src, int nhwc, const int* boxes, float* out, int N, int H, int W,
                               cudaStream_t s);

struct BnApplyArgs {
  const void* y = nullptr;      // bf16 [M][C] raw conv output
  void* a = nullptr;            // bf16 [M][C] activated output
  const void* residual = nullptr;  // bf16 [M][C] or null
  int M = 0, C = 0;
  int relu = 1;
  int train = 1;                // 1: batch statistics from (sum, sq);  0: running statistics
  const float* sum = nullptr;   // [C]
  const float* sq = nullptr;    // [C]
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float* running_mean = nullptr;
  float* running_var = nullptr;
  float* save_mean = nullptr;   // [C] written in train mode (for backward)
  float* save_rstd = nullptr;
  int update_running = 1;
  // 1: sum (== sq) and sum2 (== sq2) point at the producing conv's raw fixed-point accumulators (8 64-bit words per
  // channel: 4 limbs of the sum, 4 of the sum of squares; see fx_add) instead of fp32 arrays — the engine's path.
  // 2: the same, with the round-1 fp32 moments arithmetic (A/B aid, R3M_BN_FP32=1; see bn_batch_moments)
  int stat_raw = 0;
  // 1: walk the rows last-to-first.  The engine alternates the traversal direction of consecutive passes over a tensor:
  // the rows a producer wrote LAST are still in the 126 MB L2 when its consumer starts from that end
  int reverse = 0;
  uint8_t* mask_out = nullptr;  // optional [M][C/8]: bit j of byte (row, chunk) = (a[row][8*chunk+j] > 0)
  // optional second BatchNorm whose (un-activated) output is added before the ReLU: the downsample branch of a
  // residual block, a = relu(bn(y) + bn2(y2))  (tv resnet.py:100-103, 155-161) without materialising bn2(y2)
  const void* y2 = nullptr;
  const float* sum2 = nullptr;
  const float* sq2 = nullptr;
  const float* gamma2 = nullptr;
  const float* beta2 = nullptr;
  float* running_mean2 = nullptr;
  float* running_var2 = nullptr;
  float* save_mean2 = nullptr;
  float* save_rstd2 = nullptr;
};
cudaError_t launch_bn_apply(const BnApplyArgs& a, cudaStream_t s);

// Scratch of the deterministic grid-wide reductions (see det_grid_reduce in elementwise.cu): `scratch` holds the
// fixed-point accumulators (four 64-bit words per reduced value), `tickets` one int per channel slice; both zero on
// entry and left zero on exit.  Launches that share one DetScratch must be stream-ordered.
struct DetScratch {
  float* scratch = nullptr;
  int* tickets = nullptr;
};
constexpr int kDetMaxBlocks = 1184;                             // grid cap of the kernels that reduce through it
constexpr size_t kDetScratchFloats = (size_t)3 * 2048 * 8;      // 3 * C accumulators of 32 bytes, C <= 2048
constexpr size_t kDetTickets = 128;
DetScratch device_det_scratch();  // lazily allocated process-wide instance (kernel-level C-ABI entry points)

// stem: y [N,112,112,64] -> BN -> ReLU -> maxpool 3x3 s2 p1 -> a [N,56,56,64], argmax code (0..8) per element
struct StemPoolArgs {
  const void* y = nullptr;
  void* a = nullptr;
  uint8_t* argmax = nullptr;  // may be null (eval)
  void* ymax = nullptr;       // optional bf16 [N,H/2,W/2,C]: the RAW conv output at the argmax (for the backward reduce)
  int N = 0, H = 112, W = 112, C = 64;
  int train = 1;
  const float* sum = nullptr;
  const float* sq = nullptr;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float* running_mean = nullptr;
  float* running_var = nullptr;
  float* save_mean = nullptr;
  float* save_rstd = nullptr;
  int update_running = 1;
  int stat_raw = 0;  // see BnApplyArgs
};
cudaError_t launch_stem_bn_relu_maxpool(const StemPoolArgs& a, cudaStream_t s);
// dz[n,h,w,c] = sum over pooling windows whose argmax is (h,w) of dA[window]   (dead maxima carry code 15: no match)
cudaError_t launch_maxpool_bwd(const void* dA, const uint8_t* argmax, void* dz, int N, int H, int W, int C,
                               cudaStream_t s);

// Stem backward in two launches: maxpool backward + ReLU mask recomputed from (dA, argmax) inside the BatchNorm
// backward's reduce and apply passes (the masked [N,H,W,C] gradient is never materialised).
struct StemBwdArgs {
  const void* dA = nullptr;        // bf16 [N,H/2,W/2,C] gradient w.r.t. the pooled output
  const uint8_t* argmax = nullptr; // [N,H/2,W/2,C] codes written by launch_stem_bn_relu_maxpool
  // optional bf16 [N,H/2,W/2,C] raw conv output at each window's argmax (same launch): the reduce pass then runs over
  // the pooled elements only (sum dz = sum over live windows of dA, sum dz*xhat likewise with xhat taken at ymax)
  const void* ymax = nullptr;
  const void* y = nullptr;         // bf16 [N,H,W,C] raw stem conv output
  int N = 0, H = 112, W = 112, C = 64;
  const float* mean = nullptr;
  const float* rstd = nullptr;
  const float* gamma = nullptr;
  float* sums = nullptr;           // [2][C] workspace: written by the reduce pass, read by the apply pass
  void* dy = nullptr;              // bf16 [N,H,W,C]
  float* dgamma = nullptr;
  float* dbeta = nullptr;
  DetScratch det;                  // null members: the process-wide instance
  // 1: `sums` is 2C raw fixed-point accumulators (four 64-bit words each, zero on entry): the reduce pass adds into them
  // and the apply pass converts on read — no finalize tail (the engine's path)
  int sums_raw = 0;
};
cudaError_t launch_stem_bwd(const StemBwdArgs& a, cudaStream_t s);

cudaError_t launch_avgpool_fwd(const void* a, float* out, int N, int HW, int C, cudaStream_t s);
cudaError_t launch_avgpool_bwd(const float* dE, void* dA, int N, int HW, int C, cudaStream_t s);

struct BnBwdArgs {
  const void* dA = nullptr;   // bf16 [M][C] gradient w.r.t. the layer's (activated) output
  const void* a = nullptr;    // bf16 [M][C] activated output (ReLU mask) or null when no ReLU follows / already masked
  const uint8_t* mask = nullptr;  // alternative ReLU mask: bit-packed [M][C/8] as written by bn_apply (a must be null)
  const void* y = nullptr;    // bf16 [M][C] raw conv output
  int M = 0, C = 0;
  const float* mean = nullptr;
  const float* rstd = nullptr;
  const float* gamma = nullptr;
  float* sums = nullptr;      // [2][C] workspace: sum(dz), sum(dz * xhat); written by the reduce pass
  DetScratch det;             // deterministic reduction scratch (null members: the process-wide instance)
  int single_rows = 0;  // A/B aid (R3M_AB bit 1): reduce pass without the four-row software pipeline
  int rev_reduce = 0, rev_apply = 0;  // traversal direction of the two passes (see BnApplyArgs::reverse)
  int sums_raw = 0;           // 1: sums / sums2 are raw fixed-point accumulators (2C / C entries of four 64-bit words,
                              // zero on entry); the apply pass converts on read (the engine's path)
  void* dy = nullptr;         // bf16 [M][C] gradient w.r.t. the raw conv output
  void* dz_out = nullptr;     // optional bf16 [M][C]: masked gradient (feeds the residual branch)
  float* dgamma = nullptr;    // [C]
  float* dbeta = nullptr;
  // optional second BatchNorm fed by the same masked gradient (the downsample branch): y2 raw output, dy2 result
  const void* y2 = nullptr;
  const float* mean2 = nullptr;
  const float* rstd2 = nullptr;
  const float* gamma2 = nullptr;
  float* sums2 = nullptr;     // [C] workspace: sum(dz * xhat2); written by the reduce pass
  void* dy2 = nullptr;
  float* dgamma2 = nullptr;
  float* dbeta2 = nullptr;
};
cudaError_t launch_bn_bwd_reduce(const BnBwdArgs& a, cudaStream_t s);
cudaError_t launch_bn_bwd_apply(const BnBwdArgs& a, cudaStream_t s);
// Both passes in one launch with a grid barrier between them (layers that stay L2 resident; see bn_bwd_fused_kernel).
// counter: one int, zero on entry (the engine keeps it in the step's zeroed region); error_flag: the device watchdog
// flag (a lost block sets it instead of hanging the grid).
bool bn_bwd_can_fuse(const BnBwdArgs& a);
cudaError_t launch_bn_bwd_fused(const BnBwdArgs& a, int* counter, int* error_flag, cudaStream_t s);

// Inference: fold BatchNorm (running statistics) into a per-channel affine, for all layers in one launch.
struct BnFoldEntry {
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float* scale;  // gamma / sqrt(var + eps)
  float* shift;  // beta - mean * scale
  int C;
};
cudaError_t launch_bn_fold(const BnFoldEntry* table_dev, int entries, cudaStream_t s);

// Adam (torch.optim.Adam defaults, ref r3m/models/models_r3m.py:76): flat fp32 p/g/m/v, also emits the bf16 copy.
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n, float lr, float beta1,
                        float beta2, float eps, int step, float grad_scale, cudaStream_t s);
cudaError_t launch_cast_bf16(const float* src, void* dst, size_t n, cudaStream_t s);
// dst (device) <- host_mapped (pinned host memory, device-accessible under unified addressing), copied by a kernel so
// that it does not share the H2D copy engine's queue with bulk uploads.  bytes: a multiple of 16; both 16-byte aligned.
cudaError_t launch_pull_host(const void* host_mapped, void* dst, size_t bytes, cudaStream_t s);

// dgrad filter: out[c][t][k] = w[k][src_tap[t]][c]   (w fp32 [Cout][T][Cin] -> bf16 [Cin][nt][Cout])
cudaError_t launch_pack_dgrad(const float* w, void* out, int Cout, int T, int Cin, int nt, const int* src_tap,
                              cudaStream_t s);
// the same for every filter of a network in ONE launch (61 launches of a few microseconds each otherwise); the table
// lives in device memory, block_begin is the running sum of ceil(Cin/32) * ceil(Cout/32) * nt
struct PackDgradEntry {
  const float* w;
  void* out;
  int Cout, T, Cin, nt;
  int taps[16];
  int block_begin;
};
cudaError_t launch_pack_dgrad_multi(const PackDgradEntry* table_dev, int entries, int total_blocks, cudaStream_t s);
// stem filter OIHW fp32 [64,3,7,7] <-> packed [64][4][64] (bf16 forward operand / fp32 gradient)
cudaError_t launch_stem_pack(const float* w_oihw, void* wp_bf16, cudaStream_t s);
cudaError_t launch_stem_unpack_grad(const float* dwp, float* dw_oihw, cudaStream_t s);

// ---- tf32 tier of the inference path: fp32 storage holding tf32-rounded values (see conv_igemm.cu PREC = 1)
cudaError_t launch_round_tf32(const float* src, float* dst, size_t n, cudaStream_t s);
cudaError_t launch_preprocess_stem_f32(const void* obs, int format, float* xs, int N, cudaStream_t s);
cudaError_t launch_maxpool_f32(const float* y, float* a, int N, int H, int W, int C, cudaStream_t s);
cudaError_t launch_avgpool_fwd_f32(const float* a, float* out, int N, int HW, int C, cudaStream_t s);
cudaError_t launch_stem_pack_f32(const float* w_oihw, float* wp, cudaStream_t s);

// out[0] = the EXACT sum of x[0..n) rounded once to fp32, through the fixed-point accumulators every deterministic
// reduction of the step uses (test hook: the result must not depend on `blocks`, i.e. on grouping and arrival order)
// BnApplyArgs / StemPoolArgs::stat_raw of the engine's path: 1, or 2 under R3M_BN_FP32=1
int bn_stat_mode();
cudaError_t launch_ordered_sum(const float* x, size_t n, float* out, int blocks, cudaStream_t s);
cudaError_t launch_ordered_moments(const float* x, int n, float* out, int blocks, cudaStream_t s);

}  // namespace r3m
