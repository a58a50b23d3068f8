// Internal launch interface of the tcgen05 implicit-GEMM kernels (conv_igemm.cu, wgrad.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace r3m {

constexpr int kMaxTaps = 16;

struct ConvKernelParams {
  // GEMM rows: flattened (n, p, q) base positions of the im2col traversal
  int M_total;
  int PQ, Q;
  int stride;          // traversal stride of the base pixel (conv stride)
  int base_w, base_h;  // lower corner of the bounding box (= -pad for a forward conv)
  // GEMM K: taps x 64-channel blocks
  int num_taps;
  int cblocks;
  uint16_t tap_w[kMaxTaps];
  uint16_t tap_h[kMaxTaps];
  // GEMM N
  int Cout;
  int num_m_tiles, num_n_tiles;
  int pair;   // 1: cta_group::2 CTA pairs (clusters of two along grid x, two M tiles per filter tile; tmB's box is BN / 2 rows)
  int rev_m;  // 1: M tiles are walked last-to-first (L2 reuse of the rows the producing kernel wrote last)
  // output
  void* out;     // bf16
  int out_mode;  // 0: row m at m*ldo;  1: row m=(n,p,q) scattered to ((n*oH + p*o_stride+o_h0)*oW + q*o_stride+o_w0)*ldo
  int ldo;
  int oH, oW, o_stride, o_h0, o_w0;
  int accumulate;  // out += result (bf16 read-modify-write)
  // optional per-channel statistics of the (bf16-rounded) output: WRITTEN (not accumulated), deterministically: every
  // CTA adds its column sums to fixed-point accumulators in stat_scratch (Cout * 2 values * 32 bytes, zero on entry and
  // exit), the last CTA of each column block (stat_ticket, one zeroed int per column block, self-resetting) converts
  // them to fp32.  The grid must be a multiple of num_n_tiles.
  float* stat_sum;
  float* stat_sq;
  float* stat_scratch;
  int* stat_ticket;
  int stat_raw;  // 1: stat_sum points at the layer's own accumulators (Cout * 8 64-bit words, zero on entry); the kernel
                 // only adds into them and the consumer converts (no scratch, no ticket, no finalize tail)
  int stat_rows;  // ticket mode (stat_raw == 0) only; > 0: the last CTA of a column block publishes the batch MOMENTS of
                  // stat_rows values per channel — stat_sum[c] = mean, stat_sq[c] = biased variance (fx_moments) —
                  // instead of the fp32-rounded sums
  // optional fused epilogue (inference: BatchNorm folded into a per-channel affine): out = [relu](acc * ep_scale[c] +
  // ep_shift[c] [+ ep_res[m][c]]).  Dense outputs only; mutually exclusive with the statistics.
  const float* ep_scale;
  const float* ep_shift;
  const void* ep_res;  // bf16 [M][ldo] (fp32 in the tf32 tier) or null
  int ep_relu;  // 0 none, 1 ReLU, 2 exact GELU (tf32 tier only)
  int ep_exact;  // tf32 tier: 1 = store the fp32 result as is (consumer is not a tensor-core operand), 0 = round to tf32
  int tf32;  // 1: fp32 operands / output through kind::tf32 (cblocks counts 32-channel blocks); dense mode only
  int* error_flag;
};

cudaError_t conv_igemm_launch(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                              const ConvKernelParams& p, int grid, cudaStream_t stream);

// 3x3 / stride-1 / 64 -> 64 channel convolutions with TAP REUSE (halo3x3.cu): an M tile is a 16 x 8 patch of output pixels,
// ONE tiled TMA load brings its 18 x 10 x 64 input halo, and the nine taps are nine shifted views of that shared memory.
struct HaloKernelParams {
  int N, H, W;            // images; output (= input) height and width: H % 4 == 0, W % 8 == 0
  int tiles_h, tiles_w;   // ceil(H / 16), W / 8
  int num_taps;           // <= 9
  uint16_t tap_w[9], tap_h[9];  // tap t reads input pixel (p + org_h + tap_h[t], q + org_w + tap_w[t]) with filter K block t
  int org_h, org_w;       // halo origin relative to the tile's first output pixel: (-1, -1) for the 3x3 convs, (-2, 0) for
                          // the stem's four vertical taps
  int patch_bytes;        // bytes of one halo load: (16 + max tap_h) rows x 10 columns x 128
  int rev;                // walk the tiles last-to-first
  unsigned long long* stat_acc;  // optional raw fixed-point accumulators (64 channels x {sum, sum of squares}), added into
  int acc_stages;         // accumulator stages in tensor memory (tiles in flight between the MMA issuer and the epilogue)
  int debug;              // experiment switches (R3M_HALO_DEBUG): 1 no epilogue work, 2 one MMA per tile, 4 no patch loads
  int base_offset_mode;   // experiment switch: 1 = descriptors carry the swizzle base offset of their start row
  int* error_flag;
};
cudaError_t halo3x3_launch(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmY,
                           const HaloKernelParams& p, int grid, cudaStream_t stream);

struct WgradKernelParams {
  int M_total;
  int PQ, Q;
  int stride;
  int base_w, base_h;
  int cblocks;      // 64-channel blocks of the activation per tap
  int num_items;    // taps * cblocks   (one item = one 64-wide slice of the filter's K axis)
  int group;        // items per CTA (<= 8); pair mode: items per CTA PAIR (even, <= 8)
  int pair;         // 1: wgrad_pair_kernel (cta_group::2, clusters of two CTAs along the k-tile axis; Cout % 256 == 0)
  uint16_t tap_w[kMaxTaps];
  uint16_t tap_h[kMaxTaps];
  int Cout;         // rows of dW
  int ldw;          // elements per dW row (= taps * Cin)
  int Cin;
  int pix_block;          // reduction rows (pixels) per pipeline stage: 64 or 128
  int mblocks_total;      // ceil(M_total / pix_block)
  int mblocks_per_split;
  int num_stages;
  float* dW;        // fp32, accumulated into
  // split-K (splits > 1): fp32 scratch of splits * dw_elems floats; every split writes its partial dW there and a
  // second kernel adds them to dW in split order (deterministic).  null: one split, added straight into dW.
  float* scratch;
  size_t dw_elems;  // Cout * ldw
  int* error_flag;
};

cudaError_t wgrad_launch(const CUtensorMap& tmDy, const CUtensorMap& tmX, const WgradKernelParams& p, int splits,
                         int groups, int ktiles, cudaStream_t stream);
int wgrad_smem_bytes(int group, int num_stages, int pix_block);

}  // namespace r3m
