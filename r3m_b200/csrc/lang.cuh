// Video-language alignment head of Trainer.update (reference r3m/trainer.py:63-118, LanguageReward MLP
// r3m/models/models_language.py:37-55).  All 15 get_reward() evaluations of one update are batched into ONE
// [15*B, 2D+768] pass (the reference runs 15 separate MLP calls, 75 small cuBLAS launches).  fp32 throughout: the
// reference's Linear layers run in true fp32 (cuda.matmul.allow_tf32 == False) and the loss tolerance is 1e-4.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "convops.h"

namespace r3m {

struct LangDims {
  int B = 0;   // clips
  int D = 0;   // embedding dim
  int L = 768; // sentence-embedding dim
  int H = 0;   // hidden units
  __host__ __device__ int rows() const { return 15 * B; }
  __host__ __device__ int k1() const { return 2 * D + L; }
};

struct LangParams {  // device pointers into the flat parameter / gradient buffers
  const float* w[5];
  const float* b[5];
  float* dw[5];
  float* db[5];
};

struct LangWorkspace {  // device scratch, sized by lang_workspace_floats()
  // Layer 1 is factorised: cat([e0, e_t, l]) . W1^T = e0 . W1a^T + e_t . W1b^T + l . W1c^T, and only B distinct e0 rows,
  // 5B distinct e_t rows and B distinct sentence rows exist among the 15B evaluations, so the three products are
  // computed once per distinct row (3.4x fewer layer-1 FLOPs than the reference's 15 concatenated passes).
  float* U;        // [B][H]   e0 . W1a^T
  float* V;        // [5B][H]  e  . W1b^T
  float* Lc;       // [B][H]   l  . W1c^T
  float* dU;       // gradients of the three products (gathered from the 15B rows)
  float* dV;
  float* dLc;
  float* Hact[4];  // post-ReLU hidden activations [rows][H]
  float* S;        // [rows] scores
  float* dS;       // [rows]
  float* dH[2];    // ping-pong [rows][H]
  float* splitk;   // slices of the split-K GEMMs (summed in slice order: deterministic)
};
// Tensor-core path of the three hidden Linear(H, H) layers (forward, dX and dW: nine 15B x H x H products, 85 % of the
// head's FLOPs).  The reference computes them in true fp32, so every fp32 operand is split into two tf32 terms
// (x = hi + lo) and a . w ~= a_hi . w_hi + a_lo . w_hi + a_hi . w_lo runs as ONE tcgen05 kind::tf32 GEMM over a
// three-times-longer reduction axis with fp32 accumulation (the conv kernel's tf32 tier as a plain GEMM: bias and ReLU
// in its epilogue) — error ~1e-6 relative, the same parity band as the fp32 SIMT path.  Small fused kernels produce the
// split (and transposed split) operand copies and apply the ReLU gates.
struct LangTc {
  bool enabled = false;
  int Rp = 0;            // 15B rounded up to a multiple of 32 (reduction axis of the dW products, zero padded)
  ConvPlan fwd[3];       // H_l = relu(H_{l-1} W_l^T + b_l),   l = 1..3  (index l - 1)
  ConvPlan dx[3];        // dH_{l-1} (before its ReLU gate) = dH_l W_l
  ConvPlan dw[3];        // dW_l = dH_l^T H_{l-1}
  float* Hs[3];          // split H_{l-1}                [15B][3H]  (hi | lo | hi)
  float* HTs[3];         // split H_{l-1}^T              [H][3Rp]   (hi | hi | lo)
  float* dHs;            // split dH_l                   [15B][3H]  (hi | lo | hi)
  float* dHTs;           // split dH_l^T                 [H][3Rp]   (hi | lo | hi)
  float* Ws[3];          // split W_l                    [H][3H]    (hi | hi | lo)
  float* WTs[3];         // split W_l^T                  [H][3H]    (hi | hi | lo)
  float* ones;           // [H]
};
size_t lang_tc_floats(const LangDims& d);
// Carves `base` (lang_tc_floats floats, 256-byte aligned) and encodes the nine GEMM plans against the head's parameter /
// gradient / workspace pointers.  Returns an empty string on success.
std::string lang_tc_plan(const LangDims& d, const LangParams& p, const LangWorkspace& ws, float* base, LangTc* tc);

size_t lang_workspace_floats(const LangDims& d);
void lang_carve_workspace(float* base, const LangDims& d, LangWorkspace* ws);

// Forward + InfoNCE loss (+ metrics).  When dE != null also the full backward: parameter gradients are WRITTEN to
// p.dw / p.db and d(langw * rewloss)/dE is accumulated into dE (ordered reductions only: no atomics).  Returns the number of kernels launched
// through *launches.
// tc: optional tensor-core path of the hidden layers (null or !enabled: fp32 SIMT everywhere).
cudaError_t lang_head_run(const LangDims& d, const LangParams& p, const LangWorkspace& ws, const float* E, float* dE,
                          const int* perms, const float* lang_emb, const float* lang_mask, float langw,
                          float* metrics, int* launches, cudaStream_t s, const LangTc* tc = nullptr);

}  // namespace r3m
