// Video-language alignment head of Trainer.update (reference r3m/trainer.py:63-118, LanguageReward MLP
// r3m/models/models_language.py:37-55).  All 15 get_reward() evaluations of one update are batched into ONE
// [15*B, 2D+768] pass (the reference runs 15 separate MLP calls, 75 small cuBLAS launches).  fp32 throughout: the
// reference's Linear layers run in true fp32 (cuda.matmul.allow_tf32 == False) and the loss tolerance is 1e-4.
#pragma once
#include <cuda_runtime.h>

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
size_t lang_workspace_floats(const LangDims& d);
void lang_carve_workspace(float* base, const LangDims& d, LangWorkspace* ws);

// Forward + InfoNCE loss (+ metrics).  When dE != null also the full backward: parameter gradients are WRITTEN to
// p.dw / p.db and d(langw * rewloss)/dE is accumulated into dE (ordered reductions only: no atomics).  Returns the number of kernels launched
// through *launches.
cudaError_t lang_head_run(const LangDims& d, const LangParams& p, const LangWorkspace& ws, const float* E, float* dE,
                          const int* perms, const float* lang_emb, const float* lang_mask, float langw,
                          float* metrics, int* launches, cudaStream_t s);

}  // namespace r3m
