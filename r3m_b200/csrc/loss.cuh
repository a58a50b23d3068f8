// Loss heads of Trainer.update (reference r3m/trainer.py:51-59 LP norms, :63-118 language InfoNCE, :120-150 TCN InfoNCE).
#pragma once
#include <cuda_runtime.h>

namespace r3m {

// slots of the device metrics buffer (fp32[16]); same key set as the reference's metrics dict
enum Metric {
  kL2 = 0, kL1 = 1, kL0 = 2, kRewLoss = 3, kRewAcc1 = 4, kRewAcc2 = 5, kRewAcc3 = 6, kTcnLoss = 7, kAligned = 8,
  kFullLoss = 9, kDeviceFlag = 15 /* != 0: a tcgen05 pipeline watchdog fired during the step */, kNumMetrics = 16
};
constexpr float kLossEps = 1e-8f;  // r3m/trainer.py:18

// One block per embedding row: L2 / L1 / L0 row norms -> metrics (means) and, when dE != null,
// dE[row] = l2w * e/(|e|_2 * rows) + l1w * sign(e)/rows   (this INITIALISES dE; the other heads accumulate on top).
// All reductions are ordered (per-row / per-clip partials in `scratch`, summed by one block): bit-identical run to run.
// scratch: rows * 4 floats (loss_lp) / B * 40 floats (loss_tcn); null: a process-wide buffer (stream-ordered use).
cudaError_t launch_loss_lp(const float* E, float* dE, int rows, int D, float l2w, float l1w, float* metrics,
                           cudaStream_t s, float* scratch = nullptr);

// One block per clip: 9 "sim" values (3 in-clip, 3+3 against permuted clips), the two InfoNCE terms, the `aligned`
// metric and, when dE != null, accumulation of d(tcnw * tcnloss)/dE by a gather pass (one block per embedding row).  l2dist != 0: sim = negative L2 distance
// (the default, models_r3m.py:102-104); l2dist == 0: nn.CosineSimilarity(dim=1) (:105-107).  perms: int32 [15][B] in
// the reference's draw order (rows 9..14 are the TCN permutations: es0 then es2 per iteration).
cudaError_t launch_loss_tcn(const float* E, float* dE, const int* perms, int B, int D, float tcnw, int l2dist,
                            float* metrics, cudaStream_t s, float* scratch = nullptr);

// metrics[kDeviceFlag] = (float)*flag  — lets the single metrics read-back also carry the kernels' error flag
cudaError_t launch_publish_flag(const int* flag, float* metrics, cudaStream_t s);
// out[0 .. n) <- NaN when the device-side pipeline watchdog flag is set (forward-only calls: no metrics read-back)
cudaError_t launch_poison_on_flag(const int* flag, float* out, size_t n, cudaStream_t s);

}  // namespace r3m
