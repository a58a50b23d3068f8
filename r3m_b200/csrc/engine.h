// The R3M pretraining engine: a static launch schedule (forward, loss heads, backward, Adam) over one arena of
// device memory, for one backbone size and one frame count.  Mirrors what R3M.forward (r3m/models/models_r3m.py:84)
// and Trainer.update (r3m/trainer.py:25) make ATen do, re-planned for sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "convops.h"
#include "lang.cuh"

namespace r3m {

enum OpFamily { kFamConv = 0, kFamWgrad = 1, kFamNorm = 2, kFamPool = 3, kFamLoss = 4, kFamOptim = 5, kFamLang = 6,
                kNumFamilies = 8 };

enum TensorKind {
  kConvKRSC = 0,   // conv filter, stored [Cout][R][S][Cin] fp32 (state_dict: OIHW)
  kStemOIHW = 1,   // the 7x7 stem filter, stored OIHW fp32
  kVector = 2,     // BN gamma / beta
  kRunMean = 3,    // BN running_mean  (buffers region)
  kRunVar = 4,     // BN running_var   (buffers region)
  kLinearW = 5,    // language head weight [out][in]
  kLinearB = 6,    // language head bias
};

struct TensorInfo {
  std::string name;  // state_dict key without the "module." prefix
  int kind = 0;
  size_t offset = 0;  // element offset inside the flat parameter buffer (kinds 0,1,2,5,6) or the buffers region (3,4)
  int dims[4] = {1, 1, 1, 1};  // logical (state_dict) shape
  int ndim = 1;
};

struct Hyper {
  float l2weight = 1e-5f, l1weight = 1e-5f, langweight = 0.f, tcnweight = 1.f;
};

class Engine {
 public:
  // size in {18, 34, 50}; frames = images per forward (5 * clips for update()).
  static std::string create(int size, int frames, int lang_head, int hidden_dim, Engine** out);
  ~Engine();

  size_t workspace_bytes() const { return ws_bytes_; }
  size_t param_block_bytes() const { return pws_bytes_; }
  // params: the model's parameter block (shared between engines of different frame counts; caller-initialised);
  // ws: this engine's activation workspace.
  std::string bind(void* params, size_t param_bytes, void* ws, size_t bytes, cudaStream_t stream);
  const std::vector<TensorInfo>& tensors() const { return tensors_; }
  size_t num_params() const { return nparams_; }
  size_t num_buffer_floats() const { return nbuf_; }
  // 0: params fp32, 1: grads fp32, 2: adam m, 3: adam v, 4: BN buffers fp32, 5: embeddings fp32 [frames][D],
  // 6: d(loss)/d(embeddings) fp32, 7: metrics fp32[16]
  void* region(int which) const;
  int embed_dim() const { return D_; }
  void set_l2dist(bool v) { l2dist_ = v; }
  int frames() const { return N_; }

  std::string sync_weights(cudaStream_t stream);  // params fp32 -> bf16 operands (+ dgrad / stem re-packs)
  // format of the frames passed to forward / update_grads (ObsFormat, elementwise.cuh): fp32 NCHW (default), uint8 NCHW,
  // uint8 NHWC.  Sticky until changed.
  std::string set_obs_format(int format);
  // 0: bf16 storage (default; training and fast inference), 1: the tf32 parity tier — EVAL-mode forward only: fp32
  // storage rounded to tf32, kind::tf32 tensor-core arithmetic, embeddings within 1e-3 of the fp32 reference
  std::string set_precision(int precision) {
    if (precision != 0 && precision != 1) return "precision must be 0 (bf16) or 1 (tf32)";
    precision_ = precision;
    return std::string();
  }
  std::string forward(const void* obs, int train, float* out, cudaStream_t stream);
  std::string update_grads(const void* obs, const int* perms, const float* lang_emb, const float* lang_mask,
                           const Hyper& h, int eval, cudaStream_t stream);
  // Backward pass alone, for callers that compute the loss themselves (the reference's own Trainer through
  // torch.autograd): dE = d(loss)/d(embeddings) fp32 [frames][D] of the preceding train-mode forward().  Filter
  // gradients are ACCUMULATED into the gradient region, BatchNorm gradients written.
  std::string backward(const float* dE, cudaStream_t stream);
  // Gradient chunks for an all-reduce that overlaps the backward pass: chunk k covers the flat gradient elements
  // [begin, end) — chunk 0 the language head and layer 4, then layer 3, layer 2, and last the stem with layer 1 — in the
  // order in which the backward pass completes them.  wait_grad_chunk makes `stream` wait (cudaStreamWaitEvent) until
  // the LAST update_grads / backward enqueued on this engine has produced chunk k; call it after that call returned.
  int num_grad_chunks() const;
  std::string grad_chunk(int k, size_t* begin, size_t* end) const;
  std::string wait_grad_chunk(int k, cudaStream_t stream);
  // Test hooks (tests/test_block_backward_gpu.py): the buffers of residual block `block` (what: 0 input activation,
  // 1 output activation, 2 incoming gradient, 3 outgoing gradient; all bf16 NHWC) and a run of ONLY that block's
  // backward ops on one stream, after a train-mode forward.
  int num_blocks() const;
  std::string debug_block(int block, int what, void** ptr, size_t* count) const;
  std::string debug_run_block_backward(int block, cudaStream_t stream);
  // step: 1-based Adam step count (bias correction); the caller owns it because engines share a parameter block
  std::string adam_step(float lr, float grad_scale, int step, cudaStream_t stream);
  int launches_last_call() const { return launches_; }
  int graph_replays() const { return graph_replays_; }  // cudaGraphLaunch calls of the training step so far
  // per-launch records of the last profile_update: {family, ms, flops, bytes} each
  const std::vector<double>& last_profile_ops() const { return prof_last_; }
  const std::vector<std::string>& last_profile_labels() const { return prof_labels_; }
  // Runs ONE update_grads + adam_step with a CUDA-event pair around every launch (serialises nothing: events are
  // recorded in-stream) and accumulates per-family device time.  out: [kNumFamilies][4] = {ms, flops, bytes, launches}.
  std::string profile_update(const void* obs, const int* perms, const float* lang_emb, const float* lang_mask,
                             const Hyper& h, float lr, int step, double* out, cudaStream_t stream);
  void param_block_layout(size_t* offsets5) const {
    offsets5[0] = off_P_; offsets5[1] = off_G_; offsets5[2] = off_M_; offsets5[3] = off_V_; offsets5[4] = off_buf_;
  }

 private:
  Engine() {}
  struct Conv;
  struct Block;
  struct Op {  // one kernel launch of the static schedule + what it costs algorithmically
    std::function<cudaError_t(cudaStream_t)> fn;
    int family = 0;      // OpFamily
    double flops = 0.0;  // algorithmic FLOPs (2*MAC) of the launch
    double bytes = 0.0;  // algorithmic HBM bytes (operands read once + results written once)
    std::string label;   // what the launch is (layer / role), for the per-launch profile
    int nlaunch = 1;     // kernels the op launches (fused two-pass ops count both)
    int block = -1;      // backward ops: index of the residual block they belong to (-1: head / stem)
    // cross-stream schedule of the backward pass: ops with side = true run on the engine's second stream
    bool side = false;
    std::vector<int> wait;  // event ids the op's stream waits on before the launch
    int record = -1;        // event id recorded on the op's stream after the launch
    Op() {}
    template <class F>
    Op(F f, int fam = 0, double fl = 0.0, double by = 0.0) : fn(f), family(fam), flops(fl), bytes(by) {}
  };

  std::string plan_all();
  std::string run(const std::vector<Op>& ops, cudaStream_t stream);
  void add_bn_apply(std::vector<Op>& ops, const Conv& c, const void* residual, void* dst, int relu, int train,
                    const Conv* second);

  int size_ = 0, N_ = 0, D_ = 0, B_ = 0, lang_ = 0, hidden_ = 0;
  bool bottleneck_ = false;
  bool l2dist_ = true;  // R3M.sim: negative L2 distance (default) or cosine similarity
  int obs_format_ = 0;  // ObsFormat of the frames handed to forward / update_grads
  int precision_ = 0;   // inference tier: 0 bf16, 1 tf32
  // tf32 tier: byte offsets inside the aliased activation region (relative to off_xs_)
  size_t t32_xs_ = 0, t32_buf_[5] = {0, 0, 0, 0, 0}, t32_params_ = 0, t32_stem_w_ = 0, t32_E_ = 0;
  std::vector<Conv*> convs_;
  std::vector<Block*> blocks_;
  std::vector<TensorInfo> tensors_;
  size_t nparams_ = 0, nbuf_ = 0;
  size_t ws_bytes_ = 0, pws_bytes_ = 0;
  uint8_t* ws_ = nullptr;
  uint8_t* pws_ = nullptr;
  bool bound_ = false;
  int launches_ = 0;
  bool fwd_train_valid_ = false;  // a train-mode forward() has run since the last update_grads (see obs == NULL there)
  int fwd_launches_ = 0;
  bool profiling_ = false;
  std::vector<cudaEvent_t> prof_events_;
  std::vector<int> prof_ops_family_;
  std::vector<double> prof_flops_, prof_bytes_, prof_last_;
  std::vector<std::string> prof_labels_, prof_labels_run_;
  cudaError_t launch(const Op& op, cudaStream_t stream);

  // Whole-step CUDA graphs.  The training step is two fixed launch sequences — the train-mode forward (memset,
  // preprocess, ~105 launches) and everything behind it (loss heads, language head, the two-stream backward pass: ~250
  // launches) — whose only variable inputs are pointers and a few scalars.  The first call with a given set of them runs
  // plainly (it also configures every kernel's attributes), the second is captured (relaxed stream capture, the side
  // stream joins through the schedule's own events) and instantiated, every later one is ONE cudaGraphLaunch.
  // R3M_STEP_GRAPH=0 disables it.  The gradient-chunk markers are additionally recorded as EXTERNAL events so that
  // wait_grad_chunk works on a replayed graph.
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int seen = 0;
    int launches = 0;
    bool failed = false;
    uint64_t last_use = 0;
  };
  uint64_t graph_clock_ = 0;
  std::string run_cached(const std::vector<uint64_t>& key, cudaStream_t stream,
                         const std::function<std::string(cudaStream_t)>& body, bool* graphed);
  std::map<std::vector<uint64_t>, GraphEntry> graphs_;
  bool use_graph_ = true;
  bool capturing_ = false;
  bool last_bwd_graphed_ = false;
  int graph_replays_ = 0;
  cudaStream_t cap_ = nullptr;
  std::vector<cudaEvent_t> ext_evs_;   // external twins of the gradient-chunk events (null elsewhere)

  // arena offsets (bytes)
  size_t off_P_ = 0, off_G_ = 0, off_M_ = 0, off_V_ = 0, off_Pb_ = 0, off_buf_ = 0, off_saved_ = 0, off_zero_ = 0,
         zero_bytes_ = 0, off_metrics_ = 0, off_stem_dwp_ = 0, off_wd_ = 0, off_stem_wp_ = 0, off_xs_ = 0, off_argmax_ = 0, off_ymax_ = 0,
         off_E_ = 0, off_dE_ = 0, off_g_[7] = {0, 0, 0, 0, 0, 0, 0};
  size_t nsaved_ = 0, nwd_ = 0;
  // deterministic-reduction scratch (see Engine::create)
  size_t off_det_ = 0, off_det_bn_ = 0, off_det_loss_ = 0, det_small_bytes_ = 0, off_wgrad_scratch_[2] = {0, 0};
  // language head (optional)
  size_t lang_w_off_[5] = {0, 0, 0, 0, 0}, lang_b_off_[5] = {0, 0, 0, 0, 0};
  size_t off_lang_ws_ = 0, off_lang_tc_ = 0;
  LangTc lang_tc_;  // tensor-core path of the language head's hidden layers (plans + split operand buffers)
  size_t off_fold_ = 0;  // BnFoldEntry table (device) for the inference path
  size_t off_pack_ = 0;  // PackDgradEntry table (device): all dgrad filter re-packs in one launch
  // small-batch inference: the eval forward is replayed as a CUDA graph from a fixed staging copy of the frames
  size_t off_obs_stage_ = 0;
  cudaGraphExec_t eval_graph_ = nullptr;
  int eval_calls_ = 0;
  LangDims lang_dims_;

  std::vector<Op> fwd_train_, fwd_eval_, fwd_eval_tf32_, bwd_, repack_;
  struct GradChunk {
    size_t begin = 0;     // first flat gradient element of the chunk (it ends where the previous chunk begins)
    int main_event = -1;  // recorded on the main stream / the side stream when the chunk's last producers are enqueued
    int side_event = -1;
  };
  std::vector<GradChunk> chunks_;
  // Filter gradients run on a second stream: nothing in the backward chain consumes them, and a wgrad CTA (tensor /
  // L2 bound, one per SM) co-resides with the HBM-bound BatchNorm-backward CTAs of the layer below.
  cudaStream_t side_ = nullptr;
  std::vector<cudaEvent_t> evs_;
  bool use_side_ = true;
  bool fuse_bn_bwd_ = false;  // small layers: BatchNorm backward as one launch with a grid barrier (R3M_FUSE_BN_BWD=1)
  double l2_order_min_bytes_ = 0.0;  // tensors below this size keep the first-to-last walk (R3M_L2_ORDER_MIN_MB; default: none)
  bool l2_order_ = true;  // alternate the traversal direction of consecutive passes over a tensor (R3M_L2_ORDER=0: off)
};

}  // namespace r3m
