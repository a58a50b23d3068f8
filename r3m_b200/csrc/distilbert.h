// Frozen DistilBERT sentence encoder of the language branch (reference: r3m/models/models_language.py:13-35 —
// AutoModel("distilbert-base-uncased")(ids, attention_mask).last_hidden_state.mean(1)), SURVEY.md §8 f2.
// The arithmetic of transformers' DistilBertModel (embeddings + LayerNorm, 6 x [multi-head attention, residual + LayerNorm,
// GELU feed-forward, residual + LayerNorm]) on the sm_100a kernels of this library: every Linear is the tcgen05
// implicit-GEMM kernel in its tf32 tier (a 1x1 "convolution" over the token axis; bias, residual and GELU fused into the
// epilogue) with fp32 operands split into two tf32 terms (three products: fp32-grade results, see distilbert.cu);
// attention / LayerNorm / pooling are small fused fp32 kernels.  Inference only (the reference freezes it).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "convops.h"
#include "engine.h"

namespace r3m {

struct BertDims {
  int vocab = 30522, max_pos = 512, dim = 768, heads = 12, layers = 6, ffn = 3072;
};

class DistilBert {
 public:
  static std::string create(const BertDims& d, DistilBert** out);
  ~DistilBert();
  // HF state_dict names ("embeddings.word_embeddings.weight", "transformer.layer.0.attention.q_lin.weight", ...) ->
  // element offsets inside the flat fp32 parameter buffer; Linear weights keep their [out][in] layout
  const std::vector<TensorInfo>& tensors() const { return tensors_; }
  size_t num_params() const { return nparams_; }
  size_t workspace_bytes(int max_tokens) const;
  // params: flat fp32 parameter buffer (device, num_params() floats, caller-filled); ws: workspace_bytes(max_tokens)
  std::string bind(float* params, void* ws, size_t ws_bytes, int max_tokens);
  std::string sync_weights(cudaStream_t stream);  // tf32-rounded operand copies; call after (re)loading parameters
  // ids int32 [B][T], mask fp32 [B][T] (1 = token, 0 = padding), out fp32 [B][dim] = mean over ALL T positions of the
  // last hidden state (padding included, like the reference's `.mean(1)`).  hidden (optional): fp32 [B][T][dim], the
  // last hidden state itself.
  std::string forward(const int* ids, const float* mask, int B, int T, float* out, float* hidden, cudaStream_t stream);
  int dim() const { return d_.dim; }
  int launches_last_call() const { return launches_; }

 private:
  struct Layer {
    size_t w[6], b[6];  // q, k, v, out, ff1, ff2 (offsets in the parameter buffer)
    size_t wt[6];       // offsets of the split operand copies [Cout][hi | hi | lo]
    size_t ln1_w, ln1_b, ln2_w, ln2_b;
  };
  struct Plans {
    std::vector<ConvPlan> gemm;  // per layer: q, k, v, out, ff1 (column slices), ff2
    int per_layer = 0;
  };
  std::string plan_for(int M, Plans** out);
  std::string enqueue(const int* ids, const float* mask, int B, int T, float* out, float* hidden, cudaStream_t stream);
  struct Graph {
    cudaGraphExec_t exec = nullptr;
    int seen = 0, launches = 0;
    bool failed = false;
    uint64_t last_use = 0;
  };
  uint64_t graph_clock_ = 0;
  std::map<std::vector<uint64_t>, Graph> graphs_;  // (buffers, shape) -> captured forward
  cudaStream_t cap_ = nullptr;

  BertDims d_;
  std::vector<TensorInfo> tensors_;
  std::vector<Layer> layers_;
  size_t word_off_ = 0, pos_off_ = 0, eln_w_ = 0, eln_b_ = 0, nparams_ = 0;
  size_t lin_begin_ = 0, lin_end_ = 0;  // the transformer layers' parameters
  size_t lin_weight_floats_ = 0;        // Linear weights of all layers (their split copies take three times this)
  float* P_ = nullptr;   // fp32 master parameters
  float* Pt_ = nullptr;  // split operand copies of the Linear weights: [Cout][hi | hi | lo] (distilbert.cu: split_tf32)
  int max_tokens_ = 0;
  // activations (fp32): x residual stream, xr / ctx / ffs the split [hi | lo | hi] GEMM operands (LayerNorm output,
  // attention output, GELU output), qkv [M][3*dim], h pre-LayerNorm sums, ff GELU(lin1) before the split
  float *x_ = nullptr, *xr_ = nullptr, *qkv_ = nullptr, *ctx_ = nullptr, *h_ = nullptr, *ff_ = nullptr, *ffs_ = nullptr,
        *ones_ = nullptr;
  std::map<int, Plans> plans_;
  int launches_ = 0;
};

}  // namespace r3m
