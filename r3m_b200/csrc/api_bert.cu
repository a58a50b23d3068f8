// extern "C" boundary of the DistilBERT sentence encoder (include/r3m_b200.h, "Sentence encoder" section).
#include "../../include/r3m_b200.h"

#include <cstdio>
#include <string>

#include "api_util.h"
#include "distilbert.h"

using namespace r3m;

#define BERT_OR_FAIL(h)                                                        \
  if (!(h)) return fail(R3M_B200_ERR_INVALID, "null sentence-encoder handle"); \
  DistilBert* enc = static_cast<DistilBert*>(h)

extern "C" {

int r3m_b200_distilbert_create(int vocab, int max_pos, int dim, int heads, int layers, int ffn, void** handle) {
  if (!handle) return fail(R3M_B200_ERR_INVALID, "null handle pointer");
  BertDims d;
  d.vocab = vocab;
  d.max_pos = max_pos;
  d.dim = dim;
  d.heads = heads;
  d.layers = layers;
  d.ffn = ffn;
  DistilBert* m = nullptr;
  const std::string err = DistilBert::create(d, &m);
  if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
  *handle = m;
  return R3M_B200_OK;
}

int r3m_b200_distilbert_destroy(void* handle) {
  BERT_OR_FAIL(handle);
  delete enc;
  return R3M_B200_OK;
}

int r3m_b200_distilbert_num_params(void* handle, size_t* count) {
  BERT_OR_FAIL(handle);
  *count = enc->num_params();
  return R3M_B200_OK;
}

int r3m_b200_distilbert_num_tensors(void* handle, int* count) {
  BERT_OR_FAIL(handle);
  *count = (int)enc->tensors().size();
  return R3M_B200_OK;
}

int r3m_b200_distilbert_tensor_info(void* handle, int index, char* name, int name_capacity, long long* offset, int* ndim,
                                    int* dims2) {
  BERT_OR_FAIL(handle);
  if (index < 0 || index >= (int)enc->tensors().size()) return fail(R3M_B200_ERR_INVALID, "tensor index out of range");
  const TensorInfo& t = enc->tensors()[index];
  if ((int)t.name.size() + 1 > name_capacity) return fail(R3M_B200_ERR_INVALID, "name buffer too small");
  std::snprintf(name, name_capacity, "%s", t.name.c_str());
  *offset = (long long)t.offset;
  *ndim = t.ndim;
  dims2[0] = t.dims[0];
  dims2[1] = t.dims[1];
  return R3M_B200_OK;
}

int r3m_b200_distilbert_workspace_bytes(void* handle, int max_tokens, size_t* bytes) {
  BERT_OR_FAIL(handle);
  if (max_tokens < 1) return fail(R3M_B200_ERR_INVALID, "max_tokens must be positive");
  *bytes = enc->workspace_bytes(max_tokens);
  return R3M_B200_OK;
}

int r3m_b200_distilbert_bind(void* handle, float* params, void* workspace, size_t bytes, int max_tokens) {
  BERT_OR_FAIL(handle);
  const std::string err = enc->bind(params, workspace, bytes, max_tokens);
  if (!err.empty()) return fail(R3M_B200_ERR_INVALID, err);
  return R3M_B200_OK;
}

int r3m_b200_distilbert_sync_weights(void* handle, void* stream) {
  BERT_OR_FAIL(handle);
  const std::string err = enc->sync_weights(static_cast<cudaStream_t>(stream));
  if (!err.empty()) return fail(R3M_B200_ERR_CUDA, err);
  return R3M_B200_OK;
}

int r3m_b200_distilbert_forward(void* handle, const int* ids, const float* mask, int B, int T, float* out, float* hidden,
                                void* stream) {
  BERT_OR_FAIL(handle);
  const std::string err = enc->forward(ids, mask, B, T, out, hidden, static_cast<cudaStream_t>(stream));
  if (!err.empty()) return fail(R3M_B200_ERR_CUDA, err);
  return R3M_B200_OK;
}

int r3m_b200_distilbert_launches(void* handle, int* count) {
  BERT_OR_FAIL(handle);
  *count = enc->launches_last_call();
  return R3M_B200_OK;
}

}  // extern "C"
