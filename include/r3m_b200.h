/* r3m_b200 — C ABI of the B200-native R3M pretraining hot path.
 *
 * The reference (facebookresearch/r3m) is pure Python; the arithmetic of its hot path lives in the ATen ops that
 * torchvision.models.resnet{18,34,50} and r3m/trainer.py dispatch.  Each entry point below replaces the ATen call(s)
 * named in its comment (reference file:line).  All pointers are DEVICE pointers unless stated otherwise; activations
 * are NHWC bf16, statistics / parameters / gradients fp32.  `stream` is a cudaStream_t passed as void*.
 *
 * Every function returns 0 on success and a negative code on failure; r3m_b200_last_error() returns a
 * thread-local message for the last failure.  Nothing here falls back to a CPU or library path: on a machine
 * without an sm_100 GPU the compute entry points fail.
 */
#ifndef R3M_B200_H_
#define R3M_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R3M_B200_OK 0
#define R3M_B200_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define R3M_B200_ERR_CUDA (-2)      /* CUDA runtime or driver error */
#define R3M_B200_ERR_KERNEL (-3)    /* a device-side pipeline watchdog fired */
#define R3M_B200_ERR_STATE (-4)     /* call made in the wrong engine state */

const char* r3m_b200_last_error(void);
int r3m_b200_abi_version(void);
/* Reads and clears the device-side watchdog flag (synchronises the device). 0 = clean. */
int r3m_b200_check_device_flag(void);

/* ------------------------------------------------------------------------------------------------------------
 * Kernel-level entry points (unit parity tests and micro-benchmarks call these; the engine uses the same kernels)
 * ------------------------------------------------------------------------------------------------------------ */

/* Forward convolution, bias-free (replaces aten::cudnn_convolution for tv resnet.py:42-56 conv3x3/conv1x1).
 *   x  bf16 [N,H,W,Cin]      w  bf16 [Cout,R,S,Cin]      y  bf16 [N,P,Q,Cout]   (Cin, Cout multiples of 64)
 *   stat_sum/stat_sq: optional fp32 [Cout], accumulated (+=) per-channel sum / sum of squares of y (train-mode BN). */
int r3m_b200_conv_fwd(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int R, int S,
                      int stride, int pad, float* stat_sum, float* stat_sq, void* stream);

/* Inference form of the forward convolution: BatchNorm folded into the epilogue (replaces cudnn_convolution +
 * cudnn_batch_norm(eval) + add_ + relu_ of one residual-block stage, tv resnet.py:89-105,143-163).
 *   y = [relu](conv(x, w) * scale[c] + shift[c] [+ residual]);  scale/shift fp32 [Cout], residual bf16 like y or NULL. */
int r3m_b200_conv_fwd_affine(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int R, int S,
                             int stride, int pad, const float* scale, const float* shift, const void* residual,
                             int relu, void* stream);

/* Re-pack a master filter fp32 [Cout,R,S,Cin] into the dgrad operand (bf16, Cout*R*S*Cin elements: one
 * [Cin][taps][Cout] block per output-parity class, classes in (ph,pw) row-major order). */
int r3m_b200_pack_dgrad_filter(const float* w, void* w_dgrad, int Cout, int R, int S, int Cin, int stride, int pad,
                               void* stream);

/* Data gradient (replaces aten::cudnn_convolution_backward_input).  dy bf16 [N,P,Q,Cout] -> dx bf16 [N,H,W,Cin].
 *   accumulate != 0: dx += result. */
int r3m_b200_conv_dgrad(const void* dy, const void* w_dgrad, void* dx, int N, int H, int W, int Cin, int Cout, int R,
                        int S, int stride, int pad, int accumulate, void* stream);

/* Filter gradient (replaces aten::cudnn_convolution_backward_weight).  dw fp32 [Cout,R,S,Cin] is ACCUMULATED into. */
int r3m_b200_conv_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                        int stride, int pad, void* stream);

/* Input normalisation + stem re-layout (replaces aten::div/sub/div of models_r3m.py:97-98 fused with the layout the
 * stem conv consumes).  obs fp32 NCHW [N,3,224,224] in [0,255] -> xs bf16 [N,112,112,64]; channel
 * j = kw*16 + (dy*2+dx)*4 + c holds normalise(obs[n,c,2i+dy,2(q-2+kw)+dx]) (zero outside the image / for c == 3). */
int r3m_b200_preprocess_stem(const float* obs, void* xs, int N, void* stream);
/* The same for frames that arrive as uint8 (format 1: uint8 NCHW [N,3,224,224] — what torchvision.io.read_image
 * yields, r3m/utils/data_loaders.py:30-32; format 2: uint8 NHWC [N,224,224,3]; format 0: fp32 NCHW): the
 * `obs.float()` of models_r3m.py:97 happens in registers, a quarter of the bytes cross PCIe and HBM. */
int r3m_b200_preprocess_stem_format(const void* obs, int format, void* xs, int N, void* stream);

/* torchvision.transforms.RandomResizedCrop(224, ...) arithmetic of the reference loader's augmentation
 * (r3m/utils/data_loaders.py:47-50,81-102) on the GPU: for frame n, crop box boxes[n] = (top, left, h, w) of the uint8
 * source frame ([N,3,H,W], or [N,H,W,3] when nhwc != 0) resized to 224x224 with antialiased bilinear interpolation
 * (aten::_upsample_bilinear2d_aa; borders clamp to the crop).  out: fp32 NCHW [N,3,224,224] in [0,255], the loader's
 * output contract.  boxes: DEVICE int32 [N][4], drawn by the host with the reference's sampling law. */
int r3m_b200_random_resized_crop(const uint8_t* src, int nhwc, int N, int H, int W, const int* boxes, float* out,
                                 void* stream);

/* BatchNorm2d forward on a raw conv output (replaces aten::cudnn_batch_norm + relu_ [+ add_], tv resnet.py:89-105,
 * 143-163).  y, a, residual: bf16 [M][C].  train != 0: batch statistics from (sum, sq) = per-channel sum / sum of
 * squares of y; writes save_mean / save_rstd and updates running_mean / running_var (momentum 0.1, unbiased var).
 * train == 0: running statistics.  a = [relu](gamma * xhat + beta [+ residual] [+ bn2(y2)]).
 *   mask_out (optional, uint8 [M][C/8]): bit-packed (a > 0), the ReLU mask the backward pass consumes.
 *   y2 ... save_rstd2 (optional, all NULL otherwise): a SECOND BatchNorm — the downsample branch of a residual
 *   block — whose un-activated output is added before the ReLU without being materialised. */
int r3m_b200_bn_apply(const void* y, void* a, const void* residual, int M, int C, int relu, int train, const float* sum,
                      const float* sq, const float* gamma, const float* beta, float* running_mean, float* running_var,
                      float* save_mean, float* save_rstd, uint8_t* mask_out, const void* y2, const float* sum2,
                      const float* sq2, const float* gamma2, const float* beta2, float* running_mean2,
                      float* running_var2, float* save_mean2, float* save_rstd2, void* stream);

/* BatchNorm2d backward (replaces aten::cudnn_batch_norm_backward + threshold_backward).  dA: gradient w.r.t. the
 * activated output; ReLU mask either from a (activated output, bf16) or from mask (bit-packed, as written by
 * bn_apply); both NULL: no ReLU.  y: raw conv output; sums: fp32 [2*C] scratch that must be zero on entry.  Writes
 * dy (gradient w.r.t. y), optionally dz (masked gradient, the residual branch's share), dgamma, dbeta.
 *   y2 ... dbeta2 (optional): the second (downsample) BatchNorm fed by the same masked gradient; sums2 fp32 [C]
 *   zeroed scratch, dy2 its data gradient. */
int r3m_b200_bn_backward(const void* dA, const void* a, const uint8_t* mask, const void* y, int M, int C,
                         const float* mean, const float* rstd, const float* gamma, float* sums, void* dy, void* dz,
                         float* dgamma, float* dbeta, const void* y2, const float* mean2, const float* rstd2,
                         const float* gamma2, float* sums2, void* dy2, float* dgamma2, float* dbeta2, void* stream);

/* Stem tail: BN + ReLU + MaxPool2d(3,2,1) (tv resnet.py:198-200) and its backward.  y bf16 [N,H,W,C] ->
 * a bf16 [N,H/2,W/2,C] plus the argmax code per output element (0..8: scan-order window position of the maximum;
 * 15: the maximum is 0, i.e. clipped by the ReLU, and carries no gradient).  maxpool_backward returns the gradient
 * w.r.t. the BN output with the ReLU mask applied (`a` is unused and may be null).  stem_backward is the fused form
 * the engine runs: aten::max_pool2d_with_indices_backward + threshold_backward + cudnn_batch_norm_backward in two
 * passes that recompute the pooled gradient scatter on the fly (sums: fp32 [2*C] zeroed scratch).  ymax (optional in both
 * calls): bf16 [N,H/2,W/2,C], the RAW conv output at each window's argmax; when given, the backward's reduce pass runs
 * over the pooled elements only. */
int r3m_b200_stem_bn_relu_maxpool(const void* y, void* a, uint8_t* argmax, void* ymax, int N, int H, int W, int C, int train,
                                  const float* sum, const float* sq, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float* save_mean, float* save_rstd,
                                  void* stream);
int r3m_b200_maxpool_backward(const void* dA, const void* a, const uint8_t* argmax, void* dz, int N, int H, int W, int C,
                              void* stream);
int r3m_b200_stem_backward(const void* dA, const uint8_t* argmax, const void* ymax, const void* y, int N, int H, int W, int C,
                           const float* mean, const float* rstd, const float* gamma, float* sums, void* dy,
                           float* dgamma, float* dbeta, void* stream);

/* Test hook of the deterministic reductions: out[0] = the exact sum of x[0..n) rounded once to fp32, accumulated by
 * `blocks` thread blocks through the 128-bit fixed-point accumulators that replace fp32 atomics in the step's
 * BatchNorm statistics / BatchNorm-backward sums (aten's cudnn_batch_norm reductions are order dependent; these are
 * not).  The result is independent of `blocks`. */
int r3m_b200_ordered_sum(const float* x, size_t n, float* out, int blocks, void* stream);
/* Test hook of the BatchNorm batch statistics as the engine computes them (aten::cudnn_batch_norm's mean / biased
 * variance, tv resnet.py:89-105): out[0] = mean, out[1] = variance of x[0..n) from the fixed-point sum and sum of
 * squares, with the E[x^2] - mean^2 subtraction carried out in fp64 so that channels with |mean| >> std do not lose
 * their variance to cancellation. */
int r3m_b200_ordered_moments(const float* x, int n, float* out, int blocks, void* stream);

/* Side-band upload of the step's small host inputs (replaces the `.cuda()` calls of r3m/trainer.py:108 and the index
 * tensors of :86-92,135-137): dst (device) <- host_pinned (page-locked host memory, device-accessible under unified
 * addressing), copied by a kernel on `stream` instead of the H2D copy engine, so it never queues behind a bulk frame
 * upload.  bytes must be a multiple of 16 and both pointers 16-byte aligned. */
int r3m_b200_pull_host(const void* host_pinned, void* dst, size_t bytes, void* stream);

/* AdaptiveAvgPool2d((1,1)) + flatten (tv resnet.py:278-279) and its backward.  a bf16 [N,HW,C] <-> fp32 [N,C]. */
int r3m_b200_avgpool_forward(const void* a, float* out, int N, int HW, int C, void* stream);
int r3m_b200_avgpool_backward(const float* dE, void* dA, int N, int HW, int C, void* stream);

/* Loss heads on embeddings E fp32 [5*B][D] (trainer.py:51-59 and :120-150).  metrics: fp32[16] accumulated into
 * (zero it first); dE (may be NULL): loss_lp WRITES d/dE of the weighted penalty, loss_tcn ACCUMULATES on top. */
int r3m_b200_loss_lp(const float* E, float* dE, int rows, int D, float l2weight, float l1weight, float* metrics,
                     void* stream);
int r3m_b200_loss_tcn(const float* E, float* dE, const int* perms, int B, int D, float tcnweight, float* metrics,
                      void* stream);
/* The same head with the similarity of R3M.sim selectable (models_r3m.py:102-107): l2dist != 0 negative L2 distance
 * (what r3m_b200_loss_tcn computes), l2dist == 0 nn.CosineSimilarity(dim=1). */
int r3m_b200_loss_tcn_sim(const float* E, float* dE, const int* perms, int B, int D, float tcnweight, int l2dist,
                          float* metrics, void* stream);

/* torch.optim.Adam step over a flat fp32 buffer (defaults beta 0.9/0.999, eps 1e-8), also emitting the bf16 copy. */
int r3m_b200_adam(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n, float lr, int step,
                  float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Engine: the whole hot path behind R3M.forward (r3m/models/models_r3m.py:84-100) and Trainer.update
 * (r3m/trainer.py:25-162) for one backbone size and one frame count.  The caller owns two device allocations
 * (1024-byte aligned, e.g. torch uint8 tensors): the PARAMETER BLOCK (r3m_b200_engine_param_block_bytes(): flat
 * fp32 parameters / gradients / Adam moments, BN running statistics, bf16 operand copies; depends only on the
 * model, so engines for different frame counts of one model share it; the caller zero-fills it and then loads
 * weights) and the per-engine WORKSPACE (r3m_b200_engine_workspace_bytes(): activations, saved statistics).
 * ------------------------------------------------------------------------------------------------------------ */

/* size: 18 | 34 | 50 (torchvision resnet, models_r3m.py:44-52); frames: images per call (5 * clips for update);
 * lang_head: build the LanguageReward MLP (models_language.py:37-55) with hidden_dim units. */
int r3m_b200_engine_create(int size, int frames, int lang_head, int hidden_dim, void** handle);
int r3m_b200_engine_destroy(void* handle);
int r3m_b200_engine_workspace_bytes(void* handle, size_t* bytes);
int r3m_b200_engine_param_block_bytes(void* handle, size_t* bytes);
/* Attaches both allocations and builds the launch schedule (TMA descriptors are encoded here). */
int r3m_b200_engine_bind(void* handle, void* param_block, size_t param_bytes, void* workspace, size_t bytes,
                         void* stream);

/* state_dict bridge: table of named tensors.  kind: 0 conv filter stored [Cout][R][S][Cin] (state_dict OIHW),
 * 1 stem filter stored OIHW, 2 BN weight/bias, 3 running_mean, 4 running_var, 5 linear weight, 6 linear bias.
 * offset: element offset in the parameter region (kinds 0,1,2,5,6) or the BN-buffer region (kinds 3,4). */
int r3m_b200_engine_num_tensors(void* handle, int* count);
int r3m_b200_engine_tensor_info(void* handle, int index, char* name, int name_capacity, int* kind, long long* offset,
                                int* ndim, int* dims4);
/* which: 0 params, 1 grads, 2 Adam m, 3 Adam v (fp32, same flat layout), 4 BN buffers fp32, 5 embeddings fp32
 * [frames][D], 6 d(loss)/d(embeddings) fp32, 7 metrics fp32[16]
 * (slots 0-9: l2loss,l1loss,l0loss,rewloss,rewacc1,rewacc2,rewacc3,tcnloss,aligned,full_loss; slot 15: device-side
 * pipeline-watchdog flag of the step, 0 = clean). count = number of elements. */
int r3m_b200_engine_region(void* handle, int which, void** ptr, size_t* count);
/* what: 0 embedding dim, 1 frames, 2 kernels launched by the last engine call (replayed graphs count their kernels),
 *       3 cudaGraphLaunch calls so far: the training step is replayed as two captured CUDA graphs (train-mode forward;
 *         loss heads + language head + two-stream backward) from the second call with the same input pointers on;
 *         environment R3M_STEP_GRAPH=0 keeps plain launches */
int r3m_b200_engine_get_int(void* handle, int what, int* value);
/* what: 0 similarity of the TCN head: value != 0 negative L2 distance (default; R3M(l2dist=True)), 0 cosine
 *       1 format of the frames `obs` of forward / update_grads / profile_update: 0 fp32 NCHW (default), 1 uint8 NCHW,
 *         2 uint8 NHWC (see r3m_b200_preprocess_stem_format); sticky
 *       2 precision tier of the EVAL-mode forward: 0 bf16 storage (default), 1 tf32 — fp32 storage rounded to tf32,
 *         kind::tf32 tensor cores, fp32 accumulation: embeddings within 1e-3 (relative) of the fp32 reference
 *         (r3m/models/models_r3m.py:97-99 computes in fp32); train-mode forwards and update() always run bf16 */
int r3m_b200_engine_set_int(void* handle, int what, int value);
/* Byte offsets of {params, grads, Adam m, Adam v, BN buffers} inside the parameter block, and the element counts of
 * the flat parameter buffer / the BN-buffer region.  Valid before bind (pure layout query; no GPU needed). */
int r3m_b200_engine_param_block_layout(void* handle, size_t* offsets5, size_t* num_params, size_t* num_buffer_floats);

/* After writing the parameter region: refresh the bf16 operand copies (forward filters, dgrad re-packs, stem). */
int r3m_b200_engine_sync_weights(void* handle, void* stream);

/* R3M.forward: obs [frames,3,224,224] in [0,255] (fp32 NCHW unless set_int(1) says otherwise) -> out fp32 [frames][D] (out may be NULL: result stays in
 * region 5).  train != 0 uses batch statistics and updates the running ones (nn.BatchNorm2d semantics). */
int r3m_b200_engine_forward(void* handle, const void* obs, int train, float* out, void* stream);

/* Trainer.update up to (and excluding) the optimiser step: forward, LP / TCN / language losses, backward.
 *   perms: int32 [15][clips] permutations in the reference's draw order (9 language, then 6 TCN; trainer.py:86-92,
 *   135-137); lang_emb fp32 [clips][768] sentence embeddings and lang_mask fp32 [clips] (both may be NULL when
 *   langweight == 0).  eval != 0: eval-mode BN, no gradients (trainer.py:28-29,155).  Metrics land in region 7.
 *   obs == NULL (training only): the forward pass was already enqueued with r3m_b200_engine_forward(obs, train = 1,
 *   out = NULL) on the same stream, so the host may prepare perms / lang inputs while it runs. */
int r3m_b200_engine_update_grads(void* handle, const void* obs, const int* perms, const float* lang_emb,
                                 const float* lang_mask, float l2weight, float l1weight, float langweight,
                                 float tcnweight, int eval, void* stream);
/* The backward pass alone (what `full_loss.backward()` of r3m/trainer.py:157 runs below the embeddings), for hosts
 * that compute the loss themselves — the reference's own Trainer through torch.autograd:
 *   dE fp32 [frames][D] = d(loss)/d(embeddings) of the preceding r3m_b200_engine_forward(train = 1) on this engine.
 * Filter gradients are ACCUMULATED into region 1 (zero it for a fresh gradient), BatchNorm gradients are written. */
int r3m_b200_engine_backward(void* handle, const float* dE, void* stream);

/* Overlapping the step's one collective with the backward pass (replaces the gradient reduction of nn.DataParallel,
 * r3m/train_representation.py:30).  The flat gradient buffer (region 1) is cut into r3m_b200_engine_num_grad_chunks()
 * chunks [begin, end) (elements) in the order the backward pass completes them: the language head + layer 4 first, the
 * stem + layer 1 last.  After r3m_b200_engine_update_grads / _backward has RETURNED, r3m_b200_engine_wait_grad_chunk
 * makes `stream` wait (cudaStreamWaitEvent) for chunk k of that call, so the host can enqueue one all-reduce per chunk
 * on a communication stream while the rest of the backward pass is still running. */
int r3m_b200_engine_num_grad_chunks(void* handle, int* count);
int r3m_b200_engine_grad_chunk(void* handle, int k, size_t* begin, size_t* end);
int r3m_b200_engine_wait_grad_chunk(void* handle, int k, void* stream);

/* Test hooks for the block-level backward parity test: the bf16 NHWC buffers of residual block `block`
 * (0 .. r3m_b200_engine_num_blocks()-1, forward order) — what: 0 input activation, 1 output activation, 2 incoming
 * gradient (read by the block's backward), 3 outgoing gradient (written) — and a run of ONLY that block's backward
 * kernels (BatchNorm backward, dgrad, wgrad of its convs; tv resnet.py:89-105,143-163) after a train-mode forward. */
int r3m_b200_engine_num_blocks(void* handle, int* count);
int r3m_b200_engine_debug_block(void* handle, int block, int what, void** ptr, size_t* count);
int r3m_b200_engine_debug_run_block_backward(void* handle, int block, void* stream);

/* torch.optim.Adam step (models_r3m.py:76, defaults) on grads * grad_scale, then refresh of the bf16 operands.
 * step is the 1-based step count (bias correction).  Between update_grads and adam_step the caller may all-reduce
 * region 1 across ranks (trainer DDP path: ONE NCCL all-reduce per step). */
int r3m_b200_engine_adam_step(void* handle, float lr, float grad_scale, int step, void* stream);

/* Measurement aid for bench.py: runs ONE full step (update_grads + adam_step) with a CUDA-event pair recorded
 * in-stream around every kernel launch and returns, per kernel family f (0 conv_igemm [forward + dgrad], 1 wgrad,
 * 2 BatchNorm/normalise, 3 pooling, 4 loss heads, 5 optimiser + filter re-packs, 6 language head, 7 unused):
 * out32[4*f + {0,1,2,3}] = {device ms, algorithmic FLOPs, algorithmic HBM bytes, launches}.  out32 is HOST memory. */
int r3m_b200_engine_profile_update(void* handle, const void* obs, const int* perms, const float* lang_emb,
                                   const float* lang_mask, float l2weight, float l1weight, float langweight,
                                   float tcnweight, float lr, int step, double* out32, void* stream);

/* Per-launch records of the last r3m_b200_engine_profile_update, in launch order: out[4*i + {0,1,2,3}] =
 * {family, device ms, algorithmic FLOPs, algorithmic HBM bytes}.  out is HOST memory of 4*capacity_ops doubles. */
int r3m_b200_engine_profile_ops(void* handle, double* out, int capacity_ops, int* num_ops);
/* Human-readable label (layer / role) of launch `index` of the last profile. */
int r3m_b200_engine_profile_label(void* handle, int index, char* out, int capacity);

/* ------------------------------------------------------------------------------------------------------------
 * Sentence encoder: the frozen DistilBERT of the language branch (r3m/models/models_language.py:13-35:
 * AutoModel("distilbert-base-uncased")(input_ids, attention_mask).last_hidden_state.mean(1) under no_grad).
 * Tokenisation stays on the host (the reference's AutoTokenizer); everything from the token ids on runs here:
 * embeddings + LayerNorm, `layers` x [q/k/v Linear, softmax attention, out Linear + residual + LayerNorm, Linear +
 * exact GELU, Linear + residual + LayerNorm], mean over positions.  The Linears run on the tcgen05 kernel in its tf32
 * tier (fp32 storage, fp32 accumulation); everything else is fp32.
 * ------------------------------------------------------------------------------------------------------------ */

/* distilbert-base-uncased: vocab 30522, max_pos 512, dim 768, heads 12, layers 6, ffn 3072 (head size must be 64). */
int r3m_b200_distilbert_create(int vocab, int max_pos, int dim, int heads, int layers, int ffn, void** handle);
int r3m_b200_distilbert_destroy(void* handle);
/* Flat fp32 parameter buffer: element count, and the table of named tensors in transformers' state_dict naming
 * ("embeddings.word_embeddings.weight", "transformer.layer.0.attention.q_lin.weight", ...; Linear weights [out][in]). */
int r3m_b200_distilbert_num_params(void* handle, size_t* count);
int r3m_b200_distilbert_num_tensors(void* handle, int* count);
int r3m_b200_distilbert_tensor_info(void* handle, int index, char* name, int name_capacity, long long* offset, int* ndim,
                                    int* dims2);
/* The caller owns both device allocations: params (num_params floats, 16-byte aligned, filled from the checkpoint)
 * and a workspace (1024-byte aligned) sized for at most max_tokens = sentences x padded length per call. */
int r3m_b200_distilbert_workspace_bytes(void* handle, int max_tokens, size_t* bytes);
int r3m_b200_distilbert_bind(void* handle, float* params, void* workspace, size_t bytes, int max_tokens);
/* After (re)writing params: refresh the tf32-rounded operand copies. */
int r3m_b200_distilbert_sync_weights(void* handle, void* stream);
/* ids int32 [B][T], mask fp32 [B][T] (1 token, 0 padding; every sentence needs at least one token) ->
 * out fp32 [B][dim] = last_hidden_state.mean(1) (padding positions included, as the reference does);
 * hidden (optional) fp32 [B][T][dim] = last_hidden_state. */
int r3m_b200_distilbert_forward(void* handle, const int* ids, const float* mask, int B, int T, float* out, float* hidden,
                                void* stream);
/* Kernels launched by the last r3m_b200_distilbert_forward. */
int r3m_b200_distilbert_launches(void* handle, int* count);

#ifdef __cplusplus
}
#endif
#endif /* R3M_B200_H_ */
