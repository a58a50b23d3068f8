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

#ifdef __cplusplus
}
#endif
#endif /* R3M_B200_H_ */
