"""ctypes binding of libr3m_b200.so (the C ABI declared in include/r3m_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this module on a machine where the shared
library has not been built raises ImportError; calling a compute entry point without an sm_100 GPU raises
RuntimeError with the library's own message.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libr3m_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C r3m_b200/csrc` (nvcc, sm_100a). r3m_b200 has no fallback path."
    )

lib = ctypes.CDLL(LIB_PATH)

c_int = ctypes.c_int
c_void_p = ctypes.c_void_p
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t

lib.r3m_b200_last_error.restype = ctypes.c_char_p
lib.r3m_b200_last_error.argtypes = []


class R3MB200Error(RuntimeError):
    pass


def check(code):
    if code != 0:
        msg = lib.r3m_b200_last_error().decode("utf-8", "replace")
        raise R3MB200Error(f"r3m_b200 error {code}: {msg}")


def _sig(name, argtypes):
    fn = getattr(lib, name)
    fn.restype = c_int
    fn.argtypes = argtypes
    return fn


_sig("r3m_b200_abi_version", [])
_sig("r3m_b200_check_device_flag", [])
_sig("r3m_b200_conv_fwd", [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_void_p, c_void_p, c_void_p])
_sig("r3m_b200_conv_fwd_affine", [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_void_p, c_void_p, c_void_p, c_int,
                                                                                c_void_p])
_sig("r3m_b200_pack_dgrad_filter", [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p])
_sig("r3m_b200_conv_dgrad", [c_void_p, c_void_p, c_void_p] + [c_int] * 10 + [c_void_p])
_sig("r3m_b200_conv_wgrad", [c_void_p, c_void_p, c_void_p] + [c_int] * 9 + [c_void_p])
_sig("r3m_b200_preprocess_stem", [c_void_p, c_void_p, c_int, c_void_p])
_sig("r3m_b200_preprocess_stem_format", [c_void_p, c_int, c_void_p, c_int, c_void_p])
_sig("r3m_b200_random_resized_crop", [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p])
_sig("r3m_b200_bn_apply", [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 19)
_sig("r3m_b200_bn_backward", [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 17)
_sig("r3m_b200_stem_bn_relu_maxpool", [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int] + [c_void_p] * 9)
_sig("r3m_b200_maxpool_backward", [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p])
_sig("r3m_b200_loss_tcn_sim", [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p])
_sig("r3m_b200_engine_set_int", [c_void_p, c_int, c_int])
_sig("r3m_b200_ordered_sum", [c_void_p, c_size_t, c_void_p, c_int, c_void_p])
_sig("r3m_b200_ordered_moments", [c_void_p, c_int, c_void_p, c_int, c_void_p])
_sig("r3m_b200_pull_host", [c_void_p, c_void_p, ctypes.c_size_t, c_void_p])
_sig("r3m_b200_stem_backward", [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 8)
_sig("r3m_b200_avgpool_forward", [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p])
_sig("r3m_b200_avgpool_backward", [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p])
_sig("r3m_b200_loss_lp", [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p])
_sig("r3m_b200_loss_tcn", [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p])
_sig("r3m_b200_adam", [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_int, c_float, c_void_p])

c_void_pp = ctypes.POINTER(c_void_p)
c_size_p = ctypes.POINTER(c_size_t)
c_int_p = ctypes.POINTER(c_int)
_sig("r3m_b200_engine_create", [c_int, c_int, c_int, c_int, c_void_pp])
_sig("r3m_b200_engine_destroy", [c_void_p])
_sig("r3m_b200_engine_workspace_bytes", [c_void_p, c_size_p])
_sig("r3m_b200_engine_param_block_bytes", [c_void_p, c_size_p])
_sig("r3m_b200_engine_bind", [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p])
_sig("r3m_b200_engine_num_tensors", [c_void_p, c_int_p])
_sig("r3m_b200_engine_tensor_info", [c_void_p, c_int, ctypes.c_char_p, c_int, c_int_p,
                                     ctypes.POINTER(ctypes.c_longlong), c_int_p, c_int_p])
_sig("r3m_b200_engine_region", [c_void_p, c_int, c_void_pp, c_size_p])
_sig("r3m_b200_engine_get_int", [c_void_p, c_int, c_int_p])
_sig("r3m_b200_engine_param_block_layout", [c_void_p, c_size_p, c_size_p, c_size_p])
_sig("r3m_b200_engine_sync_weights", [c_void_p, c_void_p])
_sig("r3m_b200_engine_forward", [c_void_p, c_void_p, c_int, c_void_p, c_void_p])
_sig("r3m_b200_engine_update_grads", [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                      c_float, c_int, c_void_p])
_sig("r3m_b200_engine_adam_step", [c_void_p, c_float, c_float, c_int, c_void_p])
_sig("r3m_b200_engine_backward", [c_void_p, c_void_p, c_void_p])
_sig("r3m_b200_engine_num_blocks", [c_void_p, c_int_p])
_sig("r3m_b200_engine_num_grad_chunks", [c_void_p, c_int_p])
_sig("r3m_b200_engine_grad_chunk", [c_void_p, c_int, c_size_p, c_size_p])
_sig("r3m_b200_engine_wait_grad_chunk", [c_void_p, c_int, c_void_p])
_sig("r3m_b200_engine_debug_block", [c_void_p, c_int, c_int, c_void_pp, c_size_p])
_sig("r3m_b200_engine_debug_run_block_backward", [c_void_p, c_int, c_void_p])
_sig("r3m_b200_engine_profile_ops", [c_void_p, ctypes.POINTER(ctypes.c_double), c_int, c_int_p])
_sig("r3m_b200_engine_profile_label", [c_void_p, c_int, ctypes.c_char_p, c_int])
_sig("r3m_b200_engine_profile_update", [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                        c_float, c_float, c_int, ctypes.POINTER(ctypes.c_double), c_void_p])

c_float_p = ctypes.POINTER(c_float)
_sig("r3m_b200_distilbert_create", [c_int] * 6 + [c_void_pp])
_sig("r3m_b200_distilbert_destroy", [c_void_p])
_sig("r3m_b200_distilbert_num_params", [c_void_p, c_size_p])
_sig("r3m_b200_distilbert_num_tensors", [c_void_p, c_int_p])
_sig("r3m_b200_distilbert_tensor_info", [c_void_p, c_int, ctypes.c_char_p, c_int, ctypes.POINTER(ctypes.c_longlong),
                                         c_int_p, c_int_p])
_sig("r3m_b200_distilbert_workspace_bytes", [c_void_p, c_int, c_size_p])
_sig("r3m_b200_distilbert_bind", [c_void_p, c_void_p, c_void_p, c_size_t, c_int])
_sig("r3m_b200_distilbert_sync_weights", [c_void_p, c_void_p])
_sig("r3m_b200_distilbert_forward", [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p])
_sig("r3m_b200_distilbert_launches", [c_void_p, c_int_p])


def ptr(t):
    """Device (or host) address of a torch tensor, or None."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
