"""Thin Python handle over the C-ABI engine (include/r3m_b200.h "Engine").  torch is used for device memory and
streams only; every kernel is launched from inside libr3m_b200.so."""
import ctypes

import torch

from . import _lib as L

METRIC_KEYS = ("l2loss", "l1loss", "l0loss", "rewloss", "rewacc1", "rewacc2", "rewacc3", "tcnloss", "aligned",
               "full_loss")  # slot order of region 7; key set of r3m/trainer.py:55-57,114-117,148-149,152

KIND_CONV, KIND_STEM, KIND_VECTOR, KIND_RUN_MEAN, KIND_RUN_VAR, KIND_LINEAR_W, KIND_LINEAR_B = range(7)


class TensorInfo:
    __slots__ = ("name", "kind", "offset", "shape")

    def __init__(self, name, kind, offset, shape):
        self.name, self.kind, self.offset, self.shape = name, kind, offset, shape

    @property
    def numel(self):
        n = 1
        for d in self.shape:
            n *= d
        return n


class Layout:
    """Parameter-block layout of one model (size, language head): needs no GPU."""

    def __init__(self, size, lang_head=False, hidden_dim=1024):
        h = ctypes.c_void_p()
        L.check(L.lib.r3m_b200_engine_create(size, 5, int(lang_head), hidden_dim, ctypes.byref(h)))
        try:
            n = ctypes.c_int()
            L.check(L.lib.r3m_b200_engine_num_tensors(h, ctypes.byref(n)))
            self.tensors = []
            buf = ctypes.create_string_buffer(256)
            for i in range(n.value):
                kind, ndim = ctypes.c_int(), ctypes.c_int()
                off = ctypes.c_longlong()
                dims = (ctypes.c_int * 4)()
                L.check(L.lib.r3m_b200_engine_tensor_info(h, i, buf, 256, ctypes.byref(kind), ctypes.byref(off),
                                                          ctypes.byref(ndim), dims))
                self.tensors.append(TensorInfo(buf.value.decode(), kind.value, off.value,
                                               tuple(dims[j] for j in range(ndim.value))))
            offs = (ctypes.c_size_t * 5)()
            npar, nbuf, nbytes = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
            L.check(L.lib.r3m_b200_engine_param_block_layout(h, offs, ctypes.byref(npar), ctypes.byref(nbuf)))
            L.check(L.lib.r3m_b200_engine_param_block_bytes(h, ctypes.byref(nbytes)))
            self.region_offsets = tuple(int(o) for o in offs)  # bytes: params, grads, m, v, buffers
            self.num_params = int(npar.value)
            self.num_buffer_floats = int(nbuf.value)
            self.param_block_bytes = int(nbytes.value)
        finally:
            L.lib.r3m_b200_engine_destroy(h)


def _aligned_empty(nbytes, device, zero):
    """uint8 tensor whose data pointer is 1024-byte aligned."""
    raw = (torch.zeros if zero else torch.empty)(nbytes + 1024, dtype=torch.uint8, device=device)
    shift = (-raw.data_ptr()) % 1024
    return raw[shift:shift + nbytes]


class Engine:
    """One launch schedule: (model parameter block, frame count)."""

    def __init__(self, size, frames, param_block, lang_head=False, hidden_dim=1024, l2dist=True):
        if not param_block.is_cuda:
            raise L.R3MB200Error("r3m_b200 computes on an sm_100 GPU only; the parameter block is on "
                                 f"{param_block.device} (there is no CPU path)")
        self.frames = frames
        self.device = param_block.device
        self._h = ctypes.c_void_p()
        L.check(L.lib.r3m_b200_engine_create(size, frames, int(lang_head), hidden_dim, ctypes.byref(self._h)))
        nbytes = ctypes.c_size_t()
        L.check(L.lib.r3m_b200_engine_workspace_bytes(self._h, ctypes.byref(nbytes)))
        self.workspace_bytes = int(nbytes.value)
        with torch.cuda.device(self.device):
            self._ws = _aligned_empty(self.workspace_bytes, self.device, zero=False)
            self._params = param_block
            L.check(L.lib.r3m_b200_engine_bind(self._h, L.ptr(param_block), param_block.numel(), L.ptr(self._ws),
                                               self._ws.numel(), L.current_stream()))
        d = ctypes.c_int()
        L.check(L.lib.r3m_b200_engine_get_int(self._h, 0, ctypes.byref(d)))
        self.embed_dim = d.value
        L.check(L.lib.r3m_b200_engine_set_int(self._h, 0, int(bool(l2dist))))  # R3M.sim: -L2 distance or cosine
        self._metrics_host = torch.empty(16, dtype=torch.float32).pin_memory()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            L.lib.r3m_b200_engine_destroy(h)

    def _region(self, which, dtype=torch.float32):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        L.check(L.lib.r3m_b200_engine_region(self._h, which, ctypes.byref(p), ctypes.byref(n)))
        return p.value, int(n.value)

    def launches(self):
        v = ctypes.c_int()
        L.check(L.lib.r3m_b200_engine_get_int(self._h, 2, ctypes.byref(v)))
        return v.value

    def graph_replays(self):
        """cudaGraphLaunch calls so far (the training step replays two captured graphs; R3M_STEP_GRAPH=0 disables)."""
        v = ctypes.c_int()
        L.check(L.lib.r3m_b200_engine_get_int(self._h, 3, ctypes.byref(v)))
        return v.value

    def sync_weights(self):
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_sync_weights(self._h, L.current_stream()))

    def _set_format(self, obs, nhwc=False):
        """Frames are consumed as they are: float32 NCHW (the reference loader's contract), uint8 NCHW or uint8 NHWC
        (``nhwc=True``: [frames,224,224,3]) — the ``obs.float()`` of models_r3m.py:97 happens inside the kernel."""
        assert obs.is_cuda and obs.is_contiguous() and obs.numel() == self.frames * 3 * 224 * 224, tuple(obs.shape)
        if obs.dtype == torch.float32 and not nhwc:
            fmt = 0
        elif obs.dtype == torch.uint8:
            fmt = 2 if nhwc else 1
        else:
            raise L.R3MB200Error(f"frames must be float32 NCHW or uint8 (got {obs.dtype}, nhwc={nhwc})")
        if fmt != getattr(self, "_obs_format", 0):
            L.check(L.lib.r3m_b200_engine_set_int(self._h, 1, fmt))
            self._obs_format = fmt

    def set_precision(self, tier):
        """Tier of the EVAL-mode forward: "bf16" (default) or "tf32" (fp32 storage, kind::tf32 tensor cores)."""
        L.check(L.lib.r3m_b200_engine_set_int(self._h, 2, {"bf16": 0, "tf32": 1}[tier]))

    def forward(self, obs, train, nhwc=False):
        """obs: contiguous CUDA [frames,3,224,224] float32 or uint8 ([frames,224,224,3] uint8 with nhwc) in [0,255]
        -> new float32 [frames, D]."""
        self._set_format(obs, nhwc)
        out = torch.empty(self.frames, self.embed_dim, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_forward(self._h, L.ptr(obs), int(train), L.ptr(out), L.current_stream()))
        return out

    def forward_train_async(self, obs, nhwc=False):
        """Enqueue the train-mode forward only (embeddings stay in the engine); pair with update_grads(obs=None)."""
        self._set_format(obs, nhwc)
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_forward(self._h, L.ptr(obs), 1, None, L.current_stream()))

    def update_grads(self, obs, perms, lang_emb, lang_mask, l2w, l1w, langw, tcnw, eval_mode, nhwc=False):
        """obs None: the forward pass was enqueued with forward_train_async(obs) (training only)."""
        if obs is not None:
            self._set_format(obs, nhwc)
        assert perms.dtype == torch.int32 and perms.is_cuda and perms.is_contiguous()
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_update_grads(self._h, L.ptr(obs), L.ptr(perms), L.ptr(lang_emb),
                                                       L.ptr(lang_mask), l2w, l1w, langw, tcnw, int(eval_mode),
                                                       L.current_stream()))

    def adam_step(self, lr, grad_scale, step):
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_adam_step(self._h, lr, grad_scale, step, L.current_stream()))

    def backward(self, dE):
        """Backward pass alone: dE float32 [frames, D] = d(loss)/d(embeddings) of the preceding train-mode forward.
        Accumulates filter gradients into the gradient region (see r3m_b200_engine_backward)."""
        assert dE.dtype == torch.float32 and dE.is_cuda and dE.is_contiguous()
        assert dE.numel() == self.frames * self.embed_dim
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_backward(self._h, L.ptr(dE), L.current_stream()))

    def grad_chunks(self):
        """[(begin, end)] element ranges of the flat gradient buffer in the order the backward pass completes them."""
        n = ctypes.c_int()
        L.check(L.lib.r3m_b200_engine_num_grad_chunks(self._h, ctypes.byref(n)))
        out = []
        for k in range(n.value):
            b, e = ctypes.c_size_t(), ctypes.c_size_t()
            L.check(L.lib.r3m_b200_engine_grad_chunk(self._h, k, ctypes.byref(b), ctypes.byref(e)))
            out.append((int(b.value), int(e.value)))
        return out

    def wait_grad_chunk(self, k, stream):
        """Make `stream` (torch.cuda.Stream) wait until chunk k of the last enqueued update_grads / backward is final."""
        L.check(L.lib.r3m_b200_engine_wait_grad_chunk(self._h, k, ctypes.c_void_p(stream.cuda_stream)))

    # ---- test hooks (tests/test_block_backward_gpu.py)
    def num_blocks(self):
        v = ctypes.c_int()
        L.check(L.lib.r3m_b200_engine_num_blocks(self._h, ctypes.byref(v)))
        return v.value

    def block_buffer(self, block, what):
        """bf16 alias of a residual block's buffer: 0 input, 1 output, 2 incoming gradient, 3 outgoing gradient."""
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        L.check(L.lib.r3m_b200_engine_debug_block(self._h, block, what, ctypes.byref(p), ctypes.byref(n)))
        with torch.cuda.device(self.device):
            raw = torch.as_tensor(_CudaArrayView(p.value, int(n.value), "<u2"), device=self.device)
        return raw.view(torch.bfloat16)

    def run_block_backward(self, block):
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_debug_run_block_backward(self._h, block, L.current_stream()))

    FAMILIES = ("conv_igemm", "wgrad", "norm", "pool", "loss", "optim", "lang", "other")

    def profile_update(self, obs, perms, lang_emb, lang_mask, l2w, l1w, langw, tcnw, lr, step):
        """One instrumented step -> {family: {ms, flops, bytes, launches}} (see r3m_b200_engine_profile_update)."""
        out = (ctypes.c_double * 32)()
        self._set_format(obs)
        with torch.cuda.device(self.device):
            L.check(L.lib.r3m_b200_engine_profile_update(self._h, L.ptr(obs), L.ptr(perms), L.ptr(lang_emb),
                                                         L.ptr(lang_mask), l2w, l1w, langw, tcnw, lr, step, out,
                                                         L.current_stream()))
        return {name: {"ms": out[4 * i], "flops": out[4 * i + 1], "bytes": out[4 * i + 2],
                       "launches": int(out[4 * i + 3])} for i, name in enumerate(self.FAMILIES)}

    def profile_ops(self):
        """Per-launch (family, ms, flops, bytes) of the last profile_update, in launch order."""
        cap = 4096
        out = (ctypes.c_double * (4 * cap))()
        n = ctypes.c_int()
        L.check(L.lib.r3m_b200_engine_profile_ops(self._h, out, cap, ctypes.byref(n)))
        buf = ctypes.create_string_buffer(160)
        rows = []
        for i in range(min(n.value, cap)):
            L.check(L.lib.r3m_b200_engine_profile_label(self._h, i, buf, 160))
            rows.append((self.FAMILIES[int(out[4 * i])], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3],
                         buf.value.decode()))
        return rows

    def read_metrics(self):
        """ONE device->host copy of the 16-float metrics buffer (the reference does ~10 .item() syncs)."""
        p, n = self._region(7)
        dev = _as_tensor(p, n, torch.float32, self.device)
        self._metrics_host.copy_(dev, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        vals = self._metrics_host.tolist()
        if vals[15] != 0.0:  # slot 15 carries the device-side pipeline watchdog flag of the step's tcgen05 kernels
            L.lib.r3m_b200_check_device_flag()  # clears it
            raise L.R3MB200Error(f"a tcgen05/TMA pipeline watchdog fired during the step (code {int(vals[15])}); "
                                 "the step's results are invalid")
        return vals

    def embeddings(self):
        p, n = self._region(5)
        return _as_tensor(p, n, torch.float32, self.device).view(self.frames, self.embed_dim)

    def embedding_grads(self):
        p, n = self._region(6)
        return _as_tensor(p, n, torch.float32, self.device).view(self.frames, self.embed_dim)


class _CudaArrayView:
    """__cuda_array_interface__ carrier so torch can alias raw device memory owned by the engine's allocations."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def _as_tensor(ptr, n, dtype, device):
    assert dtype == torch.float32
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaArrayView(ptr, n, "<f4"), device=device)
