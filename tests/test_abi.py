"""The C-ABI shared library: loads without a GPU, exports every symbol include/r3m_b200.h declares, answers layout
queries on the host, and refuses compute without a device instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "r3m_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(r3m_b200_\w+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib.lib, n)]
    assert not missing, missing


def test_abi_version_and_error_string(lib):
    assert lib.lib.r3m_b200_abi_version() == 1
    assert isinstance(lib.lib.r3m_b200_last_error(), bytes)


def test_engine_layout_without_gpu(lib):
    from r3m_b200.engine import KIND_CONV, KIND_LINEAR_W, KIND_STEM, Layout

    for size, nparams in ((18, 11176512), (34, 21284672), (50, 23508032)):
        lay = Layout(size)
        assert sum(t.numel for t in lay.tensors if t.kind in (0, 1, 2)) == nparams
        assert lay.num_params >= nparams and lay.param_block_bytes > 18 * nparams
        assert [t for t in lay.tensors if t.kind == KIND_STEM][0].shape == (64, 3, 7, 7)
        offs = sorted((t.offset, t.numel) for t in lay.tensors if t.kind in (0, 1, 2))
        assert all(o0 + n0 <= o1 for (o0, n0), (o1, _) in zip(offs, offs[1:]))  # no overlap
    lay = Layout(50, lang_head=True, hidden_dim=1024)
    lin = [t for t in lay.tensors if t.kind == KIND_LINEAR_W]
    assert [t.shape for t in lin] == [(1024, 4864), (1024, 1024), (1024, 1024), (1024, 1024), (1, 1024)]
    assert sum(t.numel for t in lay.tensors if t.kind in (5, 6)) == 8131585
    assert any(t.kind == KIND_CONV and t.name == "convnet.layer4.2.conv3.weight" for t in lay.tensors)


def test_invalid_arguments_are_errors_not_crashes(lib):
    h = ctypes.c_void_p()
    assert lib.lib.r3m_b200_engine_create(101, 5, 0, 1024, ctypes.byref(h)) < 0
    assert b"18, 34 or 50" in lib.lib.r3m_b200_last_error()
    assert lib.lib.r3m_b200_engine_create(18, 0, 0, 1024, ctypes.byref(h)) < 0
    # a model with a language head embeds any number of frames, like the reference (only update() needs 5 * clips)
    assert lib.lib.r3m_b200_engine_create(18, 7, 1, 1024, ctypes.byref(h)) == 0
    assert lib.lib.r3m_b200_engine_destroy(h) == 0
    n = ctypes.c_size_t()
    assert lib.lib.r3m_b200_engine_workspace_bytes(None, ctypes.byref(n)) < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point fails loudly."""
    assert lib.lib.r3m_b200_check_device_flag() < 0
    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16)
    w = torch.zeros(64, 1, 1, 64, dtype=torch.bfloat16)
    y = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16)
    rc = lib.lib.r3m_b200_conv_fwd(lib.ptr(x), lib.ptr(w), lib.ptr(y), 1, 8, 8, 64, 64, 1, 1, 1, 0, None, None, None)
    assert rc < 0 and len(lib.lib.r3m_b200_last_error()) > 0
    from r3m_b200 import R3M
    from r3m_b200._lib import R3MB200Error

    m = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0)
    with pytest.raises(R3MB200Error):
        m(torch.zeros(1, 3, 224, 224))
