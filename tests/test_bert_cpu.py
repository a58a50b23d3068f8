"""Host-side checks of the sentence-encoder boundary (no GPU): the flat parameter table is transformers'
DistilBertModel.state_dict() — same names, shapes and total — and compute is refused without a device."""
import ctypes

import pytest
import torch


def test_tensor_table_is_the_transformers_state_dict():
    from transformers import DistilBertConfig, DistilBertModel

    from r3m_b200.bert import DistilBertLayout

    lay = DistilBertLayout()
    with torch.device("meta"):
        sd = DistilBertModel(DistilBertConfig()).state_dict()
    assert set(lay.tensors) == set(sd)
    assert all(tuple(sd[k].shape) == shape for k, (_, shape) in lay.tensors.items())
    assert lay.num_params == sum(v.numel() for v in sd.values()) == 66362880
    spans = sorted((off, off + int(torch.tensor(shape).prod())) for off, shape in lay.tensors.values())
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == lay.num_params


def test_bad_dimensions_are_rejected(lib):
    h = ctypes.c_void_p()
    assert lib.lib.r3m_b200_distilbert_create(30522, 512, 768, 8, 6, 3072, ctypes.byref(h)) != 0  # head size != 64
    assert b"head size" in lib.lib.r3m_b200_last_error()
    assert lib.lib.r3m_b200_distilbert_create(30522, 512, 700, 12, 6, 3072, ctypes.byref(h)) != 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_path():
    from r3m_b200._lib import R3MB200Error
    from r3m_b200.bert import DistilBertEncoder

    with pytest.raises(R3MB200Error):
        DistilBertEncoder({}, "cpu")
