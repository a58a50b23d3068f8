"""Host-side laws of the input pipeline against what the REAL reference loader drew
(tests/golden/loader_law.json, recorded by oracle/make_golden.py from r3m.utils.data_loaders.R3MBuffer with the JPEG
reader patched out): clip indices and labels (data_loaders.py:64-79), RandomResizedCrop boxes (:47-50,81-102)."""
import json
import os
import random

import numpy as np
import pytest
import torch

from r3m_b200.data import R3MBufferU8, draw_crop_boxes, random_resized_crop_params

GOLD = os.path.join(os.path.dirname(__file__), "golden", "loader_law.json")


def _cases():
    with open(GOLD) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["doaug"])
def test_clip_indices_labels_and_boxes_match_the_reference_loader(case):
    import pandas as pd

    lens = case["lens"]
    manifest = pd.DataFrame({"path": [f"/vid{i}" for i in range(len(lens))], "len": lens,
                             "txt": [f"C does thing {i}" for i in range(len(lens))]})
    seen = []

    def decoder(path):
        vid, name = path.rsplit("/", 1)
        seen.append([vid, int(name.split(".")[0])])
        return torch.zeros(3, case["H"], case["W"], dtype=torch.uint8)

    random.seed(case["seed"])
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    buf = R3MBufferU8("unused/", case["alpha"], ["ego4d"], manifest=manifest, decoder=decoder)
    labels, boxes = [], []
    for _ in range(6):
        im, label = buf._sample()
        assert im.shape == (5, 3, case["H"], case["W"]) and im.dtype == torch.uint8
        labels.append(label)
        if case["doaug"] != "none":
            b = draw_crop_boxes(1, case["H"], case["W"], case["doaug"])
            boxes += b.tolist()[::5] if case["doaug"] == "rctraj" else b.tolist()
    assert seen == case["frames"]
    assert labels == case["labels"]
    assert boxes == case["boxes"]


def test_crop_params_follow_torchvision_including_the_fallback():
    tv = pytest.importorskip("torchvision.transforms")
    for seed in range(40):
        H, W = [(224, 224), (480, 640), (300, 200), (1080, 1920), (30, 400), (400, 30)][seed % 6]
        torch.manual_seed(seed)
        want = tv.RandomResizedCrop.get_params(torch.zeros(3, H, W), (0.2, 1.0), (3 / 4, 4 / 3))
        after_want = torch.rand(1).item()
        torch.manual_seed(seed)
        got = random_resized_crop_params(H, W)
        assert tuple(got) == tuple(want) and torch.rand(1).item() == after_want  # same box, same generator state


def test_invalid_dataset_and_aug_names():
    with pytest.raises(NameError):
        R3MBufferU8("x/", 0.2, ["kinetics"], manifest=[1])
    with pytest.raises(ValueError):
        draw_crop_boxes(1, 224, 224, "none")
