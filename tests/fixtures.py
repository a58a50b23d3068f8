"""Seeded inputs of the golden cases (oracle/make_golden.py), shared by the CPU and GPU tests."""
import json
import os

import numpy as np
import torch

from oracle import r3m_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HYPER = dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, lr=1e-4)
FRAME_KINDS = {"randint": O.synthetic_frames, "structured": O.structured_frames, "varied": O.varied_frames}


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def case_inputs(case):
    """(params, buffers, frames, perms, lang_emb, sentences, mask) of a case dict {size, clips, langweight, frames,
    seeds[, last_gamma]} — exactly what oracle/make_golden.py fed the reference."""
    sw, sf, sp, sl = case["seeds"]
    lang = case["langweight"] > 0
    params, buffers = O.init_state(case["size"], sw, lang=lang)
    if "last_gamma" in case:
        O.scale_last_gamma(params, case["size"], case["last_gamma"])
    frames = FRAME_KINDS[case["frames"]](case["clips"], sf)
    perms = O.draw_permutations(case["clips"], sp)
    lang_emb = O.stub_lang_embedding(case["clips"], sl) if lang else None
    sentences = ["C does something %d" % i for i in range(case["clips"])]
    if lang and case["clips"] >= 4:
        sentences[1] = ""
    if not lang:
        sentences = [""] * case["clips"]
    mask = torch.tensor([1.0 * (s != "") for s in (sentences if lang else ["x"] * case["clips"])])
    return params, buffers, frames, perms, lang_emb, sentences, mask


def load_case(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    case = json.loads(bytes(z["case_json"]).decode())
    gold_metrics = json.loads(bytes(z["metrics_json"]).decode())
    return (z, case, gold_metrics) + case_inputs(case)


def projections(grads, names, rows=8, seed=99):
    """The fixed random 1-D projections stored for the well-conditioned goldens (gproj::<name>), recomputed for
    `grads` (same generator stream as oracle/make_golden.py: tensors visited in sorted-name order)."""
    pg = torch.Generator().manual_seed(seed)
    out = {}
    for k in names:
        g = grads[k].detach().float().cpu()
        if g.dim() > 1:
            proj = torch.randn(rows, g.numel(), generator=pg)
            out[k] = proj @ g.flatten()
    return out
