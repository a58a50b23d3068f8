"""Exact-resume snapshots, a superset of the reference's (r3m/train_representation.py:123-138).

The reference saves ``{"r3m": model.state_dict(), "global_step": step}`` and restores exactly that, so a resumed run
restarts Adam from zero moments and an unrelated random stream.  ``save_snapshot`` writes the SAME two entries — the
reference's ``Workspace.load_snapshot`` and ``load_r3m``-style loaders keep working on the file — plus what an exact
resume needs: Adam's moments and step count, torch's CPU / CUDA generators and the ``random`` / ``numpy`` streams the
loader draws its clips from.  Under one-process-per-GPU only rank 0 writes (every rank holds identical weights after
the all-reduced step); every rank reads."""
import os
import random

import numpy as np
import torch

FORMAT = 1


def _rank():
    import torch.distributed as dist

    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _module(model):
    return model.module if isinstance(model, torch.nn.DataParallel) else model


def save_snapshot(path, model, global_step, extra=None):
    """model: ``DataParallel(R3M)`` (what the reference's Workspace holds) or a bare ``R3M``.  Returns True on the
    writing rank.  The file is written atomically (temp file + rename)."""
    if _rank() != 0:
        return False
    m = _module(model)
    sdict = {"r3m": model.state_dict(), "global_step": int(global_step)}
    dev = m._block.device
    sdict["r3m_b200"] = {
        "format": FORMAT,
        "encoder_opt": {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in m.encoder_opt.state_dict().items()},
        "rng": {"torch_cpu": torch.get_rng_state(),
                "torch_cuda": torch.cuda.get_rng_state(dev) if dev.type == "cuda" else None,
                "python": random.getstate(), "numpy": np.random.get_state()},
        "extra": extra,
    }
    tmp = f"{path}.tmp.{os.getpid()}"
    torch.save(sdict, tmp)
    os.replace(tmp, path)
    return True


def load_snapshot(path, model, restore_rng=True):
    """Inverse of ``save_snapshot``; also accepts a snapshot written by the reference (weights + global_step only:
    Adam then restarts from zero, as it does there).  Returns ``(global_step, extra)``."""
    m = _module(model)
    payload = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict(payload["r3m"])
    step = int(payload.get("global_step", 0))
    ours = payload.get("r3m_b200")
    if ours is None:
        return step, None
    if ours.get("format") != FORMAT:
        raise ValueError(f"{path}: unknown r3m_b200 snapshot format {ours.get('format')!r}")
    m.encoder_opt.load_state_dict(ours["encoder_opt"])
    if restore_rng:
        rng = ours["rng"]
        torch.set_rng_state(rng["torch_cpu"])
        if rng.get("torch_cuda") is not None and m._block.device.type == "cuda":
            torch.cuda.set_rng_state(rng["torch_cuda"], m._block.device)
        random.setstate(rng["python"])
        np.random.set_state(rng["numpy"])
    return step, ours.get("extra")
