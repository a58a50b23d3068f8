"""r3m_b200 — B200-native (sm_100a) implementation of the R3M pretraining hot path.

Public surface mirrors the reference package (r3m/__init__.py): ``R3M`` (r3m/__init__.py:5), ``load_r3m``
(:44-75), ``load_r3m_reproduce`` (:77-113); the update step lives in ``r3m_b200.trainer.Trainer``
(reference r3m/trainer.py:21).  Importing the package requires the built shared library; there is no fallback.
"""
import copy
import os
from os.path import expanduser

import torch

from . import _lib  # noqa: F401  (fails loudly when libr3m_b200.so is missing)
from .model import R3M, set_lang_encoder_factory  # noqa: F401
from .trainer import Trainer  # noqa: F401
from .data import FrameFeeder, GpuAugment, R3MBufferU8, nvjpeg_batch_decoder  # noqa: F401
from .bert import DistilBertEncoder  # noqa: F401
from .checkpoint import load_snapshot, save_snapshot  # noqa: F401

__all__ = ["R3M", "Trainer", "load_r3m", "load_r3m_reproduce", "set_lang_encoder_factory", "FrameFeeder", "GpuAugment",
           "R3MBufferU8", "nvjpeg_batch_decoder", "DistilBertEncoder", "save_snapshot", "load_snapshot", "check_device"]


def check_device():
    """Synchronise the device and raise ``R3MB200Error`` if a tcgen05 / TMA pipeline watchdog fired in any kernel since
    the last check.  ``Trainer.update`` performs this check every step (the flag travels with the metrics read-back);
    forward-only use (``model.eval(); model(frames)``) enqueues its kernels without synchronising, so a serving loop
    calls this where a synchronisation is acceptable — e.g. once per batch after copying the embeddings out."""
    _lib.check(_lib.lib.r3m_b200_check_device_flag())

VALID_ARGS = ["_target_", "device", "lr", "hidden_dim", "size", "l2weight", "l1weight", "langweight", "tcnweight",
              "l2dist", "bs"]  # r3m/__init__.py:15

_MODELS = {  # r3m/__init__.py:46-57: folder under ~/.r3m, Google-Drive ids of model.pt / config.yaml
    "resnet50": ("r3m_50", "1Xu0ssuG0N1zjZS54wmWzJ7-nb0-7XzbA", "10jY2VxrrhfOdNPmsFdES568hjjIoBJx8"),
    "resnet34": ("r3m_34", "15bXD3QRhspIRacOKyWPw5y2HpoWUCEnE", "1RY0NS-Tl4G7M1Ik_lOym0b5VIBxX9dqW"),
    "resnet18": ("r3m_18", "1A1ic-p4KtYlKXdXHcV2QV0cUzI4kn0u-", "1nitbHQ-GRorxc7vMUiEHjHWP5N11Jvc6"),
}
_REPRODUCE = {  # r3m/__init__.py:79-94 (the reference's `modelif` typos make two of these unreachable there)
    "r3m": ("original_r3m", "1jLb1yldIMfAcGVwYojSQmMpmRM7vqjp9", "1cu-Pb33qcfAieRIUptNlG1AQIMZlAI-q"),
    "r3m_noaug": ("original_r3m_noaug", "1k_ZlVtvlktoYLtBcfD0aVFnrZcyCNS9D", "1hPmJwDiWPkd6GGez6ywSC7UOTIX7NgeS"),
    "r3m_nol1": ("original_r3m_nol1", "1LpW3aBMdjoXsjYlkaDnvwx7q22myM_nB", "1rZUBrYJZvlF1ReFwRidZsH7-xe7csvab"),
    "r3m_nolang": ("original_r3m_nolang", "1FXcniRei2JDaGMJJ_KlVxHaLy0Fs_caV", "192G4UkcNJO4EKN46ECujMcH0AQVhnyQe"),
}


def cleanup_config(agent_cfg):
    """r3m/__init__.py:21-33: keep the constructor arguments, drop the language head."""
    cfg = {k: v for k, v in copy.deepcopy(dict(agent_cfg)).items() if k in VALID_ARGS}
    cfg.pop("_target_", None)
    cfg["langweight"] = 0
    return cfg


def remove_language_head(state_dict):
    """r3m/__init__.py:35-42."""
    for key in list(state_dict.keys()):
        if ("lang_enc" in key) or ("lang_rew" in key):
            del state_dict[key]
    return state_dict


def load_config(path):
    """``OmegaConf.load`` (r3m/__init__.py:68) for the checkpoint's config.yaml without omegaconf: PyYAML plus the two
    things it does not do — hydra interpolations ``${key}`` / ``${a.b}`` resolved against the top-level config, and
    numbers written like ``1e-4`` (a string for YAML 1.1) coerced."""
    import re

    import yaml

    with open(path) as f:
        cfg = yaml.safe_load(f)

    def lookup(dotted):
        node = cfg
        for part in dotted.split("."):
            node = node[part]
        return node

    def resolve(v, depth=0):
        if isinstance(v, dict):
            return {k: resolve(x, depth) for k, x in v.items()}
        if isinstance(v, list):
            return [resolve(x, depth) for x in v]
        if isinstance(v, str):
            m = re.fullmatch(r"\$\{([^}]+)\}", v.strip())
            if m and depth < 8:
                try:
                    return resolve(lookup(m.group(1)), depth + 1)
                except (KeyError, TypeError):
                    raise ValueError(f"{path}: cannot resolve interpolation {v!r}") from None
            try:
                return int(v) if re.fullmatch(r"[+-]?\d+", v.strip()) else float(v)
            except ValueError:
                return v
        return v

    cfg = resolve(cfg)
    agent = cfg.get("agent", {})
    for key, typ in (("lr", float), ("hidden_dim", int), ("size", int)):
        if key in agent:
            try:
                agent[key] = typ(agent[key])
            except (TypeError, ValueError):
                raise ValueError(f"{path}: agent.{key} = {agent[key]!r} is not a {typ.__name__}") from None
    return cfg


def _load(table, modelid):
    if modelid not in table:
        raise NameError("Invalid Model ID")  # r3m/__init__.py:59
    foldername, model_id, config_id = table[modelid]
    home = os.path.join(expanduser("~"), ".r3m")
    os.makedirs(os.path.join(home, foldername), exist_ok=True)
    modelpath = os.path.join(home, foldername, "model.pt")
    configpath = os.path.join(home, foldername, "config.yaml")
    if not os.path.exists(modelpath):  # r3m/__init__.py:65-67
        try:
            import gdown
        except ImportError as e:
            raise RuntimeError(f"{modelpath} is not cached and gdown is not installed: place model.pt and "
                               f"config.yaml there, or install gdown") from e
        gdown.download("https://drive.google.com/uc?id=" + model_id, modelpath, quiet=False)
        gdown.download("https://drive.google.com/uc?id=" + config_id, configpath, quiet=False)
    cfg = cleanup_config(load_config(configpath)["agent"])
    device = "cuda" if torch.cuda.is_available() else "cpu"
    cfg["device"] = device
    rep = R3M(**cfg)
    # one GPU per process: with the default device_ids DataParallel would scatter a batch over every visible GPU and
    # replicate a module whose parameters alias ONE device block
    ids = [torch.cuda.current_device()] if device == "cuda" else None
    rep = torch.nn.DataParallel(rep.to(device), device_ids=ids)
    payload = torch.load(modelpath, map_location=torch.device(device), weights_only=False)["r3m"]
    rep.load_state_dict(remove_language_head(payload))
    return rep


def load_r3m(modelid):
    """r3m/__init__.py:44-75 -> DataParallel(R3M) with the cached checkpoint loaded."""
    return _load(_MODELS, modelid)


def load_r3m_reproduce(modelid):
    """r3m/__init__.py:77-113."""
    return _load(_REPRODUCE, modelid)
