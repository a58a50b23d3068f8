"""Host-side mirror of the reference surface (no GPU): state_dict contract, loader, permutation stream, config."""
import os
import sys

import pytest
import torch
import torchvision

import r3m_b200
from r3m_b200 import R3M, Trainer
from r3m_b200.trainer import draw_permutations


class _StubEnc:
    lang_size = 768

    def __init__(self, device):
        pass

    def __call__(self, s):
        return torch.zeros(len(s), 768)


@pytest.mark.parametrize("size", [18, 34, 50])
def test_state_dict_matches_torchvision(size):
    """module.convnet.* keys, shapes, dtypes identical to torchvision's ResNet (SURVEY.md §5 checkpoint contract)."""
    m = R3M("cpu", 1e-4, 1024, size=size, langweight=0.0)
    tv = getattr(torchvision.models, f"resnet{size}")()
    tv.fc = torch.nn.Identity()
    want = {"convnet." + k: v for k, v in tv.state_dict().items()}
    got = m.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype, k
        assert got[k].is_contiguous()
    # reference checkpoints load (strict) and survive a round trip bit-exactly
    m.load_state_dict(want)
    back = m.state_dict()
    assert all(torch.equal(back[k], want[k]) for k in want)
    # and ours load into torchvision
    tv.load_state_dict({k[len("convnet."):]: v for k, v in back.items()})
    dp = torch.nn.DataParallel(m)
    assert next(iter(dp.state_dict())).startswith("module.convnet.")


def test_language_head_keys_and_init_laws():
    r3m_b200.set_lang_encoder_factory(_StubEnc)
    m = R3M("cpu", 1e-4, 1024, size=50, langweight=1.0)
    sd = m.state_dict()
    for i, shape in zip((0, 2, 4, 6, 8), ((1024, 4864), (1024, 1024), (1024, 1024), (1024, 1024), (1, 1024))):
        assert sd[f"lang_rew.pred.{i}.weight"].shape == shape
        assert sd[f"lang_rew.pred.{i}.bias"].shape == (shape[0],)
    w = sd["convnet.layer1.0.conv2.weight"]  # kaiming_normal_(fan_out, relu): std = sqrt(2 / (Cout*k*k))
    assert abs(float(w.std()) - (2.0 / (64 * 9)) ** 0.5) < 0.1 * (2.0 / (64 * 9)) ** 0.5
    assert torch.all(sd["convnet.bn1.weight"] == 1) and torch.all(sd["convnet.bn1.bias"] == 0)
    assert torch.all(sd["convnet.bn1.running_var"] == 1) and torch.all(sd["convnet.bn1.running_mean"] == 0)
    b = 1.0 / 4864 ** 0.5
    assert float(sd["lang_rew.pred.0.weight"].abs().max()) <= b
    assert sum(p.numel() for p in m.parameters()) == 23508032 + 8131585
    assert m.num_negatives == 3 and m.outdim == 2048


def test_constructor_signature_and_attributes():
    import inspect

    sig = inspect.signature(R3M.__init__)
    assert list(sig.parameters)[1:] == ["device", "lr", "hidden_dim", "size", "l2weight", "l1weight", "langweight",
                                        "tcnweight", "l2dist", "bs"]  # models_r3m.py:22-23
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["size"], d["l2weight"], d["l1weight"], d["langweight"], d["tcnweight"], d["l2dist"], d["bs"]) == \
        (34, 1.0, 1.0, 1.0, 0.0, True, 16)
    with pytest.raises(NameError):
        R3M("cpu", 1e-4, 1024, size=101, langweight=0.0)
    m = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0, tcnweight=1.0)
    a, b = torch.randn(3, 8), torch.randn(3, 8)
    assert torch.allclose(m.sim(a, b), -torch.linalg.norm(a - b, dim=-1))
    assert hasattr(m.encoder_opt, "zero_grad") and hasattr(m.encoder_opt, "step")


def test_parameters_alias_the_flat_block_and_survive_to():
    m = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0)
    w = m.convnet.layer1._modules["0"].conv1.weight
    flat = m._flat(0)
    with torch.no_grad():
        w[3, 5, 1, 2] = 42.0
    assert (flat == 42.0).sum() == 1  # the OIHW view writes through to the KRSC master copy
    assert w.grad is not None and w.grad.shape == w.shape
    m.to("cpu")
    m.float()
    assert m.convnet.layer1._modules["0"].conv1.weight[3, 5, 1, 2] == 42.0


def test_permutation_stream_matches_reference_draw_order():
    """trainer.py:86-92 then :135-137: 9 language permutations (if on) then 6 TCN ones from the global CPU RNG."""
    torch.manual_seed(7)
    want = [torch.randperm(6) for _ in range(15)]
    torch.manual_seed(7)
    got = draw_permutations(6, 1.0, 1.0)
    assert all(torch.equal(got[i].long(), want[i]) for i in range(15))
    torch.manual_seed(7)
    got = draw_permutations(6, 0.0, 1.0)  # language off: the TCN draws come first in the stream
    assert all(torch.equal(got[9 + i].long(), want[i]) for i in range(6))
    assert got.dtype == torch.int32 and got.shape == (15, 6)


def test_load_r3m_offline_from_cache(tmp_path, monkeypatch):
    """r3m/__init__.py:61-74: an existing ~/.r3m/<folder>/model.pt is used without downloading; the language head is
    stripped; invalid ids raise NameError."""
    monkeypatch.setenv("HOME", str(tmp_path))
    folder = tmp_path / ".r3m" / "r3m_18"
    folder.mkdir(parents=True)
    src = torch.nn.DataParallel(R3M("cpu", 1e-4, 1024, size=18, langweight=0.0))
    sd = src.state_dict()
    sd["module.lang_rew.pred.0.weight"] = torch.zeros(3, 3)  # must be ignored (remove_language_head)
    torch.save({"r3m": sd, "global_step": 5}, folder / "model.pt")
    (folder / "config.yaml").write_text(
        "agent:\n  _target_: r3m.R3M\n  device: cuda\n  lr: 0.0001\n  hidden_dim: 1024\n  size: 18\n  l2weight: 0.00001\n"
        "  l1weight: 0.00001\n  tcnweight: 1.0\n  langweight: 1.0\n  l2dist: true\n  bs: 32\n  extra_key: 1\n")
    rep = r3m_b200.load_r3m("resnet18")
    assert isinstance(rep, torch.nn.DataParallel) and rep.module.langweight == 0
    got = rep.state_dict()
    assert all(torch.equal(got[k], v) for k, v in src.state_dict().items())
    with pytest.raises(NameError):
        r3m_b200.load_r3m("resnet101")


def test_trainer_contract_needs_gpu_but_keeps_signature():
    import inspect

    params = list(inspect.signature(Trainer.update).parameters)
    assert params[:5] == ["self", "model", "batch", "step", "eval"]
    assert Trainer(eval_freq=20000).eval_freq == 20000


def test_optimizer_state_roundtrip_for_exact_resume():
    """the reference does not checkpoint Adam state; ours exposes it (encoder_opt.state_dict) so a
    resumed run continues bit-exactly, while the model state_dict stays loadable by the reference."""
    m = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0)
    m._flat(2).normal_()
    m._flat(3).uniform_()
    m.encoder_opt.steps = 7
    sd = m.encoder_opt.state_dict()
    m2 = R3M("cpu", 3e-4, 1024, size=18, langweight=0.0)
    m2.encoder_opt.load_state_dict(sd)
    assert m2.encoder_opt.steps == 7 and m2.encoder_opt.param_groups[0]["lr"] == 1e-4
    assert torch.equal(m2._flat(2), m._flat(2)) and torch.equal(m2._flat(3), m._flat(3))
    m.encoder_opt.zero_grad()
    assert float(m._flat(1).abs().max()) == 0.0


def test_resize_center_crop_branch_matches_torchvision():
    """models_r3m.py:85-90: frames that are not 3x224x224 go through transforms.Resize(256) + CenterCrop(224)."""
    tv_t = pytest.importorskip("torchvision.transforms")
    from r3m_b200.model import _resize256_center_crop224

    ref = tv_t.Compose([tv_t.Resize(256), tv_t.CenterCrop(224)])
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 3, 240, 320), (1, 3, 480, 300), (3, 3, 256, 256)):
        x = torch.rand(shape, generator=g) * 255
        got = _resize256_center_crop224(x)
        assert got.shape == (shape[0], 3, 224, 224)
        assert torch.allclose(got, ref(x), atol=1e-4)


def test_load_config_resolves_hydra_interpolations(tmp_path):
    """The checkpoints' config.yaml is written by hydra (r3m/cfgs/config_rep.yaml:30-41: ``lr: ${lr}``,
    ``bs: ${batch_size}``); without omegaconf the interpolations and YAML-1.1 floats like 1e-4 must still resolve."""
    path = tmp_path / "config.yaml"
    path.write_text("lr: 1e-4\nbatch_size: 32\ndevice: cuda\nouter:\n  hd: 512\nagent:\n  _target_: r3m.R3M\n"
                    "  device: ${device}\n  lr: ${lr}\n  hidden_dim: ${outer.hd}\n  size: 34\n  bs: ${batch_size}\n"
                    "  l2dist: true\n")
    cfg = r3m_b200.load_config(str(path))["agent"]
    assert cfg["lr"] == 1e-4 and isinstance(cfg["lr"], float)
    assert cfg["hidden_dim"] == 512 and cfg["bs"] == 32 and cfg["device"] == "cuda" and cfg["l2dist"] is True
    path.write_text("agent:\n  lr: ${missing}\n")
    with pytest.raises(ValueError):
        r3m_b200.load_config(str(path))
    path.write_text("agent:\n  lr: fast\n")
    with pytest.raises(ValueError):
        r3m_b200.load_config(str(path))


def test_data_parallel_replication_is_refused():
    """One GPU per process: the parameters are views into one flat block, so nn.DataParallel's replicate
    must fail loudly instead of sharing one engine between device threads."""
    from r3m_b200._lib import R3MB200Error

    m = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0)
    with pytest.raises(R3MB200Error):
        m._replicate_for_data_parallel()


def test_snapshot_is_a_superset_of_the_reference_format(tmp_path):
    """r3m/train_representation.py:123-138: {"r3m": DataParallel state_dict, "global_step"}; ours adds Adam + RNG."""
    import random

    import numpy as np

    m = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0)
    model = torch.nn.DataParallel(m)
    m._flat(2).normal_()
    m._flat(3).uniform_()
    m.encoder_opt.steps = 11
    torch.manual_seed(5)
    random.seed(6)
    np.random.seed(7)
    path = tmp_path / "snapshot.pt"
    assert r3m_b200.save_snapshot(str(path), model, 42, extra={"note": "x"})
    want = (torch.rand(3), random.random(), np.random.rand())
    payload = torch.load(path, weights_only=False)
    assert set(payload) == {"r3m", "global_step", "r3m_b200"} and payload["global_step"] == 42
    assert all(k.startswith("module.") for k in payload["r3m"])        # what the reference's load_snapshot feeds
    m2 = R3M("cpu", 3e-4, 1024, size=18, langweight=0.0)
    model2 = torch.nn.DataParallel(m2)
    step, extra = r3m_b200.load_snapshot(str(path), model2)
    assert step == 42 and extra == {"note": "x"} and m2.encoder_opt.steps == 11
    assert torch.equal(m2._flat(0), m._flat(0)) and torch.equal(m2._flat(2), m._flat(2))
    got = (torch.rand(3), random.random(), np.random.rand())
    assert torch.equal(got[0], want[0]) and got[1:] == want[1:]       # the random streams continue where they were
    # a snapshot written by the reference (no r3m_b200 entry) still loads: Adam restarts, as it does there
    torch.save({"r3m": model.state_dict(), "global_step": 7}, path)
    m3 = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0)
    assert r3m_b200.load_snapshot(str(path), torch.nn.DataParallel(m3)) == (7, None) and m3.encoder_opt.steps == 0
