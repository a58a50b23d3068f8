"""Parity at the BASELINE.json sizes and on the well-conditioned fixtures, against the oracle run ON the B200 in strict
fp32 (TF32 off) and against the goldens the real reference produced.

Why a second family of fixtures: at torchvision's default init a random ResNet-50's residual stream is dominated by a
per-channel constant, a 1x1 conv's output then has |mean| >> std, and storing it in bf16 ahead of the BatchNorm loses
the signal — the fp32 gradient moves by 25 % (RN18) ... 130 % (RN50) under ANY bf16 storage policy, so a gradient
comparison there cannot fail.  With the last BatchNorm weight of every block at 0.1
(torchvision's zero_init_residual idea, O.scale_last_gamma) and frames that differ from each other (O.varied_frames),
two fp32 implementations agree to 1e-4 and the bf16 policy costs 3-6 %: thresholds of 0.1 per layer group separate a
correct backward from a broken one (an all-zero gradient scores 1.0; dropping one skip-path accumulation > 0.5).

Tolerances (all relative L2, written where they are used): bf16-tier embeddings 2e-2 train / 5e-3 eval; loss heads on
identical embeddings 1e-4 (north star); gradients 0.1 per layer group and 0.2 per tensor."""
import json
import os

import numpy as np
import pytest
import torch

from fixtures import load_case, projections
from gpu_common import (HYPER, build_model, group_distances, oracle_update_on_gpu, rel, strict_fp32,
                        well_conditioned_state)
from oracle import r3m_oracle as O

pytestmark = pytest.mark.gpu
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _report(name, payload):
    if os.path.isdir(REPORT):
        path = os.path.join(REPORT, "r2_parity_report.json")
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = payload
        with open(path, "w") as f:
            json.dump(data, f, indent=1)


def _sentences(clips, lang):
    return ["" if i % 10 == 9 else "C does something %d" % i for i in range(clips)] if lang else [""] * clips


def _update_case(size, clips, lang, params, buffers, frames, perms, lang_emb, sentences, mask):
    """One Trainer.update of ours + the fp32 oracle on the GPU; returns everything the checks need."""
    from r3m_b200 import Trainer
    from test_engine_gpu import _check_loss_heads

    hyper = dict(HYPER, langweight=float(lang))
    m, model = build_model(size, params, buffers, float(lang), lang_emb)
    tr = Trainer(100)
    metrics, _ = tr.update(model, (frames.cuda(), sentences), 0, perms=perms, lang_emb=lang_emb)
    eng = m._any_engine()
    emb = eng.embeddings().clone()
    named = dict(m.named_parameters())
    ours = {k: v.grad.detach().clone() for k, v in named.items()}
    o_metrics, o_grads, o_emb, o_post, o_buf = oracle_update_on_gpu(size, params, buffers, frames, perms, hyper,
                                                                    lang_emb, mask)
    # loss heads on IDENTICAL embeddings: north-star 1e-4
    _check_loss_heads(eng, named, params, emb.cpu(), perms, hyper, lang_emb, mask, metrics, clips,
                      enumerate_kinks=clips <= 16)
    return m, metrics, emb, ours, o_metrics, o_grads, o_emb, o_buf


@pytest.mark.parametrize("size,clips,lang", [(50, 64, 1), (34, 128, 0), (18, 64, 1)],
                         ids=["c3_rn50_64clips_lang", "c4_rn34_128clips", "rn18_64clips_lang"])
def test_full_size_update_against_gpu_oracle(size, clips, lang):
    """BASELINE.json configs[2] (c3: ResNet-50, 64 clips x 5 frames, TCN + language + L1/L2) and configs[3] (c4:
    ResNet-34, 128 clips) on the well-conditioned state: embeddings, metrics and per-layer-group gradients against the
    fp32 oracle evaluated on the same GPU.  Tile tails, split-K and >2^31-element offsets only show at these sizes."""
    params, buffers = well_conditioned_state(size, 21, bool(lang))
    frames = O.varied_frames(clips, 22)
    perms = O.draw_permutations(clips, 23)
    lang_emb = O.stub_lang_embedding(clips, 24) if lang else None
    sentences = _sentences(clips, lang)
    mask = torch.tensor([1.0 * (s != "") for s in sentences]) if lang else None
    m, metrics, emb, ours, o_metrics, o_grads, o_emb, o_buf = _update_case(size, clips, lang, params, buffers, frames,
                                                                           perms, lang_emb, sentences, mask)
    d_emb = rel(emb, o_emb)
    assert d_emb < 2e-2, d_emb  # bf16 tier, train-mode BN (measured 5e-3)
    for k in ("l2loss", "l1loss", "tcnloss", "full_loss") + (("rewloss",) if lang else ()):
        assert abs(metrics[k] - o_metrics[k]) <= 2e-2 * abs(o_metrics[k]) + 1e-6, (k, metrics[k], o_metrics[k])
    dist = group_distances(ours, o_grads)
    for g, d in dist.items():
        assert d < 0.1, (g, d, dist)
    if lang:
        d_lang = group_distances(ours, o_grads, ("lang_rew",))["lang_rew"]
        assert d_lang < 0.3, d_lang  # driven by the embeddings' bf16 noise through a near-cancelling InfoNCE gradient
        dist["lang_rew"] = d_lang
    # every tensor on its own: a single mis-wired layer cannot hide in a group norm
    worst = ("", 0.0)
    for k, g in o_grads.items():
        if not k.startswith("convnet.") or float(g.norm()) < 1e-5 * max(float(v.norm()) for v in o_grads.values()):
            continue
        d = rel(ours[k], g)
        if d > worst[1]:
            worst = (k, d)
        assert d < 0.25, (k, d)
    # BatchNorm running statistics after the step (momentum 0.1, unbiased variance)
    sd = m.state_dict()
    for k in ("convnet.bn1.running_mean", "convnet.bn1.running_var", "convnet.layer3.1.bn2.running_var",
              "convnet.layer4.0.downsample.1.running_mean"):
        assert rel(sd[k], o_buf[k]) < 1e-2, k
    _report(f"full_size_rn{size}_{clips}clips", {"embeddings": d_emb, "grad_groups": dist, "worst_tensor": worst,
                                                 "metrics": {k: (metrics[k], o_metrics[k]) for k in metrics}})


@pytest.mark.parametrize("name", ["rn18_wc", "rn34_wc", "rn50_wc"])
def test_well_conditioned_goldens_of_the_real_reference(name):
    """The committed outputs of facebookresearch/r3m's own Trainer.update on the well-conditioned fixtures: embeddings,
    metrics, every BatchNorm gradient in full and 8 random projections of every filter gradient."""
    z, case, gold, params, buffers, frames, perms, lang_emb, sentences, mask = load_case(name)
    lang = case["langweight"] > 0
    m, metrics, emb, ours, o_metrics, o_grads, o_emb, _ = _update_case(case["size"], case["clips"], lang, params,
                                                                       buffers, frames, perms, lang_emb, sentences, mask)
    assert set(metrics) == set(gold)
    assert rel(o_emb, z["embeddings"]) < 1e-4          # the GPU fp32 oracle reproduces the reference's CPU run
    assert rel(emb, z["embeddings"]) < 2e-2
    for k in ("l2loss", "l1loss", "tcnloss", "full_loss"):
        assert abs(metrics[k] - gold[k]) <= 2e-2 * abs(gold[k]), (k, metrics[k], gold[k])
    names = json.loads(bytes(z["grad_names_json"]).decode())
    proj = projections(ours, names)
    scale = {k: float(z["grad_norms"][i]) for i, k in enumerate(names)}
    gmax = float(z["grad_norms"].max())
    report = {}
    for k in names:
        if not k.startswith("convnet.") or scale[k] < 1e-5 * gmax:
            continue
        if "gproj::" + k in z.files:
            d = float((proj[k] - torch.as_tensor(z["gproj::" + k])).norm()) / (8 ** 0.5 * scale[k])
        else:
            d = rel(ours[k], z["grad::" + k])
        report[k] = d
        assert d < 0.3, (k, d)  # 8 projections estimate the relative error to +-35 %
    dist = group_distances(ours, o_grads)
    for g, d in dist.items():
        assert d < 0.1, (g, d, dist)
    _report("golden_" + name, {"grad_groups": dist, "worst_tensor": max(report.items(), key=lambda kv: kv[1]),
                               "embeddings": rel(emb, z["embeddings"])})


@pytest.mark.parametrize("mode", ["eval", "train_wc", "train_default"])
def test_c2_forward_batch256_against_gpu_oracle(mode):
    """BASELINE.json configs[1]: ResNet-50 encoder forward, batch 256, both BatchNorm modes."""
    from r3m_b200 import R3M

    strict_fp32()
    if mode == "eval":
        params, buffers = O.eval_fixture_state(50)
    elif mode == "train_wc":
        params, buffers = well_conditioned_state(50, 31, False)
    else:
        params, buffers = O.init_state(50, 31)
    frames = O.varied_frames(52, 32).reshape(-1, 3, 224, 224)[:256]
    m = R3M("cuda", 1e-4, 1024, size=50, langweight=0.0)
    sd = dict(params)
    sd.update(buffers)
    m.load_state_dict(sd)
    m = m.cuda()
    train = mode != "eval"
    m.train(train)
    with torch.no_grad():
        out = m(frames.cuda())
    dev = torch.device("cuda")
    p = {k: v.to(dev) for k, v in params.items()}
    res = {}
    for pol in ("fp32", "bf16"):
        b = {k: v.to(dev) for k, v in buffers.items()}
        with torch.no_grad():
            res[pol] = O.r3m_forward(p, b, frames.to(dev), 50, train, policy=pol)
        if pol == "fp32":
            post = b
    d, d_pol = rel(out, res["fp32"]), rel(res["bf16"], res["fp32"])
    assert out.shape == (256, 2048) and out.dtype == torch.float32
    if mode == "eval":
        assert d < 5e-3, d
        assert torch.equal(m.state_dict()["convnet.bn1.running_mean"].cpu(), buffers["convnet.bn1.running_mean"])
    elif mode == "train_wc":
        assert d < 1.5e-2, d
    else:
        assert d < 1.5 * d_pol + 1e-3 and d < 0.2, (d, d_pol)  # default init: the bf16 policy itself is ~0.1 away
    if train:
        sd2 = m.state_dict()
        assert rel(sd2["convnet.layer2.0.bn1.running_var"], post["convnet.layer2.0.bn1.running_var"]) < 2e-2
        assert int(sd2["convnet.bn1.num_batches_tracked"]) == 1
    _report("c2_forward_" + mode, {"ours_vs_fp32": d, "bf16_policy_vs_fp32": d_pol})


@pytest.mark.parametrize("size", [18, 50])
def test_tf32_tier_eval_embeddings_within_1e3_of_the_reference(size):
    """North-star tolerance (embeddings <= 1e-3 relative) on the inference path: the tf32 tier (fp32 storage,
    kind::tf32 tensor cores, round-to-nearest operands) against the goldens the REAL reference produced
    (tests/golden/rn{18,50}_eval_b4.npz; reference semantics r3m/models/models_r3m.py:97-99: fp32 in, fp32 conv)."""
    from r3m_b200 import R3M

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"rn{size}_eval_b4.npz"))
    params, buffers = O.eval_fixture_state(size)
    frames = O.synthetic_frames(1, 7)[0, :4]
    m = R3M("cuda", 1e-4, 1024, size=size, langweight=0.0)
    sd = dict(params)
    sd.update(buffers)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with torch.no_grad():
        bf16 = m(frames.cuda())
        m.set_eval_precision("tf32")
        tf32 = m(frames.cuda())
        tf32_u8 = m(frames.to(torch.uint8).cuda())
        again = m(frames.cuda())
    d_bf16, d_tf32 = rel(bf16, z["embeddings"]), rel(tf32, z["embeddings"])
    assert d_tf32 < 1e-3, (d_tf32, d_bf16)
    assert d_bf16 < 5e-3
    assert torch.equal(tf32, again) and torch.equal(tf32, tf32_u8)  # deterministic; uint8 frames identical
    _report(f"tf32_eval_rn{size}_b4", {"tf32_vs_reference": d_tf32, "bf16_vs_reference": d_bf16})


def test_tf32_tier_at_c2_size_and_back_to_bf16():
    """BASELINE configs[1] size (ResNet-50, batch 256, eval BatchNorm) in the tf32 tier against the fp32 oracle on the
    GPU; switching tiers back and forth leaves the bf16 results and a following training step intact."""
    from r3m_b200 import R3M, Trainer

    strict_fp32()
    params, buffers = O.eval_fixture_state(50)
    frames = O.varied_frames(52, 33).reshape(-1, 3, 224, 224)[:256].cuda()
    m = R3M("cuda", 1e-4, 1024, size=50, langweight=0.0)
    sd = dict(params)
    sd.update(buffers)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    dev = torch.device("cuda")
    with torch.no_grad():
        want = O.r3m_forward({k: v.to(dev) for k, v in params.items()}, {k: v.to(dev) for k, v in buffers.items()},
                             frames, 50, False)
        a = m(frames)
        b = m.set_eval_precision("tf32")(frames)
        c = m.set_eval_precision("bf16")(frames)
    assert rel(b, want) < 1e-3, rel(b, want)
    assert rel(a, want) < 5e-3 and torch.equal(a, c)
    _report("tf32_c2_forward_eval", {"tf32_vs_fp32": rel(b, want), "bf16_vs_fp32": rel(a, want)})
    # a training step on the same model afterwards (the tf32 tier aliases the activation arena)
    m.set_eval_precision("tf32")
    model = torch.nn.DataParallel(m, device_ids=[torch.cuda.current_device()])
    clips = 4
    batch = O.varied_frames(clips, 34).cuda()
    perms = O.draw_permutations(clips, 35)
    tr = Trainer(100)
    m1, _ = tr.update(model, (batch, [""] * clips), 0, perms=perms)
    with torch.no_grad():
        m.eval()
        m(frames[:20])
    m2, _ = tr.update(model, (batch, [""] * clips), 1, perms=perms)
    assert all(np.isfinite(list(x.values())).all() for x in (m1, m2))
