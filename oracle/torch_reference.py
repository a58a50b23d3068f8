"""The reference pipeline as torch modules, for the SAME-BOX GPU reference arm — TEST / BENCH INFRASTRUCTURE ONLY.

``/root/reference`` does not travel to the GPU box, but what it executes there is torchvision + torch (cuDNN):
``torchvision.models.resnet{18,34,50}`` with ``fc = Identity`` behind ``x/255 -> Normalize`` (r3m/models/
models_r3m.py:44-62,84-100), the ``LanguageReward`` MLP (r3m/models/models_language.py:37-55), ``torch.optim.Adam``
(models_r3m.py:76) and the update step of r3m/trainer.py:25-162 with its ``.item()`` read-backs.  This module
restates that pipeline on the stock torch / torchvision modules so that bench.py (``gpu_reference``),
tools/gpu_reference.py and the full-size GPU parity tests can run "the reference on this B200 through cuDNN":

  * ``variant="as_written"``  fp32 API exactly like the reference (cuDNN convs may use TF32 — torch's default —
    matmuls fp32), NCHW, ``cudnn.benchmark = True`` (train_representation.py:24);
  * ``variant="bf16_channels_last"``  the strongest fair form: bf16 autocast + channels_last + cudnn.benchmark;
  * ``variant="fp32_strict"``  TF32 off everywhere: the fp32 ground truth for parity at sizes the CPU cannot do.

Nothing under r3m_b200/ imports this module.
"""
import torch
import torch.nn as nn
import torchvision

from . import r3m_oracle as O

EPSILON = O.EPSILON


class _LanguageReward(nn.Module):
    """models_language.py:43-55: Linear(2D+L, H) ReLU (Linear(H, H) ReLU) x3 Linear(H, 1) on cat([e0, eg, le])."""

    def __init__(self, outdim, hidden_dim, lang_dim=768):
        super().__init__()
        self.pred = nn.Sequential(nn.Linear(2 * outdim + lang_dim, hidden_dim), nn.ReLU(inplace=True),
                                  nn.Linear(hidden_dim, hidden_dim), nn.ReLU(inplace=True),
                                  nn.Linear(hidden_dim, hidden_dim), nn.ReLU(inplace=True),
                                  nn.Linear(hidden_dim, hidden_dim), nn.ReLU(inplace=True),
                                  nn.Linear(hidden_dim, 1))

    def forward(self, e0, eg, le):
        return self.pred(torch.cat([e0, eg, le], -1)).squeeze(), {}


class TorchR3M(nn.Module):
    """The reference module graph (torchvision ResNet + Normalize + LanguageReward + Adam)."""

    def __init__(self, size, hidden_dim=1024, lr=1e-4, l2weight=1e-5, l1weight=1e-5, langweight=1.0, tcnweight=1.0,
                 l2dist=True, lang_embedding=None):
        super().__init__()
        self.size, self.l2weight, self.l1weight, self.langweight, self.tcnweight = size, l2weight, l1weight, langweight, tcnweight
        self.l2dist, self.num_negatives = l2dist, 3
        self.outdim = O.OUTDIM[size]
        self.convnet = getattr(torchvision.models, f"resnet{size}")(weights=None)
        self.convnet.fc = nn.Identity()
        self.register_buffer("_mean", torch.tensor(O.MEAN)[None, :, None, None], persistent=False)
        self.register_buffer("_std", torch.tensor(O.STD)[None, :, None, None], persistent=False)
        params = list(self.convnet.parameters())
        self._lang_embedding = lang_embedding  # stands in for the frozen DistilBERT output
        if langweight > 0:
            self.lang_rew = _LanguageReward(self.outdim, hidden_dim)
            params += list(self.lang_rew.parameters())
        self.encoder_opt = torch.optim.Adam(params, lr=lr)
        self.channels_last = False

    def load_oracle_state(self, params, buffers):
        sd = {k: v for k, v in list(params.items()) + list(buffers.items())}
        missing = self.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys, missing.unexpected_keys

    def forward(self, obs):
        x = obs.float() / 255.0
        x = (x - self._mean) / self._std
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        return self.convnet(x)

    def sim(self, a, b):
        if self.l2dist:
            return -torch.linalg.norm(a - b, dim=-1)
        return torch.nn.functional.cosine_similarity(a, b, dim=1)

    def get_reward(self, e0, es, sentences):
        le = self._lang_embedding  # the reference re-encodes the sentences here, 15 times per step
        return self.lang_rew(e0, es, le.to(e0.dtype))


def update(model, batch, perms=None, eval_mode=False, autocast_bf16=False, sync_metrics=True):
    """r3m/trainer.py:25-162 on ``TorchR3M``.  ``perms`` ([15, B], reference draw order) replaces the randperm draws
    when given.  ``sync_metrics`` keeps the reference's per-metric ``.item()`` read-backs."""
    b_im, b_lang = batch
    model.train(not eval_mode)
    bs = b_im.shape[0]
    draw = iter(perms) if perms is not None else None

    def randperm():
        if draw is not None:
            return next(draw).to(b_im.device)
        return torch.randperm(bs, device=b_im.device)

    item = (lambda t: t.item()) if sync_metrics else (lambda t: t.detach())
    metrics = {}
    with torch.set_grad_enabled(not eval_mode), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast_bf16):
        alles = model(b_im.reshape(bs * 5, 3, 224, 224))
    alles = alles.float()
    with torch.set_grad_enabled(not eval_mode):
        alle = alles.reshape(bs, 5, -1)
        e0, eg, es0, es1, es2 = (alle[:, i] for i in range(5))
        l2loss = torch.linalg.norm(alles, ord=2, dim=-1).mean()
        l1loss = torch.linalg.norm(alles, ord=1, dim=-1).mean()
        l0loss = torch.linalg.norm(alles, ord=0, dim=-1).mean()
        metrics["l2loss"], metrics["l1loss"], metrics["l0loss"] = item(l2loss), item(l1loss), item(l0loss)
        full = model.l2weight * l2loss + model.l1weight * l1loss
        if model.langweight > 0:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast_bf16):
                G = lambda a, b: model.get_reward(a, b, b_lang)[0].float()  # noqa: E731
                pos = [G(e0, eg), G(e0, es1), G(e0, es2)]
                negs = [[G(e0, e0)], [G(e0, es0)], [G(e0, es1)]]
                targets = [eg, es1, es2]
                for _ in range(model.num_negatives):
                    for t in range(3):
                        idx = randperm()
                        negs[t].append(G(e0[idx], targets[t][idx]))
            rew, accs = 0, []
            for t in range(3):
                n = torch.stack(negs[t], -1)
                rew = rew + -torch.log(EPSILON + torch.exp(pos[t]) / (EPSILON + torch.exp(pos[t]) + torch.exp(n).sum(-1)))
                accs.append((1.0 * (n.max(-1)[0] < pos[t])).mean())
            mask = torch.tensor([1.0 * (b != "") for b in b_lang], dtype=torch.float32).to(b_im.device)
            rewloss = ((rew / 3) * mask).mean()
            metrics["rewloss"] = item(rewloss)
            for t in range(3):
                metrics[f"rewacc{t + 1}"] = item(accs[t])
            full = full + model.langweight * rewloss
        elif draw is not None:
            for _ in range(9):
                next(draw)
        if model.tcnweight > 0:
            s02, s12, s01 = model.sim(es2, es0), model.sim(es2, es1), model.sim(es1, es0)
            neg0, neg2 = [], []
            for _ in range(model.num_negatives):
                neg0.append(model.sim(es0, es0[randperm()]))
                neg2.append(model.sim(es2, es2[randperm()]))
            neg0, neg2 = torch.stack(neg0, -1), torch.stack(neg2, -1)
            sm1 = -torch.log(EPSILON + torch.exp(s12) / (EPSILON + torch.exp(s02) + torch.exp(s12) + torch.exp(neg2).sum(-1)))
            sm2 = -torch.log(EPSILON + torch.exp(s01) / (EPSILON + torch.exp(s01) + torch.exp(s02) + torch.exp(neg0).sum(-1)))
            tcn = ((sm1 + sm2) / 2.0).mean()
            aligned = ((1.0 * (s02 < s12)) * (1.0 * (s01 > s02))).mean()
            metrics["tcnloss"], metrics["aligned"] = item(tcn), item(aligned)
            full = full + model.tcnweight * tcn
        metrics["full_loss"] = item(full)
        if not eval_mode:
            model.encoder_opt.zero_grad()
            full.backward()
            model.encoder_opt.step()
    return metrics, alles.detach()


def configure(variant):
    """Global torch flags of a variant; returns (autocast_bf16, channels_last)."""
    torch.backends.cudnn.benchmark = True  # train_representation.py:24
    if variant == "as_written":
        torch.backends.cudnn.allow_tf32 = True  # torch default: fp32 convs may run on TF32 tensor cores
        torch.backends.cuda.matmul.allow_tf32 = False
        return False, False
    if variant == "bf16_channels_last":
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        return True, True
    if variant == "fp32_strict":
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        return False, False
    raise ValueError(variant)


def build(size, variant, device, lang=True, seed=0, lang_embedding=None, hyper=None):
    hyper = hyper or dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, lr=1e-4, hidden_dim=1024)
    autocast, cl = configure(variant)
    torch.manual_seed(seed)
    m = TorchR3M(size, hidden_dim=hyper.get("hidden_dim", 1024), lr=hyper["lr"], l2weight=hyper["l2weight"],
                 l1weight=hyper["l1weight"], langweight=1.0 if lang else 0.0, tcnweight=hyper["tcnweight"],
                 lang_embedding=lang_embedding)
    m = m.to(device)
    if cl:
        m = m.to(memory_format=torch.channels_last)
        m.channels_last = True
    return m, autocast


def time_update(size, clips, variant, device, lang=True, steps=10, warmup=5):
    """Device-timed frames/s of the reference step at (size, clips) through torch/cuDNN."""
    g = torch.Generator(device=device).manual_seed(1)
    frames = torch.randint(0, 255, (clips, 5, 3, 224, 224), generator=g, device=device).float()
    emb = torch.randn(clips, 768, generator=torch.Generator().manual_seed(1234)).to(device) if lang else None
    model, autocast = build(size, variant, device, lang=lang, lang_embedding=emb)
    sentences = ["" if i % 10 == 9 else "C does something %d" % i for i in range(clips)]
    for _ in range(warmup):
        update(model, (frames, sentences), autocast_bf16=autocast)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        update(model, (frames, sentences), autocast_bf16=autocast)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model
    torch.cuda.empty_cache()
    return {"variant": variant, "frames_per_s": clips * 5 / (ms * 1e-3), "ms_per_step": ms, "clips": clips,
            "size": size, "steps": steps, "warmup": warmup}


def time_forward(size, batch, variant, device, train_bn, steps=20, warmup=5):
    """Device-timed frames/s of R3M.forward (c2) through torch/cuDNN."""
    g = torch.Generator(device=device).manual_seed(1)
    frames = torch.randint(0, 255, (batch, 3, 224, 224), generator=g, device=device).float()
    model, autocast = build(size, variant, device, lang=False)
    model.train(train_bn)

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return model(frames)

    for _ in range(warmup):
        fwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fwd()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model
    torch.cuda.empty_cache()
    return {"variant": variant, "frames_per_s": batch / (ms * 1e-3), "ms_per_call": ms, "batch": batch, "size": size,
            "train_bn": bool(train_bn)}
