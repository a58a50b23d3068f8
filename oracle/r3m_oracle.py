"""CPU oracle for the R3M pretraining hot path — TEST INFRASTRUCTURE ONLY.

This is a plain-PyTorch (CPU, fp32, torch.nn.functional) restatement of what the reference computes on the hot path:
``R3M.forward`` (r3m/models/models_r3m.py:84-100), the torchvision ResNet-18/34/50 it wraps
(torchvision/models/resnet.py:89-105 BasicBlock.forward, :143-163 Bottleneck.forward, :266-282 _forward_impl; the
reference pins torchvision==0.8.2 / torch==1.7.1 in r3m/r3m_base.yaml:60-61), ``LanguageReward``
(r3m/models/models_language.py:37-55), ``R3M.sim`` (models_r3m.py:102-107) and ``Trainer.update``
(r3m/trainer.py:25-162) including the Adam step (torch.optim.Adam defaults, models_r3m.py:76).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker / CPU baseline.  Nothing under r3m_b200/ imports it.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned against the
reference ITSELF, imported from /root/reference in the build container by oracle/make_golden.py, which checks
oracle == reference (embeddings, every loss term / metric, every gradient, post-Adam weights) to fp32 round-off and
commits the small fixtures under tests/golden/ that the tests replay.

Differences from the reference that are deliberate and documented:
  * permutations (``torch.randperm`` drawn from the global CPU generator, trainer.py:87-91,136-137) and the language
    embedding (DistilBERT, models_language.py:23-35, un-runnable offline) are INPUTS here, so both sides consume
    identical values;
  * the language mask is built on the tensor's device instead of the hard-coded ``.cuda()`` (trainer.py:108).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

EPSILON = 1e-8  # r3m/trainer.py:18
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
MEAN = (0.485, 0.456, 0.406)  # models_r3m.py:61
STD = (0.229, 0.224, 0.225)

_CFG = {
    18: ("basic", (2, 2, 2, 2)),      # tv resnet.py:705
    34: ("basic", (3, 4, 6, 3)),      # tv resnet.py:731
    50: ("bottleneck", (3, 4, 6, 3)),  # tv resnet.py:763
}
OUTDIM = {18: 512, 34: 512, 50: 2048}  # models_r3m.py:45,48,51


# ---------------------------------------------------------------------------------------------------------------
# architecture description shared by init / forward / the state-dict bridge
# ---------------------------------------------------------------------------------------------------------------
def conv_specs(size):
    """List of (name, Cin, Cout, k, stride, pad) for every conv, in torchvision parameter order, with the name of
    the BatchNorm that follows it."""
    kind, layers = _CFG[size]
    specs = [("conv1", "bn1", 3, 64, 7, 2, 3)]
    inplanes = 64
    expansion = 1 if kind == "basic" else 4
    for li, (planes, nblocks) in enumerate(zip((64, 128, 256, 512), layers)):
        for b in range(nblocks):
            stride = 2 if (li > 0 and b == 0) else 1
            pre = f"layer{li + 1}.{b}"
            if kind == "basic":
                specs.append((f"{pre}.conv1", f"{pre}.bn1", inplanes, planes, 3, stride, 1))
                specs.append((f"{pre}.conv2", f"{pre}.bn2", planes, planes, 3, 1, 1))
            else:
                specs.append((f"{pre}.conv1", f"{pre}.bn1", inplanes, planes, 1, 1, 0))
                specs.append((f"{pre}.conv2", f"{pre}.bn2", planes, planes, 3, stride, 1))  # v1.5: stride on 3x3
                specs.append((f"{pre}.conv3", f"{pre}.bn3", planes, planes * expansion, 1, 1, 0))
            if b == 0 and (stride != 1 or inplanes != planes * expansion):
                specs.append((f"{pre}.downsample.0", f"{pre}.downsample.1", inplanes, planes * expansion, 1, stride, 0))
            inplanes = planes * expansion
    return specs


def init_state(size, seed, hidden_dim=1024, lang=False, lang_dim=768):
    """Seeded random initial state with torchvision's init LAWS (kaiming_normal_(fan_out, relu) convs, BN gamma=1
    beta=0 — tv resnet.py:208-213; nn.Linear default init for the language head).  Returns (params, buffers) as
    OrderedDicts keyed like ``R3M.state_dict()`` without the DataParallel ``module.`` prefix."""
    g = torch.Generator().manual_seed(seed)
    params, buffers = OrderedDict(), OrderedDict()
    for conv, bn, cin, cout, k, _s, _p in conv_specs(size):
        std = math.sqrt(2.0 / (cout * k * k))
        params[f"convnet.{conv}.weight"] = torch.randn(cout, cin, k, k, generator=g) * std
        params[f"convnet.{bn}.weight"] = torch.ones(cout)
        params[f"convnet.{bn}.bias"] = torch.zeros(cout)
        buffers[f"convnet.{bn}.running_mean"] = torch.zeros(cout)
        buffers[f"convnet.{bn}.running_var"] = torch.ones(cout)
        buffers[f"convnet.{bn}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    if lang:
        d = OUTDIM[size]
        dims = [2 * d + lang_dim, hidden_dim, hidden_dim, hidden_dim, hidden_dim, 1]
        for i in range(5):
            fan_in = dims[i]
            bound = 1.0 / math.sqrt(fan_in)
            params[f"lang_rew.pred.{2 * i}.weight"] = (torch.rand(dims[i + 1], fan_in, generator=g) * 2 - 1) * bound
            params[f"lang_rew.pred.{2 * i}.bias"] = (torch.rand(dims[i + 1], generator=g) * 2 - 1) * bound
    return params, buffers


def eval_fixture_state(size):
    """The state of the eval-forward goldens (tests/golden/rn{18,50}_eval_b4.npz): seeded init with NON-trivial running
    statistics, so that eval-mode BatchNorm is exercised."""
    params, buffers = init_state(size, 5)
    g = torch.Generator().manual_seed(6)
    for k in buffers:
        if k.endswith("running_mean"):
            buffers[k] = 0.1 * torch.randn(buffers[k].shape, generator=g)
        elif k.endswith("running_var"):
            buffers[k] = 0.5 + torch.rand(buffers[k].shape, generator=g)
    return params, buffers


def synthetic_frames(num_clips, seed):
    """[B,5,3,224,224] float frames with integer values in [0,255) (mirrors r3m/example.py:29 randint(0,255))."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 255, (num_clips, 5, 3, 224, 224), generator=g).float()


def structured_frames(num_clips, seed):
    """Smoother frames (bilinear-upsampled 14x14 noise + 10 % pixel noise, rounded) — closer to video statistics."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(num_clips * 5, 3, 14, 14, generator=g)
    up = F.interpolate(low, size=(224, 224), mode="bilinear", align_corners=False)
    x = 255.0 * (0.9 * up + 0.1 * torch.rand(num_clips * 5, 3, 224, 224, generator=g))
    return x.round().clamp(0, 255).reshape(num_clips, 5, 3, 224, 224)


def varied_frames(num_clips, seed):
    """``structured_frames`` with a per-frame, per-colour gain in [0.1, 1.5): frames (hence embeddings) differ from each
    other by much more than the bf16 storage noise — half of what makes the "well-conditioned" parity fixtures."""
    g = torch.Generator().manual_seed(seed + 1000)
    gain = torch.rand(num_clips, 5, 3, 1, 1, generator=g) * 1.4 + 0.1
    return (structured_frames(num_clips, seed) * gain).clamp(0, 255).round()


def scale_last_gamma(params, size, value):
    """Sets the weight of the LAST BatchNorm of every residual block (bn3 of a Bottleneck, bn2 of a BasicBlock) to
    ``value`` — torchvision's ``zero_init_residual`` option (tv resnet.py:214-220) with a small non-zero value.  At the
    default init (1.0) the residual stream of a random-init ResNet-50 is dominated by a per-channel constant, a
    1x1 conv's output then has |mean| >> std, and storing it in bf16 before the BatchNorm loses the signal: the fp32
    gradient moves by >100 % under ANY bf16 storage policy (DESIGN.md "noise floor").  With 0.1 every conv sees
    relu(N(0,1))-like inputs and the same comparison is conditioned to a few per cent."""
    tail = "bn3.weight" if _CFG[size][0] == "bottleneck" else "bn2.weight"
    for k in params:
        if k.startswith("convnet.layer") and k.endswith(tail):
            params[k] = torch.full_like(params[k], value)
    return params


def draw_permutations(batch, seed):
    """15 permutations in the reference's draw order: 9 for the language branch (trainer.py:86-92: n1,n2,n3 per
    iteration), then 6 for TCN (trainer.py:135-137: es0 then es2 per iteration)."""
    g = torch.Generator().manual_seed(seed)
    return torch.stack([torch.randperm(batch, generator=g) for _ in range(15)])


def stub_lang_embedding(batch, seed, dim=768):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, dim, generator=g)


# ---------------------------------------------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------------------------------------------
class _RoundBf16(torch.autograd.Function):
    """Storage-precision model of the sm_100a path: a tensor that the CUDA path keeps in HBM as bf16 is rounded to
    bf16 here (value AND incoming gradient), all arithmetic stays fp32 — "the reference under the same dtype policy"
    (SURVEY.md §7 hard part 1(c))."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundBf16Forward(torch.autograd.Function):
    """Filters: the CUDA path reads a bf16 copy of the fp32 master filter but accumulates the filter gradient in fp32,
    so only the VALUE is rounded (straight-through gradient)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _policy_round(policy):
    if policy == "fp32":
        return lambda t: t
    if policy == "bf16":
        return _RoundBf16.apply
    raise ValueError(policy)


def _policy_round_weight(policy):
    return (lambda t: t) if policy == "fp32" else _RoundBf16Forward.apply


def _bn(x, params, buffers, name, train):
    """nn.BatchNorm2d forward; in train mode also the running-stat update (momentum 0.1, unbiased variance)."""
    w, b = params[f"{name}.weight"], params[f"{name}.bias"]
    rm, rv = buffers[f"{name}.running_mean"], buffers[f"{name}.running_var"]
    if train:
        n = x.numel() // x.shape[1]
        mean = x.mean((0, 2, 3))
        var = x.var((0, 2, 3), unbiased=False)
        with torch.no_grad():
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * n / max(n - 1, 1))
            buffers[f"{name}.num_batches_tracked"] += 1
    else:
        mean, var = rm, rv
    xhat = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + BN_EPS)
    return xhat * w[None, :, None, None] + b[None, :, None, None]


def resnet_forward(params, buffers, x, size, train, taps=None, policy="fp32"):
    """tv resnet.py:266-282 with fc = Identity (models_r3m.py:62).  ``taps`` (optional dict) collects intermediate
    activations keyed by layer name for per-layer parity.  ``policy="bf16"`` rounds exactly the tensors the CUDA path
    stores as bf16 (network input, filters, raw conv outputs, post-ReLU activations)."""
    kind, layers = _CFG[size]
    P = lambda k: params["convnet." + k]  # noqa: E731
    q = _policy_round(policy)
    qw = _policy_round_weight(policy)
    x = q(x)

    def conv(t, name, stride, pad):
        return q(F.conv2d(t, qw(P(name + ".weight")), None, stride, pad))

    def bn(t, name):
        return _bn(t, params, buffers, "convnet." + name, train)

    x = conv(x, "conv1", 2, 3)
    if taps is not None:
        taps["conv1.raw"] = x
    x = q(F.relu(bn(x, "bn1")))
    x = F.max_pool2d(x, 3, 2, 1)
    if taps is not None:
        taps["maxpool"] = x
    inplanes = 64
    expansion = 1 if kind == "basic" else 4
    for li, (planes, nblocks) in enumerate(zip((64, 128, 256, 512), layers)):
        for b in range(nblocks):
            stride = 2 if (li > 0 and b == 0) else 1
            pre = f"layer{li + 1}.{b}"
            identity = x
            if kind == "basic":
                out = q(F.relu(bn(conv(x, f"{pre}.conv1", stride, 1), f"{pre}.bn1")))
                out = bn(conv(out, f"{pre}.conv2", 1, 1), f"{pre}.bn2")
            else:
                out = q(F.relu(bn(conv(x, f"{pre}.conv1", 1, 0), f"{pre}.bn1")))
                out = q(F.relu(bn(conv(out, f"{pre}.conv2", stride, 1), f"{pre}.bn2")))
                out = bn(conv(out, f"{pre}.conv3", 1, 0), f"{pre}.bn3")
            if b == 0 and (stride != 1 or inplanes != planes * expansion):
                # the CUDA path never materialises the downsample BatchNorm's output (it is folded into the block tail)
                identity = bn(conv(x, f"{pre}.downsample.0", stride, 0), f"{pre}.downsample.1")
            x = q(F.relu(out + identity))
            if taps is not None:
                taps[pre] = x
            inplanes = planes * expansion
    return x.mean((2, 3))  # AdaptiveAvgPool2d((1,1)) + flatten


def r3m_forward(params, buffers, obs, size, train, taps=None, policy="fp32"):
    """models_r3m.py:84-100 for obs_shape == [3,224,224]: obs.float()/255 -> Normalize -> convnet."""
    x = obs.float() / 255.0
    mean = torch.tensor(MEAN, dtype=x.dtype, device=x.device)[None, :, None, None]
    std = torch.tensor(STD, dtype=x.dtype, device=x.device)[None, :, None, None]
    x = (x - mean) / std
    return resnet_forward(params, buffers, x, size, train, taps, policy)


def sim(a, b, l2dist=True):
    """models_r3m.py:102-107: negative L2 distance (l2dist=True, the default) or nn.CosineSimilarity(dim=1)."""
    if l2dist:
        return -torch.linalg.norm(a - b, dim=-1)
    return F.cosine_similarity(a, b, dim=1, eps=1e-08)


def lang_reward(params, e0, eg, le, taps=None):
    """models_language.py:43-55: 5-layer MLP on cat([e0, eg, le]).  `taps` (a list) collects (layer, pre-activation)
    of the four hidden layers — used by the tests to find ReLU inputs that sit on the kink to within round-off."""
    h = torch.cat([e0, eg, le], -1)
    for i in range(5):
        h = F.linear(h, params[f"lang_rew.pred.{2 * i}.weight"], params[f"lang_rew.pred.{2 * i}.bias"])
        if i < 4:
            if taps is not None:
                taps.append((i, h))
            h = F.relu(h)
    return h.squeeze(-1)


def losses(params, alles, perms, hyper, lang_emb=None, lang_mask=None, lang_taps=None):
    """trainer.py:42-152 given the embeddings.  Returns (full_loss tensor, metrics dict of python floats)."""
    bs = alles.shape[0] // 5
    alle = alles.reshape(bs, 5, -1)
    e0, eg, es0, es1, es2 = (alle[:, i] for i in range(5))
    metrics = OrderedDict()
    l2loss = torch.linalg.norm(alles, ord=2, dim=-1).mean()
    l1loss = torch.linalg.norm(alles, ord=1, dim=-1).mean()
    l0loss = torch.linalg.norm(alles, ord=0, dim=-1).mean()
    metrics["l2loss"], metrics["l1loss"], metrics["l0loss"] = l2loss.item(), l1loss.item(), l0loss.item()
    full = hyper["l2weight"] * l2loss + hyper["l1weight"] * l1loss
    pi = 0
    if hyper["langweight"] > 0:
        G = lambda a, b: lang_reward(params, a, b, lang_emb, lang_taps)  # noqa: E731
        pos = [G(e0, eg), G(e0, es1), G(e0, es2)]
        negs = [[G(e0, e0)], [G(e0, es0)], [G(e0, es1)]]
        targets = [eg, es1, es2]
        for _ in range(3):
            for t in range(3):
                idx = perms[pi]
                pi += 1
                negs[t].append(G(e0[idx], targets[t][idx]))
        rew = 0
        accs = []
        for t in range(3):
            n = torch.stack(negs[t], -1)
            rew = rew + -torch.log(EPSILON + torch.exp(pos[t]) / (EPSILON + torch.exp(pos[t]) + torch.exp(n).sum(-1)))
            accs.append((1.0 * (n.max(-1)[0] < pos[t])).mean())
        rewloss = ((rew / 3) * lang_mask).mean()
        metrics["rewloss"] = rewloss.item()
        for t in range(3):
            metrics[f"rewacc{t + 1}"] = accs[t].item()
        full = full + hyper["langweight"] * rewloss
    else:
        pi = 9
    if hyper["tcnweight"] > 0:
        l2d = hyper.get("l2dist", True)
        s02, s12, s01 = sim(es2, es0, l2d), sim(es2, es1, l2d), sim(es1, es0, l2d)
        neg0, neg2 = [], []
        for _ in range(3):
            neg0.append(sim(es0, es0[perms[pi]], l2d))
            neg2.append(sim(es2, es2[perms[pi + 1]], l2d))
            pi += 2
        neg0, neg2 = torch.stack(neg0, -1), torch.stack(neg2, -1)
        sm1 = -torch.log(EPSILON + torch.exp(s12) / (EPSILON + torch.exp(s02) + torch.exp(s12) + torch.exp(neg2).sum(-1)))
        sm2 = -torch.log(EPSILON + torch.exp(s01) / (EPSILON + torch.exp(s01) + torch.exp(s02) + torch.exp(neg0).sum(-1)))
        tcn = ((sm1 + sm2) / 2.0).mean()
        aligned = ((1.0 * (s02 < s12)) * (1.0 * (s01 > s02))).mean()
        metrics["tcnloss"], metrics["aligned"] = tcn.item(), aligned.item()
        full = full + hyper["tcnweight"] * tcn
    metrics["full_loss"] = full.item()
    return full, metrics


def adam_step(params, grads, opt_state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (defaults, weight_decay 0, amsgrad False) restated; opt_state = {'step', 'm', 'v'}."""
    opt_state["step"] += 1
    t = opt_state["step"]
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    for k, p in params.items():
        g = grads[k]
        m = opt_state["m"].setdefault(k, torch.zeros_like(p))
        v = opt_state["v"].setdefault(k, torch.zeros_like(p))
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def update(params, buffers, opt_state, frames, perms, hyper, size, lang_emb=None, lang_mask=None, eval_mode=False,
           policy="fp32"):
    """One ``Trainer.update`` (trainer.py:25-162).  Mutates params / buffers / opt_state in place; returns
    (metrics, grads, embeddings)."""
    bs = frames.shape[0]
    leaf = OrderedDict((k, v.detach().clone().requires_grad_(not eval_mode)) for k, v in params.items())
    with torch.set_grad_enabled(not eval_mode):
        alles = r3m_forward(leaf, buffers, frames.reshape(bs * 5, 3, 224, 224), size, train=not eval_mode,
                            policy=policy)
        full, metrics = losses(leaf, alles, perms, hyper, lang_emb, lang_mask)
    grads = None
    if not eval_mode:
        full.backward()
        grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v))) for k, v in leaf.items())
        with torch.no_grad():
            adam_step(params, grads, opt_state, hyper["lr"])
    return metrics, grads, alles.detach()


def new_opt_state():
    return {"step": 0, "m": {}, "v": {}}
