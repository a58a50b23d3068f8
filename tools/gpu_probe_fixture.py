"""Round-2 probe: which fixture makes the end-to-end gradient comparison well conditioned?  For each candidate
(initial state x frames) reports the bf16-policy oracle's and OUR distance from the fp32 oracle (all on the GPU)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r3m_oracle as O  # noqa: E402
from oracle import torch_reference as T  # noqa: E402
import r3m_b200  # noqa: E402
from r3m_b200 import R3M, Trainer  # noqa: E402

dev = torch.device("cuda")
T.configure("fp32_strict")
torch.backends.cudnn.benchmark = False
HY = dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, lr=1e-4)
rows = []


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def varied_frames(clips, seed):
    """structured frames with a per-frame colour gain / offset: embeddings of different frames differ a lot"""
    g = torch.Generator().manual_seed(seed + 1000)
    f = O.structured_frames(clips, seed)
    gain = torch.rand(clips, 5, 3, 1, 1, generator=g) * 1.4 + 0.1
    return (f * gain).clamp(0, 255).round()


def last_gamma(params, size, value):
    tail = "bn3.weight" if size == 50 else "bn2.weight"
    for k in params:
        if k.startswith("convnet.layer") and k.endswith(tail):
            params[k] = torch.full_like(params[k], value)


def train_some(params, buffers, size, lang, steps, lr, clips):
    p = {k: v.to(dev) for k, v in params.items()}
    b = {k: v.to(dev) for k, v in buffers.items()}
    opt = O.new_opt_state()
    hyper = dict(HY, langweight=float(lang), lr=lr)
    for i in range(steps):
        fr = varied_frames(clips, 500 + i).to(dev)
        le = O.stub_lang_embedding(clips, 700 + i).to(dev) if lang else None
        mask = torch.ones(clips, device=dev) if lang else None
        O.update(p, b, opt, fr, O.draw_permutations(clips, 600 + i).to(dev), hyper, size, le, mask)
    return {k: v.cpu() for k, v in p.items()}, {k: v.cpu() for k, v in b.items()}


def evaluate(tag, size, clips, lang, params, buffers, frames):
    perms = O.draw_permutations(clips, 9)
    le = O.stub_lang_embedding(clips, 10) if lang else None
    sent = ["" if i % 10 == 9 else "s%d" % i for i in range(clips)] if lang else [""] * clips
    mask = torch.tensor([1.0 * (s != "") for s in sent])
    hyper = dict(HY, langweight=float(lang))
    r3m_b200.set_lang_encoder_factory(lambda d: (lambda s: le))
    m = R3M("cuda", 1e-4, 1024, size=size, l2weight=1e-5, l1weight=1e-5, langweight=float(lang), tcnweight=1.0)
    sd = dict(params)
    sd.update(buffers)
    m.load_state_dict(sd)
    model = torch.nn.DataParallel(m.cuda(), device_ids=[0])
    Trainer(100).update(model, (frames.cuda(), sent), 0, perms=perms, lang_emb=le)
    ours_e = m._any_engine().embeddings().clone()
    ours_g = {k: v.grad.detach().clone() for k, v in m.named_parameters()}
    out = {}
    for pol in ("fp32", "bf16"):
        p = {k: v.to(dev) for k, v in params.items()}
        b = {k: v.to(dev) for k, v in buffers.items()}
        out[pol] = O.update(p, b, O.new_opt_state(), frames.to(dev), perms.to(dev), hyper, size,
                            le.to(dev) if lang else None, mask.to(dev) if lang else None, policy=pol)
    row = {"tag": tag, "size": size, "clips": clips, "lang": lang,
           "emb": [rel(ours_e, out["fp32"][2]), rel(out["bf16"][2], out["fp32"][2]), rel(ours_e, out["bf16"][2])]}
    for pre in ("convnet.layer4", "convnet.layer2", "convnet.layer1", "convnet.conv1", "convnet.", "lang_rew"):
        ks = [k for k in out["fp32"][1] if k.startswith(pre)]
        if not ks:
            continue
        cat = lambda g: torch.cat([g[k].flatten().double() for k in ks])  # noqa: E731
        row[pre] = [round(rel(cat(ours_g), cat(out["fp32"][1])), 4), round(rel(cat(out["bf16"][1]), cat(out["fp32"][1])), 4),
                    round(rel(cat(ours_g), cat(out["bf16"][1])), 4)]
    row["metrics_fp32"] = {k: round(v, 4) for k, v in out["fp32"][0].items()}
    print(json.dumps(row), flush=True)
    rows.append(row)
    with open(os.path.join(ROOT, "gpurun_out", "r2_fixture_probe.json"), "w") as f:
        json.dump(rows, f, indent=1)
    del m, model, out, ours_g
    torch.cuda.empty_cache()


for size, clips, lang in ((18, 8, 0), (50, 8, 1)):
    base_p, base_b = O.init_state(size, 7, lang=bool(lang))
    fr_s, fr_v = O.structured_frames(clips, 8), varied_frames(clips, 8)
    evaluate("default/structured", size, clips, lang, base_p, base_b, fr_s)
    evaluate("default/varied", size, clips, lang, base_p, base_b, fr_v)
    for gval in (0.3, 0.1):
        p = {k: v.clone() for k, v in base_p.items()}
        last_gamma(p, size, gval)
        evaluate(f"gamma_last={gval}/varied", size, clips, lang, p, base_b, fr_v)
    for steps, lr in ((60, 1e-3), (200, 1e-3)):
        p, b = train_some(base_p, base_b, size, lang, steps, lr, clips)
        evaluate(f"trained{steps}@{lr}/varied", size, clips, lang, p, b, fr_v)
    p = {k: v.clone() for k, v in base_p.items()}
    last_gamma(p, size, 0.1)
    p, b = train_some(p, base_b, size, lang, 60, 1e-3, clips)
    evaluate("gamma_last=0.1+trained60/varied", size, clips, lang, p, b, fr_v)
