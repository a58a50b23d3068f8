"""Round-2 probe (GPU box): (1) the same-box torch/cuDNN reference timings, (2) ours vs fp32 / bf16-policy oracle run ON
the GPU at growing sizes (embedding + per-layer-group gradient distances), written to gpurun_out/r2_probe.json."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import r3m_oracle as O  # noqa: E402
from oracle import torch_reference as T  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
dev = torch.device("cuda")
res = {"timing": [], "parity": []}


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def dump():
    with open(os.path.join(OUT, "r2_probe.json"), "w") as f:
        json.dump(res, f, indent=1)


what = sys.argv[1:] or ["timing", "parity"]
if "timing" in what:
    for variant in ("as_written", "bf16_channels_last"):
        for fn, kw in ((T.time_update, dict(size=50, clips=64, lang=True)),
                       (T.time_update, dict(size=34, clips=128, lang=True)),
                       (T.time_forward, dict(size=50, batch=256, train_bn=False)),
                       (T.time_forward, dict(size=50, batch=256, train_bn=True))):
            try:
                r = fn(variant=variant, device=dev, **kw)
            except Exception as e:  # noqa: BLE001
                r = {"variant": variant, "error": repr(e), **kw}
            r["fn"] = fn.__name__
            print(r, flush=True)
            res["timing"].append(r)
            dump()

if "parity" in what:
    import r3m_b200
    from r3m_b200 import R3M, Trainer

    T.configure("fp32_strict")
    torch.backends.cudnn.benchmark = False
    HY = dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, lr=1e-4)
    for size, clips, lang, kind in ((18, 4, 1, "randint"), (18, 32, 0, "structured"), (50, 8, 1, "structured"),
                                    (50, 32, 1, "structured"), (34, 32, 0, "structured"), (50, 64, 1, "randint")):
        t0 = time.time()
        params, buffers = O.init_state(size, 7, lang=bool(lang))
        frames = (O.synthetic_frames if kind == "randint" else O.structured_frames)(clips, 8)
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
        metrics, _ = Trainer(100).update(model, (frames.cuda(), sent), 0, perms=perms, lang_emb=le)
        ours_e = m._any_engine().embeddings().clone()
        ours_g = {k: v.grad.detach().clone() for k, v in m.named_parameters()}
        out = {}
        for pol in ("fp32", "bf16"):
            p = {k: v.to(dev) for k, v in params.items()}
            b = {k: v.to(dev) for k, v in buffers.items()}
            om, og, oe = O.update(p, b, O.new_opt_state(), frames.to(dev), perms.to(dev), hyper, size,
                                  le.to(dev) if lang else None, mask.to(dev) if lang else None, policy=pol)
            out[pol] = (om, og, oe)
            torch.cuda.empty_cache()
        keys = [k for k in out["fp32"][1] if k.startswith("convnet.")]
        row = {"size": size, "clips": clips, "lang": lang, "frames": kind,
               "emb": {"ours_fp32": rel(ours_e, out["fp32"][2]), "pol_fp32": rel(out["bf16"][2], out["fp32"][2]),
                       "ours_pol": rel(ours_e, out["bf16"][2])}, "grad": {},
               "metrics": {k: (metrics[k], out["fp32"][0][k], out["bf16"][0][k]) for k in metrics}}
        for pre in ("convnet.layer4", "convnet.layer3", "convnet.layer2", "convnet.layer1", "convnet.conv1",
                    "convnet.", "lang_rew"):
            ks = [k for k in out["fp32"][1] if k.startswith(pre)]
            if not ks:
                continue
            cat = lambda g: torch.cat([g[k].flatten().double() for k in ks])  # noqa: E731
            row["grad"][pre] = {"ours_fp32": rel(cat(ours_g), cat(out["fp32"][1])),
                                "pol_fp32": rel(cat(out["bf16"][1]), cat(out["fp32"][1])),
                                "ours_pol": rel(cat(ours_g), cat(out["bf16"][1]))}
        row["seconds"] = time.time() - t0
        print(json.dumps(row), flush=True)
        res["parity"].append(row)
        dump()
        del m, model, out, ours_g
        torch.cuda.empty_cache()
