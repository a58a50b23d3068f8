"""Engine-level diagnostic on a B200: R3M.forward / Trainer.update against the CPU oracle and the golden fixtures.
Each case runs in its own subprocess; results go to gpurun_out/engine_check.jsonl."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def load_oracle_state(model, params, buffers):
    sd = {k: v for k, v in params.items()}
    sd.update(buffers)
    model.load_state_dict(sd)


def case_eval18():
    import numpy as np
    import torch
    from oracle import r3m_oracle as O
    from r3m_b200 import R3M

    params, buffers = O.init_state(18, 5)
    g = torch.Generator().manual_seed(6)
    for k in buffers:
        if k.endswith("running_mean"):
            buffers[k] = 0.1 * torch.randn(buffers[k].shape, generator=g)
        elif k.endswith("running_var"):
            buffers[k] = 0.5 + torch.rand(buffers[k].shape, generator=g)
    frames = O.synthetic_frames(1, 7)[0, :4]
    gold = np.load(os.path.join(ROOT, "tests/golden/rn18_eval_b4.npz"))
    assert abs(float(frames.double().sum()) - float(gold["frames_checksum"][0])) < 1e-3
    m = R3M("cuda", 1e-4, 1024, size=18, langweight=0.0)
    load_oracle_state(m, params, buffers)
    m = torch.nn.DataParallel(m.cuda(), device_ids=[0])
    m.eval()
    with torch.no_grad():
        out = m(frames.cuda())
    ref = torch.from_numpy(gold["embeddings"])
    return {"emb_rel": rel(out.cpu(), ref), "finite": bool(torch.isfinite(out).all())}


def case_update(size, clips, seeds, frames_kind, golden=None, lang=False):
    import numpy as np
    import torch
    import r3m_b200
    from oracle import r3m_oracle as O
    from r3m_b200 import R3M, Trainer

    sw, sf, sp, sl = seeds
    params, buffers = O.init_state(size, sw, lang=lang)
    frames = (O.synthetic_frames if frames_kind == "randint" else O.structured_frames)(clips, sf)
    perms = O.draw_permutations(clips, sp)
    lang_emb = O.stub_lang_embedding(clips, sl) if lang else None
    sentences = ["C does something %d" % i for i in range(clips)]
    if lang and clips >= 4:
        sentences[1] = ""
    if not lang:
        sentences = [""] * clips
    lang_mask = torch.tensor([1.0 * (x != "") for x in sentences]) if lang else None
    hyper = dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, langweight=1.0 if lang else 0.0, lr=1e-4)
    r3m_b200.set_lang_encoder_factory(lambda dev: (lambda s_: lang_emb))
    m = R3M("cuda", hyper["lr"], 1024, size=size, l2weight=1e-5, l1weight=1e-5, langweight=hyper["langweight"],
            tcnweight=1.0)
    load_oracle_state(m, params, buffers)
    model = torch.nn.DataParallel(m.cuda(), device_ids=[0])
    res = {}
    # train-mode forward parity first (also exercises running-stat updates on a scratch copy of the buffers)
    o_params = {k: v.clone() for k, v in params.items()}
    o_buffers = {k: v.clone() for k, v in buffers.items()}
    o_metrics, o_grads, o_emb = O.update(o_params, o_buffers, O.new_opt_state(), frames, perms, hyper, size, lang_emb,
                                         lang_mask)
    metrics, st = Trainer(eval_freq=100).update(model, (frames.cuda(), sentences), step=0, perms=perms,
                                                lang_emb=lang_emb)
    eng = m._any_engine()
    emb = eng.embeddings().cpu()
    res["emb_rel_vs_oracle"] = rel(emb, o_emb)
    res["metrics"] = metrics
    res["oracle_metrics"] = dict(o_metrics)
    res["launches_adam"] = eng.launches()
    # loss heads on identical embeddings: feed OUR embeddings to the oracle's loss code
    full, lm = O.losses(params, emb.clone().requires_grad_(True), perms, hyper, lang_emb, lang_mask)
    res["loss_rel_same_emb"] = {k: abs(metrics[k] - lm[k]) / (abs(lm[k]) + 1e-12) for k in lm}
    e2 = emb.clone().requires_grad_(True)
    lp = {k: v.clone().requires_grad_(True) for k, v in params.items() if k.startswith("lang_rew")}
    full, _ = O.losses(lp, e2, perms, hyper, lang_emb, lang_mask)
    full.backward()
    res["dE_rel_same_emb"] = rel(eng.embedding_grads().cpu(), e2.grad)
    if lang:
        res["lang_param_grad_rel_same_emb"] = {k: rel(dict(m.named_parameters())[k].grad.cpu(), v.grad)
                                               for k, v in lp.items()}
    # gradients vs the oracle's (fp32 autograd through the whole network)
    named = dict(m.named_parameters())
    gr = {k: rel(named[k].grad.cpu(), o_grads[k]) for k in o_grads}
    cat = lambda d: torch.cat([d[k].flatten().double() for k in o_grads])  # noqa: E731
    res["grad_global_rel"] = rel(cat({k: named[k].grad.cpu() for k in o_grads}), cat(o_grads))
    worst = sorted(gr.items(), key=lambda kv: -kv[1])[:8]
    res["grad_worst"] = worst
    res["grad_rel_conv1"] = gr["convnet.conv1.weight"]
    res["grad_rel_last"] = [v for k, v in gr.items() if k.startswith("convnet.layer4")][-3:]
    # isolate the backward pass from the loss head's sensitivity to embedding noise: push OUR dE through the
    # oracle network's autograd (vector-Jacobian product) and compare parameter gradients
    leaf = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    alles = O.r3m_forward(leaf, {k: v.clone() for k, v in buffers.items()}, frames.reshape(-1, 3, 224, 224), size, True)
    alles.backward(eng.embedding_grads().cpu())
    vjp = {k: leaf[k].grad for k in o_grads if leaf[k].grad is not None}
    gv = {k: rel(named[k].grad.cpu(), vjp[k]) for k in vjp}
    catv = lambda d: torch.cat([d[k].flatten().double() for k in vjp])  # noqa: E731
    res["vjp_grad_global_rel"] = rel(catv({k: named[k].grad.cpu() for k in vjp}), catv(vjp))
    res["vjp_grad_worst"] = sorted(gv.items(), key=lambda kv: -kv[1])[:6]
    res["vjp_grad_by_layer"] = [(k, round(v, 4)) for k, v in gv.items() if k.endswith("conv1.weight") or k.endswith("bn1.weight") or "downsample" in k][:40]
    # the reference under the SAME storage policy (bf16 rounding at the points where the CUDA path stores bf16)
    b_params = {k: v.clone() for k, v in params.items()}
    b_buffers = {k: v.clone() for k, v in buffers.items()}
    b_metrics, b_grads, b_emb = O.update(b_params, b_buffers, O.new_opt_state(), frames, perms, hyper, size,
                                         lang_emb, lang_mask, policy="bf16")
    res["emb_rel_vs_bf16_oracle"] = rel(emb, b_emb)
    res["bf16_oracle_emb_rel_vs_fp32"] = rel(b_emb, o_emb)
    res["bf16_oracle_metrics"] = dict(b_metrics)
    gb = {k: rel(named[k].grad.cpu(), b_grads[k]) for k in o_grads}
    res["grad_global_rel_vs_bf16_oracle"] = rel(cat({k: named[k].grad.cpu() for k in o_grads}), cat(b_grads))
    res["bf16_oracle_grad_global_rel_vs_fp32"] = rel(cat(b_grads), cat(o_grads))
    res["grad_worst_vs_bf16_oracle"] = sorted(gb.items(), key=lambda kv: -kv[1])[:6]
    sd = m.state_dict()
    res["running_mean_rel_bn1"] = rel(sd["convnet.bn1.running_mean"].cpu(), o_buffers["convnet.bn1.running_mean"])
    res["running_var_rel_bn1"] = rel(sd["convnet.bn1.running_var"].cpu(), o_buffers["convnet.bn1.running_var"])
    d_ours = cat({k: sd[k].cpu() - params[k] for k in o_grads})
    d_orc = cat({k: o_params[k] - params[k] for k in o_grads})
    res["adam_delta_global_rel"] = rel(d_ours, d_orc)
    if golden:
        gold = np.load(os.path.join(ROOT, "tests/golden", golden + ".npz"))
        gm = json.loads(bytes(gold["metrics_json"]).decode())
        res["golden_metrics"] = gm
        res["emb_rel_vs_golden"] = rel(emb, torch.from_numpy(gold["embeddings"]))
    return res


CASES = {
    "eval18": lambda: case_eval18(),
    "update18": lambda: case_update(18, 2, (0, 11, 12, 13), "randint", "rn18_tcn"),
    "update34": lambda: case_update(34, 2, (1, 21, 22, 23), "structured", "rn34_tcn"),
    "update50": lambda: case_update(50, 2, (2, 31, 32, 33), "structured", None),
    "lang50": lambda: case_update(50, 2, (2, 31, 32, 33), "structured", "rn50_lang", lang=True),
    "lang18": lambda: case_update(18, 4, (3, 41, 42, 43), "randint", "rn18_lang_b4", lang=True),
}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        try:
            res = CASES[sys.argv[2]]()
            res["status"] = "ok"
        except Exception as e:  # noqa: BLE001
            import traceback

            res = {"status": "error", "error": repr(e)[:800], "tb": traceback.format_exc()[-1500:]}
        res["case"] = sys.argv[2]
        print("RESULT " + json.dumps(res))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    names = sys.argv[1:] or list(CASES)
    with open(os.path.join(ROOT, "gpurun_out", "engine_check.jsonl"), "w") as f:
        for name in names:
            try:
                p = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True,
                                   timeout=600)
                line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                res = json.loads(line[-1][7:]) if line else {"status": "crash", "case": name, "rc": p.returncode,
                                                             "stderr": p.stderr[-1500:]}
            except subprocess.TimeoutExpired:
                res = {"status": "timeout", "case": name}
            f.write(json.dumps(res) + "\n")
            f.flush()
            print(json.dumps({k: res.get(k) for k in ("case", "status", "error", "emb_rel", "emb_rel_vs_oracle",
                                                        "emb_rel_vs_bf16_oracle", "bf16_oracle_emb_rel_vs_fp32",
                                                        "grad_global_rel", "grad_global_rel_vs_bf16_oracle",
                                                        "bf16_oracle_grad_global_rel_vs_fp32",
                                                        "grad_worst_vs_bf16_oracle", "metrics", "loss_rel_same_emb", "dE_rel_same_emb",
                                                        "lang_param_grad_rel_same_emb", "golden_metrics",
                                                        "bf16_oracle_metrics", "adam_delta_global_rel", "tb")}),
                  flush=True)


if __name__ == "__main__":
    main()
