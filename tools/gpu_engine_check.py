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
    m = torch.nn.DataParallel(m).cuda()
    m.eval()
    with torch.no_grad():
        out = m(frames.cuda())
    ref = torch.from_numpy(gold["embeddings"])
    return {"emb_rel": rel(out.cpu(), ref), "finite": bool(torch.isfinite(out).all())}


def case_update(size, clips, seeds, frames_kind, golden=None):
    import numpy as np
    import torch
    from oracle import r3m_oracle as O
    from r3m_b200 import R3M, Trainer

    sw, sf, sp, _sl = seeds
    params, buffers = O.init_state(size, sw)
    frames = (O.synthetic_frames if frames_kind == "randint" else O.structured_frames)(clips, sf)
    perms = O.draw_permutations(clips, sp)
    hyper = dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, langweight=0.0, lr=1e-4)
    m = R3M("cuda", hyper["lr"], 1024, size=size, l2weight=1e-5, l1weight=1e-5, langweight=0.0, tcnweight=1.0)
    load_oracle_state(m, params, buffers)
    model = torch.nn.DataParallel(m).cuda()
    res = {}
    # train-mode forward parity first (also exercises running-stat updates on a scratch copy of the buffers)
    o_params = {k: v.clone() for k, v in params.items()}
    o_buffers = {k: v.clone() for k, v in buffers.items()}
    o_metrics, o_grads, o_emb = O.update(o_params, o_buffers, O.new_opt_state(), frames, perms, hyper, size)
    metrics, st = Trainer(eval_freq=100).update(model, (frames.cuda(), [""] * clips), step=0, perms=perms)
    eng = m._any_engine()
    emb = eng.embeddings().cpu()
    res["emb_rel_vs_oracle"] = rel(emb, o_emb)
    res["metrics"] = metrics
    res["oracle_metrics"] = dict(o_metrics)
    res["launches_adam"] = eng.launches()
    # loss heads on identical embeddings: feed OUR embeddings to the oracle's loss code
    full, lm = O.losses(o_params, emb.clone().requires_grad_(True), perms, hyper)
    res["loss_rel_same_emb"] = {k: abs(metrics[k] - lm[k]) / (abs(lm[k]) + 1e-12) for k in lm}
    e2 = emb.clone().requires_grad_(True)
    full, _ = O.losses(o_params, e2, perms, hyper)
    full.backward()
    res["dE_rel_same_emb"] = rel(eng.embedding_grads().cpu(), e2.grad)
    # gradients vs the oracle's (fp32 autograd through the whole network)
    named = dict(m.named_parameters())
    gr = {k: rel(named[k].grad.cpu(), o_grads[k]) for k in o_grads}
    cat = lambda d: torch.cat([d[k].flatten().double() for k in o_grads])  # noqa: E731
    res["grad_global_rel"] = rel(cat({k: named[k].grad.cpu() for k in o_grads}), cat(o_grads))
    worst = sorted(gr.items(), key=lambda kv: -kv[1])[:8]
    res["grad_worst"] = worst
    res["grad_rel_conv1"] = gr["convnet.conv1.weight"]
    res["grad_rel_last"] = [v for k, v in gr.items() if k.startswith("convnet.layer4")][-3:]
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
            print(json.dumps(res, indent=None)[:3000], flush=True)


if __name__ == "__main__":
    main()
