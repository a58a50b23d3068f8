"""Helpers shared by the GPU parity tests: build an r3m_b200 model from oracle state, run the oracle ON the GPU in strict
fp32 (so that BASELINE-size cases finish in seconds), per-layer-group gradient distances."""
import torch

from oracle import r3m_oracle as O

HYPER = dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, lr=1e-4)
GROUPS = ("convnet.layer4", "convnet.layer3", "convnet.layer2", "convnet.layer1", "convnet.conv1", "convnet.bn1",
          "convnet.")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False


def build_model(size, params, buffers, langweight, lang_emb=None, l2dist=True):
    """(R3M, DataParallel(R3M)) on cuda with the oracle's state loaded; the language encoder returns `lang_emb`."""
    import r3m_b200
    from r3m_b200 import R3M

    r3m_b200.set_lang_encoder_factory(lambda dev: (lambda s: lang_emb))
    m = R3M("cuda", HYPER["lr"], 1024, size=size, l2weight=HYPER["l2weight"], l1weight=HYPER["l1weight"],
            langweight=langweight, tcnweight=HYPER["tcnweight"], l2dist=l2dist)
    sd = dict(params)
    sd.update(buffers)
    m.load_state_dict(sd)
    return m, torch.nn.DataParallel(m.cuda(), device_ids=[torch.cuda.current_device()])


def oracle_update_on_gpu(size, params, buffers, frames, perms, hyper, lang_emb=None, mask=None, policy="fp32",
                         eval_mode=False):
    """O.update with every tensor on the GPU, TF32 off.  Returns (metrics, grads, embeddings, post_params, post_buffers)."""
    strict_fp32()
    dev = torch.device("cuda")
    p = {k: v.to(dev) for k, v in params.items()}
    b = {k: v.to(dev) for k, v in buffers.items()}
    m, g, e = O.update(p, b, O.new_opt_state(), frames.to(dev), perms.to(dev), hyper, size,
                       None if lang_emb is None else lang_emb.to(dev), None if mask is None else mask.to(dev),
                       eval_mode=eval_mode, policy=policy)
    return m, g, e, p, b


def group_distances(ours, ref, groups=GROUPS):
    """{group prefix: relative L2 distance of the concatenated gradients}."""
    out = {}
    for pre in groups:
        ks = [k for k in ref if k.startswith(pre) and k in ours]
        if not ks:
            continue
        cat = lambda d: torch.cat([d[k].detach().flatten().double().cpu() for k in ks])  # noqa: E731
        out[pre] = rel(cat(ours), cat(ref))
    return out


def well_conditioned_state(size, seed, lang):
    params, buffers = O.init_state(size, seed, lang=lang)
    O.scale_last_gamma(params, size, 0.1)
    return params, buffers
