"""Drop-in boundary on the GPU: the reference's own update step (a host that builds the loss in torch and calls
``full_loss.backward()`` / ``encoder_opt.step()``, r3m/trainer.py:41-158) driving ``r3m_b200.R3M`` through autograd,
and the engine-state rules around eval / train forwards."""
import pytest
import torch

from gpu_common import HYPER, build_model, group_distances, rel, well_conditioned_state
from oracle import r3m_oracle as O
from oracle import torch_reference as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,lang", [(18, 1), (50, 1), (34, 0)])
def test_reference_style_trainer_drives_r3m_through_autograd(size, lang):
    """oracle/torch_reference.update is the reference's Trainer.update restated on `model(...)`, `model.get_reward`,
    `model.sim`, `full_loss.backward()`, `model.encoder_opt.step()`.  Run it on r3m_b200.R3M (autograd through the
    engine, language head through torch on the same master weights) and on the fused r3m_b200.Trainer: same metrics,
    same gradients, same post-step weights."""
    from r3m_b200 import Trainer

    clips = 6
    params, buffers = well_conditioned_state(size, 50, bool(lang))
    frames = O.varied_frames(clips, 51).cuda()
    perms = O.draw_permutations(clips, 52)
    lang_emb = O.stub_lang_embedding(clips, 53) if lang else None
    sentences = ["s%d" % i for i in range(clips)]
    if lang:
        sentences[2] = ""

    m1, model1 = build_model(size, params, buffers, float(lang), lang_emb)
    metrics1, _ = Trainer(100).update(model1, (frames, sentences), 0, perms=perms, lang_emb=lang_emb)
    g1 = {k: v.grad.detach().clone() for k, v in m1.named_parameters()}

    m2, _ = build_model(size, params, buffers, float(lang), lang_emb.cuda() if lang else None)
    metrics2, emb2 = T.update(m2, (frames, sentences), perms=perms)  # autograd path: forward -> torch loss -> backward
    g2 = {k: v.grad.detach().clone() for k, v in m2.named_parameters()}
    assert emb2.shape == (clips * 5, m2.outdim)
    assert set(metrics1) == set(metrics2)
    for k in metrics1:
        if k.startswith("rewacc") or k == "aligned":
            assert abs(metrics1[k] - metrics2[k]) <= 1.0 / clips + 1e-6
        else:
            assert abs(metrics1[k] - metrics2[k]) <= 1e-3 * abs(metrics2[k]) + 1e-6, (k, metrics1[k], metrics2[k])
    # Same engine kernels below the embeddings, but dE comes from torch autograd instead of the fused loss kernels:
    # it differs in the last fp32 bits, some bf16 roundings of the stored gradients then land on the other side, and
    # the cascade saturates at the bf16 tier's own noise floor (3-5 % on this state, see test_fullsize_gpu.py) — the
    # two paths are as close to each other as either is to the fp32 reference.
    dist = group_distances(g2, g1)
    assert all(d < 0.1 for d in dist.values()), dist
    if lang:
        assert group_distances(g2, g1, ("lang_rew",))["lang_rew"] < 0.1
    # the optimiser stepped: weights moved by ~lr along -sign(g), identically (up to that noise) on both paths
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    k = "convnet.layer1.0.conv1.weight"
    assert 0.2e-4 < float((sd2[k].cpu() - params[k]).abs().mean()) < 1.1e-4
    agree = ((sd1[k].cpu() - params[k]).sign() == (sd2[k].cpu() - params[k]).sign()).float().mean()
    assert float(agree) > 0.9
    assert int(sd2["convnet.bn1.num_batches_tracked"]) == 1 and m2.encoder_opt.steps == 1


def test_autograd_accumulates_and_detects_stale_activations():
    from r3m_b200._lib import R3MB200Error

    params, buffers = well_conditioned_state(18, 54, False)
    m, _ = build_model(18, params, buffers, 0.0)
    m.train()
    x = O.varied_frames(1, 55).reshape(5, 3, 224, 224).cuda()
    e = m(x)
    assert e.requires_grad and e.grad_fn is not None
    m.encoder_opt.zero_grad()
    e.sum().backward()
    g1 = m.convnet.layer1._modules["0"].conv1.weight.grad.clone()
    assert float(g1.abs().max()) > 0
    e = m(x)
    e.sum().backward()  # no zero_grad: torch semantics accumulate
    g2 = m.convnet.layer1._modules["0"].conv1.weight.grad
    assert torch.equal(g2, 2 * g1)  # every reduction is ordered: the second pass repeats the first bit for bit
    # a second forward of the same frame count overwrites the saved activations of the first
    e_old = m(x)
    m(x)
    with pytest.raises(R3MB200Error):
        e_old.sum().backward()
    # no graph under no_grad / in eval mode
    with torch.no_grad():
        assert not m(x).requires_grad
    m.eval()
    assert not m(x).requires_grad


def test_eval_forward_invalidates_a_pending_train_forward():
    """forward(train) -> eval forward on the same engine -> update_grads(obs=NULL) must be refused (the eval
    pass overwrote the activations and the saved-statistics slots), on the CUDA-graph path (<= 16 frames) too."""
    from r3m_b200._lib import R3MB200Error

    params, buffers = O.init_state(18, 56)
    m, model = build_model(18, params, buffers, 0.0)
    x = O.synthetic_frames(1, 57).reshape(5, 3, 224, 224).cuda()
    eng = m._engine(5)
    perms = O.draw_permutations(1, 58).to(torch.int32).cuda()
    for _ in range(3):  # the third eval call replays the captured graph
        eng.forward_train_async(x)
        eng.forward(x, False)
        with pytest.raises(R3MB200Error):
            eng.update_grads(None, perms, None, None, 1e-5, 1e-5, 0.0, 1.0, False)
    eng.forward_train_async(x)
    eng.update_grads(None, perms, None, None, 1e-5, 1e-5, 0.0, 1.0, False)
    assert all(v == v for v in eng.read_metrics()[:10])


def test_language_model_embeds_any_number_of_frames():
    """R3M(langweight > 0).forward must embed N frames for any N, like the reference (the 5-frames-per-clip
    rule belongs to update() only)."""
    from r3m_b200._lib import R3MB200Error

    params, buffers = O.init_state(18, 59, lang=True)
    lang_emb = O.stub_lang_embedding(1, 60)
    m, model = build_model(18, params, buffers, 1.0, lang_emb)
    m.eval()
    x = O.synthetic_frames(1, 61).reshape(5, 3, 224, 224).cuda()
    with torch.no_grad():
        e3, e5 = m(x[:3]), m(x)
    assert e3.shape == (3, 512) and rel(e3, e5[:3]) < 1e-6
    with pytest.raises(R3MB200Error):
        m._engine(3).update_grads(x[:3].contiguous(), torch.zeros(15, 1, dtype=torch.int32, device="cuda"), None, None,
                                  1e-5, 1e-5, 0.0, 1.0, True)


@pytest.mark.parametrize("size", [18, 50])
def test_gradient_chunks_partition_the_flat_buffer_in_completion_order(size):
    """The overlapped all-reduce reduces the gradient buffer chunk by chunk (r3m_b200/trainer.py:allreduce_gradients):
    the chunks must tile [0, num_params) exactly, last layers first, and waiting for a chunk on another stream must see
    the finished gradients of that chunk (compared with a fully synchronised read)."""
    from r3m_b200 import Trainer

    params, buffers = well_conditioned_state(size, 60, True)
    lang_emb = O.stub_lang_embedding(4, 61)
    m, model = build_model(size, params, buffers, 1.0, lang_emb)
    frames = O.varied_frames(4, 62).cuda()
    perms = O.draw_permutations(4, 63)
    eng = m._engine(20)
    chunks = eng.grad_chunks()
    assert len(chunks) == 4
    assert chunks[0][1] == m._layout.num_params and chunks[-1][0] == 0
    assert all(chunks[k][0] == chunks[k + 1][1] for k in range(3)) and all(b < e for b, e in chunks)
    names = {t.offset: t.name for t in m._layout.tensors if t.kind in (0, 1)}
    assert [names[b] for b, _ in chunks] == ["convnet.layer4.0.conv1.weight", "convnet.layer3.0.conv1.weight",
                                             "convnet.layer2.0.conv1.weight", "convnet.conv1.weight"]
    G = m._flat(1)
    side = torch.cuda.Stream()
    copies = []
    # enqueue forward + losses + backward, then read every chunk on the side stream as soon as IT is final
    eng.forward_train_async(frames.reshape(-1, 3, 224, 224).contiguous())
    eng.update_grads(None, perms.to(torch.int32).cuda(), lang_emb.cuda(), torch.ones(4, device="cuda"), 1e-5, 1e-5, 1.0,
                     1.0, False)
    for k, (b, e) in enumerate(chunks):
        eng.wait_grad_chunk(k, side)
        with torch.cuda.stream(side):
            copies.append(G[b:e].clone())
    torch.cuda.synchronize()
    for (b, e), c in zip(chunks, copies):
        assert torch.equal(c, G[b:e]) and float(c.abs().max()) > 0
