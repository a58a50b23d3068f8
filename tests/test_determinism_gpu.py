"""Run-to-run determinism and exact resume.  Every reduction of the step is ordered
— BatchNorm statistics (per-CTA partials + last-CTA finalize), BatchNorm-backward sums (two-level ordered tree), wgrad
split-K (per-split partial copies summed in split order), loss heads (gathers instead of scatter-adds) — so two runs
from the same state give bit-identical gradients, and a run resumed from a snapshot continues bit for bit."""
import pytest
import torch

from gpu_common import build_model, well_conditioned_state
from oracle import r3m_oracle as O

pytestmark = pytest.mark.gpu


def _one_step(size, clips, lang, steps=1, seed=0):
    from r3m_b200 import Trainer

    params, buffers = O.init_state(size, 80 + seed, lang=bool(lang))
    lang_emb = O.stub_lang_embedding(clips, 81) if lang else None
    m, model = build_model(size, params, buffers, float(lang), lang_emb)
    tr = Trainer(100)
    sentences = ["" if i % 4 == 3 else "s%d" % i for i in range(clips)]
    out = []
    for i in range(steps):
        frames = O.structured_frames(clips, 82 + i).cuda()
        perms = O.draw_permutations(clips, 90 + i)
        metrics, _ = tr.update(model, (frames, sentences), i, perms=perms, lang_emb=lang_emb)
        out.append((metrics, m._flat(1).clone(), m._any_engine().embeddings().clone()))
    return m, out


@pytest.mark.parametrize("size,clips,lang", [(18, 6, 1), (50, 12, 1), (34, 8, 0), (50, 64, 1)])
def test_two_runs_are_bit_identical(size, clips, lang):
    """Default (ill-conditioned) init on purpose: any reordered fp32 sum would be amplified to per-cent differences."""
    ma, a = _one_step(size, clips, lang, steps=2)
    mb, b = _one_step(size, clips, lang, steps=2)
    for (m1, g1, e1), (m2, g2, e2) in zip(a, b):
        assert torch.equal(e1, e2), "embeddings differ between two identical runs"
        assert torch.equal(g1, g2), "gradients differ between two identical runs"
        assert m1 == m2, (m1, m2)
    assert torch.equal(ma._flat(0), mb._flat(0)) and torch.equal(ma._flat(4), mb._flat(4))


def test_resume_from_snapshot_is_bit_exact(tmp_path):
    """train 3 steps == train 2, save_snapshot, build a fresh model, load_snapshot, train 1 — weights, Adam moments,
    BatchNorm running statistics, step counters and the permutation stream (torch's CPU generator) all continue."""
    import r3m_b200
    from r3m_b200 import Trainer

    size, clips = 18, 6
    params, buffers = well_conditioned_state(size, 100, True)
    lang_emb = O.stub_lang_embedding(clips, 101)
    sentences = ["s%d" % i for i in range(clips)]
    batches = [O.varied_frames(clips, 102 + i).cuda() for i in range(3)]

    def fresh():
        return build_model(size, params, buffers, 1.0, lang_emb)

    m_ref, model_ref = fresh()
    torch.manual_seed(7)
    tr = Trainer(100)
    for i in range(3):
        metrics_ref, _ = tr.update(model_ref, (batches[i], sentences), i)  # permutations from the global generator

    m_a, model_a = fresh()
    torch.manual_seed(7)
    tr = Trainer(100)
    for i in range(2):
        tr.update(model_a, (batches[i], sentences), i)
    path = str(tmp_path / "snapshot.pt")
    assert r3m_b200.save_snapshot(path, model_a, 2)
    torch.manual_seed(12345)  # the resumed process starts with an unrelated generator state
    m_b, model_b = fresh()
    step, _ = r3m_b200.load_snapshot(path, model_b)
    assert step == 2 and m_b.encoder_opt.steps == 2
    metrics_b, _ = Trainer(100).update(model_b, (batches[2], sentences), step)
    assert metrics_b == metrics_ref
    for which in (0, 2, 3, 4):  # parameters, Adam m, Adam v, BatchNorm running statistics
        assert torch.equal(m_b._flat(which), m_ref._flat(which)), which
    sd_b, sd_ref = model_b.state_dict(), model_ref.state_dict()
    assert all(torch.equal(sd_b[k], sd_ref[k]) for k in sd_ref)
    assert int(sd_b["module.convnet.bn1.num_batches_tracked"]) == 3


@pytest.mark.parametrize("scale", [1.0, 1e-6, 1e5])
def test_fixed_point_accumulator_is_exact_and_order_independent(lib, scale):
    """fx_add / fx_to_float (csrc/ptx.cuh): the sum of fp32 values through the Q64.64 accumulator equals the exactly
    rounded sum (math.fsum), whatever the number of blocks (grouping, arrival order) — including heavy cancellation."""
    import math

    g = torch.Generator().manual_seed(int(scale * 7) % 1000)
    x = torch.randn(200_000, generator=g) * scale
    x[::7] *= -1e3
    x[5] = 0.0
    x = torch.cat([x, -x[:100_000] * (1 + 2 ** -12)])  # near-cancelling pairs
    want = torch.tensor(math.fsum(x.double().tolist()), dtype=torch.float64).float()
    xd = x.cuda()
    outs = []
    for blocks in (1, 7, 148, 1024):
        out = torch.zeros(1, device="cuda")
        lib.check(lib.lib.r3m_b200_ordered_sum(lib.ptr(xd), xd.numel(), lib.ptr(out), blocks, lib.current_stream()))
        outs.append(out.cpu())
    assert all(torch.equal(o, outs[0]) for o in outs)
    assert torch.equal(outs[0], want.reshape(1)), (float(outs[0]), float(want))
    # a permutation of the inputs does not change a single bit either
    perm = torch.randperm(xd.numel(), generator=g).cuda()
    out = torch.zeros(1, device="cuda")
    xp = xd[perm].contiguous()
    lib.check(lib.lib.r3m_b200_ordered_sum(lib.ptr(xp), xp.numel(), lib.ptr(out), 148, lib.current_stream()))
    assert torch.equal(out.cpu(), outs[0])


@pytest.mark.parametrize("ratio", [1.0, 30.0, 300.0, 1000.0])
def test_batchnorm_variance_survives_large_means(lib, ratio):
    """one-pass E[x^2] - mean^2 in fp32 loses eps * (mean / std)^2 of the variance (6 % at
    |mean| = 1000 std).  The engine's statistics law (csrc/elementwise.cu: bn_batch_moments — exact fixed-point sums,
    fp64 subtraction) keeps it.  Inputs sit on a 12-bit grid like a low-precision conv output, so that their squares
    are exact in fp32 and the check against fp64 can be tight."""
    import math

    g = torch.Generator().manual_seed(int(ratio))
    grid = 2.0 ** math.floor(math.log2(4095.0 / (ratio + 6.0)))
    x = torch.round((torch.randn(400_000, generator=g).clamp(-5, 5) + ratio) * grid) / grid
    want_mean, want_var = x.double().mean(), x.double().var(unbiased=False)
    xd = x.cuda()
    outs = []
    for blocks in (1, 148, 1024):
        out = torch.zeros(2, device="cuda")
        lib.check(lib.lib.r3m_b200_ordered_moments(lib.ptr(xd), xd.numel(), lib.ptr(out), blocks, lib.current_stream()))
        outs.append(out.cpu())
    assert all(torch.equal(o, outs[0]) for o in outs)
    assert abs(float(outs[0][0]) - float(want_mean)) <= 2e-7 * abs(float(want_mean))
    assert abs(float(outs[0][1]) - float(want_var)) <= 1e-6 * float(want_var), (float(outs[0][1]), float(want_var))


@pytest.mark.parametrize("size,clips,lang", [(18, 6, 1), (50, 10, 1)])
def test_step_graph_replay_equals_plain_launches(monkeypatch, size, clips, lang):
    """Whole-step CUDA graphs (engine.cu: run_cached): from the second step with the same input buffers on, the train-mode
    forward and everything behind it (loss heads, language head, two-stream backward) are ONE cudaGraphLaunch each.  Same
    kernels in the same dependency order: weights, Adam moments and metrics must equal the plainly launched run bit for
    bit, and the replay must actually have happened."""
    from r3m_b200 import Trainer

    def run(graph):
        monkeypatch.setenv("R3M_STEP_GRAPH", "1" if graph else "0")
        params, buffers = well_conditioned_state(size, 120, bool(lang))
        lang_emb = O.stub_lang_embedding(clips, 121).cuda() if lang else None
        m, model = build_model(size, params, buffers, float(lang), lang_emb)
        tr = Trainer(100)
        sentences = ["s%d" % i for i in range(clips)]
        frames = torch.empty(clips, 5, 3, 224, 224, device="cuda")  # one buffer, refilled: the feeder's contract
        metrics = []
        for i in range(5):
            frames.copy_(O.varied_frames(clips, 122 + i).reshape(frames.shape))
            mt, _ = tr.update(model, (frames, sentences), i, perms=O.draw_permutations(clips, 130 + i), lang_emb=lang_emb)
            metrics.append(mt)
        eng = m._any_engine()
        return m, metrics, eng.graph_replays(), tr.last_launches

    m0, met0, rep0, n0 = run(False)
    m1, met1, rep1, n1 = run(True)
    assert rep0 == 0 and rep1 == 2 * 4, (rep0, rep1)  # steps 2-5 (captured on the second call): forward + backward graph
    assert n0 == n1, (n0, n1)  # the launch count reported for a replayed step is the captured one
    assert met0 == met1
    for which in (0, 2, 3, 4):
        assert torch.equal(m0._flat(which), m1._flat(which)), which


def test_schedule_variants_compute_the_same_step(monkeypatch):
    """The engine's schedule knobs change WHEN and in WHAT ORDER data is touched, never the arithmetic: the fused
    BatchNorm backward (one cooperative launch with a grid barrier, R3M_FUSE_BN_BWD=1) and the L2-aware traversal order
    (R3M_L2_ORDER=0 switches it off) regroup fp32 partial sums inside a block, so results are not bit-identical to the
    default schedule — but they must agree to the bf16 tier's own noise floor (a last-bit difference
    cascades through the bf16 re-roundings to a few per cent of the gradient), far below any wiring error (O(1))."""
    from r3m_b200 import Trainer

    size, clips = 50, 10

    def run(env):
        for k in ("R3M_FUSE_BN_BWD", "R3M_L2_ORDER"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        params, buffers = well_conditioned_state(size, 140, True)
        lang_emb = O.stub_lang_embedding(clips, 141)
        m, model = build_model(size, params, buffers, 1.0, lang_emb)
        sentences = ["s%d" % i for i in range(clips)]
        metrics, _ = Trainer(100).update(model, (O.varied_frames(clips, 142).cuda(), sentences), 0,
                                         perms=O.draw_permutations(clips, 150), lang_emb=lang_emb)
        return metrics, m._flat(1).double().clone(), m._any_engine().embeddings().double().clone()

    def rel(a, b):
        return float((a - b).norm() / b.norm())

    met0, g0, e0 = run({})
    met1, g1, e1 = run({"R3M_FUSE_BN_BWD": "1"})
    assert torch.equal(e0, e1) and met0 == met1  # the forward pass is untouched by the backward variant
    assert rel(g1, g0) < 0.06, rel(g1, g0)
    met2, g2, e2 = run({"R3M_L2_ORDER": "0"})
    assert rel(e2, e0) < 5e-3 and rel(g2, g0) < 0.08, (rel(e2, e0), rel(g2, g0))
    assert all(abs(met2[k] - met0[k]) <= 2e-2 * max(1.0, abs(met0[k])) for k in met0)
