"""Multi-GPU parity (BASELINE c5 shape): one process per GPU over NCCL.  For identical initial weights
the post-all-reduce gradient must equal the MEAN over ranks of the single-GPU gradient on each rank's shard with that
rank's permutations — checked against (a) the engine's own single-GPU gradients and (b) the fp32 oracle run on each
shard — and every rank must hold bit-identical weights after the step.  Needs >= 2 GPUs (gpurun --gpus 2); the
collective it replaces: nn.DataParallel's gradient reduction, r3m/train_representation.py:30."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, size, clips, lang, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gpu_common import HYPER, build_model, group_distances, oracle_update_on_gpu, well_conditioned_state
        from oracle import r3m_oracle as O
        from r3m_b200 import Trainer

        hyper = dict(HYPER, langweight=float(lang))
        params, buffers = well_conditioned_state(size, 70, bool(lang))       # identical on every rank
        frames = O.varied_frames(clips, 71 + rank)                            # the rank's shard
        perms = O.draw_permutations(clips, 81 + rank)                         # the rank's permutation stream
        lang_emb = O.stub_lang_embedding(clips, 91 + rank) if lang else None
        sentences = ["s%d" % i for i in range(clips)]
        mask = torch.ones(clips) if lang else None

        # (a) the rank's own single-GPU gradient: the engine without the collective
        m0, _ = build_model(size, params, buffers, float(lang), lang_emb)
        eng = m0._engine(clips * 5)
        eng.update_grads(frames.reshape(-1, 3, 224, 224).cuda().contiguous(), perms.to(torch.int32).cuda(),
                         lang_emb.cuda() if lang else None, mask.cuda() if lang else None, hyper["l2weight"],
                         hyper["l1weight"], float(lang), hyper["tcnweight"], False)
        local = m0._flat(1).clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        mean_local = torch.stack(gathered).mean(0)

        # (b) the distributed step: Trainer.update with torch.distributed initialised -> ONE all-reduce, Adam * 1/world
        m1, model1 = build_model(size, params, buffers, float(lang), lang_emb)
        before = m1._flat(0).clone()
        Trainer(100).update(model1, (frames.cuda(), sentences), 0, perms=perms, lang_emb=lang_emb)
        reduced = m1._flat(1).clone() / world                                 # region 1 holds the SUM after the step
        d_engine = float((reduced - mean_local).norm() / mean_local.norm())

        # (c) the oracle on each rank's shard, averaged over ranks
        _, o_grads, _, _, _ = oracle_update_on_gpu(size, params, buffers, frames, perms, hyper, lang_emb, mask)
        keys = sorted(k for k in o_grads if k.startswith("convnet."))
        flat = torch.cat([o_grads[k].flatten() for k in keys])
        og = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(og, flat.contiguous())
        o_mean = torch.stack(og).mean(0)
        named = dict(m1.named_parameters())
        ours = torch.cat([(named[k].grad / world).flatten() for k in keys])
        d_oracle = float((ours - o_mean).norm() / o_mean.norm())

        # every rank applied the same averaged gradient: bit-identical weights, and they moved
        w = m1._flat(0)
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w.contiguous())
        identical = all(torch.equal(ws[0], x) for x in ws[1:])
        moved = float((w - before).abs().max())
        out.put((rank, d_engine, d_oracle, identical, moved))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("size,clips,lang", [(18, 6, 1), (50, 8, 1)])
def test_allreduced_gradient_is_the_mean_of_the_shard_gradients(size, clips, lang):
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, size, clips, lang, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, d_engine, d_oracle, identical, moved in res:
        assert d_engine < 2e-2, (rank, d_engine)   # same kernels; only the BatchNorm-statistics atomics reorder
        assert d_oracle < 0.1, (rank, d_oracle)    # the bf16 tier's gradient tolerance on the well-conditioned state
        assert identical and 0.0 < moved < 2e-4
