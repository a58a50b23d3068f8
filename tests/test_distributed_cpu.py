"""world_size-2 gloo test (CPU) of the multi-process host logic: one collective per step — a SUM all-reduce of the
flat gradient block — and the 1/world factor handed to Adam; rank-local permutation streams."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from r3m_b200 import R3M
        from r3m_b200.trainer import allreduce_gradients, draw_permutations

        torch.manual_seed(0)
        m = R3M("cpu", 1e-4, 1024, size=18, langweight=0.0, tcnweight=1.0)
        first = float(m.convnet.conv1.weight.detach().flatten()[0])
        g = m._flat(1)
        g.fill_(float(rank + 1))
        w = m.convnet.layer2._modules["0"].conv1.weight
        w.grad[0, 0, 0, 0] = 10.0 * (rank + 1)  # through the OIHW view
        scale = allreduce_gradients(m)
        torch.manual_seed(100 + rank)
        perms = draw_permutations(8, 0.0, 1.0)
        out.put((rank, scale, float(g.min()), float(g.max()), float(w.grad[0, 0, 0, 0]), first,
                 perms[9].tolist()))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, scale, gmin, gmax, wg, first, perm in res:
        assert scale == 0.5                       # Adam averages the summed gradient
        assert gmin == 3.0 and gmax == 30.0 and wg == 30.0   # 1 + 2 everywhere, 10 + 20 at the marked element
    assert res[0][5] == res[1][5]                 # identical initial weights on every rank (same seed)
    assert res[0][6] != res[1][6]                 # rank-local negatives (per-rank permutation stream)
