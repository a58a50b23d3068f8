"""``Trainer`` — drop-in for r3m/trainer.py:21-162 on top of the sm_100a engine.

``update(model, batch, step, eval=False) -> (metrics, st)`` keeps the reference's contract: same ``metrics`` key set
(python floats), same ``st`` timing-string format, ``model`` is the ``nn.DataParallel``-wrapped ``R3M`` that
``train_representation.py:27-31`` builds.  Inside, one engine call runs forward + loss heads + backward, the 15
permutations are drawn from torch's global CPU generator in the reference's order (trainer.py:86-92,135-137) and
shipped to the device in one copy, and the metrics come back in one 64-byte read instead of ~10 ``.item()`` syncs.

Multi-GPU: one process per GPU (``torch.distributed`` initialised, backend nccl).  The step then contains exactly
one collective — an all-reduce (sum) of the flat fp32 gradient buffer over NVLink — and Adam applies 1/world.
"""
import time

import torch

from .engine import METRIC_KEYS


def _world():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


_COMM_STREAMS = {}


def allreduce_gradients(module, engine=None):
    """The step's ONLY collective: sum the flat fp32 gradient buffer over all ranks (NCCL over NVLink on GPUs; any
    torch.distributed backend works).  Returns the factor Adam must apply to turn the sum into the mean.

    With ``engine`` (the engine whose ``update_grads`` has just been enqueued) on a CUDA device the buffer is reduced
    in the engine's gradient chunks — language head + layer 4 first, stem + layer 1 last, the order in which the backward
    pass completes them — each on a communication stream that waits only for ITS chunk
    (``r3m_b200_engine_wait_grad_chunk``), so all but the last chunk travel while the backward pass is still running.
    Same result as one all-reduce of the whole buffer (the chunks partition it)."""
    world = _world()
    if world > 1:
        import torch.distributed as dist

        G = module._flat(1)
        if engine is None or not G.is_cuda:
            dist.all_reduce(G, op=dist.ReduceOp.SUM)
        else:
            dev = G.device
            comm = _COMM_STREAMS.get(dev)
            if comm is None:
                comm = _COMM_STREAMS[dev] = torch.cuda.Stream(device=dev)
            works = []
            for k, (b, e) in enumerate(engine.grad_chunks()):
                engine.wait_grad_chunk(k, comm)
                with torch.cuda.stream(comm):
                    works.append(dist.all_reduce(G[b:e], op=dist.ReduceOp.SUM, async_op=True))
            for w in works:
                w.wait()  # the current stream waits for the collective; the host does not block
    return 1.0 / world


def draw_permutations(batch_size, langweight, tcnweight, num_negatives=3):
    """[15, B] int32 permutations drawn exactly as the reference draws them (same count, same order, global CPU
    generator): 3*num_neg for the language branch if it is on, then 2*num_neg for TCN if it is on."""
    perms = torch.zeros(15, batch_size, dtype=torch.int64)
    if langweight > 0:
        for i in range(num_negatives * 3):
            perms[i] = torch.randperm(batch_size)
    if tcnweight > 0:
        for i in range(num_negatives * 2):
            perms[9 + i] = torch.randperm(batch_size)
    return perms.to(torch.int32)


class Trainer:
    def __init__(self, eval_freq):
        self.eval_freq = eval_freq
        self.last_launches = 0  # kernels launched by the most recent update() (forward + losses + backward + Adam)
        self._side_key = None  # pinned staging buffer of the step's small inputs (see _stage_small)
        self._side_host = self._side_dev = None

    def _stage_small(self, dev, perms, mask, lang_emb):
        """The step's small host-side inputs (15 permutations, sentence mask, a host-resident sentence embedding) go
        to the device through ONE pinned staging buffer that a kernel pulls over unified addressing
        (``r3m_b200_pull_host``).  Three ``.to(device)`` copies would each wait on the H2D copy engine behind whatever
        bulk upload the input pipeline has in flight (the next batch's 193 MB of frames: +3 ms per step, measured)."""
        from . import _lib as L

        def up16(n):
            return (n + 15) // 16 * 16

        emb_host = lang_emb is not None and not lang_emb.is_cuda
        n_perm = up16(perms.numel() * 4)
        n_mask = up16(mask.numel() * 4) if mask is not None else 0
        n_emb = up16(lang_emb.numel() * 4) if emb_host else 0
        total = n_perm + n_mask + n_emb
        key = (str(dev), total)
        if self._side_key != key:
            self._side_host = torch.empty(total, dtype=torch.uint8).pin_memory()
            self._side_dev = torch.empty(total, dtype=torch.uint8, device=dev)
            self._side_key = key
        host, devb = self._side_host, self._side_dev
        perms_dev = devb[:perms.numel() * 4].view(torch.int32).view(perms.shape)
        if not perms.is_cuda:
            host[:perms.numel() * 4].view(torch.int32).copy_(perms.reshape(-1).to(torch.int32))
        mask_dev = emb_dev = None
        if mask is not None:
            host[n_perm:n_perm + mask.numel() * 4].view(torch.float32).copy_(mask)
            mask_dev = devb[n_perm:n_perm + mask.numel() * 4].view(torch.float32)
        if emb_host:
            o = n_perm + n_mask
            host[o:o + lang_emb.numel() * 4].view(torch.float32).copy_(lang_emb.reshape(-1).to(torch.float32))
            emb_dev = devb[o:o + lang_emb.numel() * 4].view(torch.float32).view(lang_emb.shape)
        elif lang_emb is not None:
            # a device-resident embedding (the native sentence encoder's output) is copied into a buffer that keeps its
            # address from step to step: the captured step graph (engine: run_cached) is keyed on its input pointers
            key_e = (str(dev), tuple(lang_emb.shape))
            if getattr(self, "_emb_key", None) != key_e:
                self._emb_buf = torch.empty(lang_emb.shape, dtype=torch.float32, device=dev)
                self._emb_key = key_e
            self._emb_buf.copy_(lang_emb)
            emb_dev = self._emb_buf
        L.check(L.lib.r3m_b200_pull_host(host.data_ptr(), devb.data_ptr(), total, L.current_stream()))
        if perms.is_cuda:  # already resident (benchmarks / tests may inject device tensors): after the pull, same stream
            perms_dev.copy_(perms)
        return perms_dev, mask_dev, emb_dev

    def update(self, model, batch, step, eval=False, perms=None, lang_emb=None):
        """``perms`` / ``lang_emb`` are optional injection points for parity tests; by default they are produced the
        way the reference produces them."""
        t0 = time.time()
        metrics = dict()
        if eval:
            model.eval()
        else:
            model.train()
        m = model.module
        t1 = time.time()
        b_im, b_lang = batch
        t2 = time.time()

        bs = b_im.shape[0]
        frames = b_im.reshape(bs * 5, 3, 224, 224)  # trainer.py:39-40
        if frames.dtype not in (torch.float32, torch.uint8):  # uint8 frames go to the engine as they are
            frames = frames.float()
        if not frames.is_cuda:
            frames = frames.to(m._block.device, non_blocking=True)
        frames = frames.contiguous()
        eng = m._engine(bs * 5)
        dev = frames.device

        if not eval:
            # the forward pass needs only the frames: enqueue it first, so that drawing the permutations and staging the
            # small inputs below overlaps it instead of delaying the step's first kernel (0.26 ms per step, measured)
            eng.forward_train_async(frames)
        if perms is None:
            perms = draw_permutations(bs, m.langweight, m.tcnweight, m.num_negatives)
        t3 = time.time()
        mask = None
        if m.langweight > 0:
            if lang_emb is None:
                lang_emb = m.lang_enc(b_lang)  # identical for all 15 get_reward calls of the reference: run it once
            mask = torch.tensor([1.0 * (b != "") for b in b_lang], dtype=torch.float32)  # trainer.py:108
        perms_dev, mask_dev, emb_dev = self._stage_small(dev, perms, mask, lang_emb)
        eng.update_grads(frames if eval else None, perms_dev, emb_dev, mask_dev, float(m.l2weight), float(m.l1weight),
                         float(m.langweight), float(m.tcnweight), bool(eval))
        self.last_launches = eng.launches() + 1  # + the side-band pull
        t5 = time.time()
        t6 = time.time()
        if not eval:
            m._nbt += 1
            m.encoder_opt.step(grad_scale=allreduce_gradients(m, eng))
            self.last_launches += eng.launches()
        vals = eng.read_metrics()
        for i, k in enumerate(METRIC_KEYS):
            if k.startswith("rew") and not m.langweight > 0:
                continue
            if k in ("tcnloss", "aligned") and not m.tcnweight > 0:
                continue
            metrics[k] = vals[i]
        t7 = time.time()
        st = (f"Load time {t1-t0}, Batch time {t2-t1}, Encode and LP tine {t3-t2}, Lang time {t5-t3}, "
              f"TCN time {t6-t5}, Backprop time {t7-t6}")
        return metrics, st
