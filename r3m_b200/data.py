"""Input pipeline of the pretraining step — the loader side of r3m/utils/data_loaders.py and the
``batch_f.cuda()`` of r3m/train_representation.py:104, rebuilt so that it can feed an engine that consumes >13 k
frames/s per GPU:

  * frames stay **uint8** from the decoder to the stem kernel (a quarter of the PCIe / HBM bytes of the reference's
    fp32 ``[B,5,3,224,224]`` batches; the ``/255`` and ``Normalize`` happen in registers);
  * ``RandomResizedCrop(224, scale=(0.2, 1.0))`` (``doaug`` "rc" / "rctraj", data_loaders.py:47-50,81-102) runs on the
    GPU (``r3m_b200_random_resized_crop``); only the crop BOXES are drawn on the host, with torchvision's law and
    torch's CPU generator, so a seeded run crops exactly where the reference would;
  * ``FrameFeeder`` overlaps the host->device copy of batch i+1 with the step on batch i (pinned staging, a copy
    stream, two slots);
  * the clip-index sampling law of ``R3MBuffer._sample`` (data_loaders.py:64-79) is kept verbatim in
    ``sample_clip_indices``.

The contract towards ``Trainer.update`` is the reference's: ``(im [B,5,3,224,224] in [0,255], labels)``.
"""
import math
import random

import numpy as np
import torch

from . import _lib as L


# ---------------------------------------------------------------------------------------------------------------
# sampling laws (host)
# ---------------------------------------------------------------------------------------------------------------
def sample_clip_indices(vidlen, alpha, rng=np.random):
    """data_loaders.py:75-79: (start, end, s0, s1, s2) frame indices of one clip, same draws in the same order."""
    start_ind = rng.randint(1, 2 + int(alpha * vidlen))
    end_ind = rng.randint(int((1 - alpha) * vidlen) - 1, vidlen)
    s1_ind = rng.randint(2, vidlen)
    s0_ind = rng.randint(1, s1_ind)
    s2_ind = rng.randint(s1_ind, vidlen + 1)
    return start_ind, end_ind, s0_ind, s1_ind, s2_ind


def random_resized_crop_params(height, width, scale=(0.2, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), generator=None):
    """torchvision.transforms.RandomResizedCrop.get_params: up to ten (area, log-uniform aspect) draws from torch's CPU
    generator, the first box that fits wins, else the central crop at the clamped aspect.  Returns (top, left, h, w).
    Consumes the generator exactly like torchvision, so seeded runs produce identical boxes."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1], generator=generator).item()
        aspect = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1], generator=generator)).item()
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,), generator=generator).item()
            j = torch.randint(0, width - w + 1, size=(1,), generator=generator).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def draw_crop_boxes(num_clips, height, width, doaug, generator=None):
    """int32 [num_clips*5, 4] crop boxes in the reference's draw order: "rc" draws one box per frame (im0, img, imts0,
    imts1, imts2 of each clip, data_loaders.py:97-101), "rctraj" one box per clip shared by its five frames (:81-95)."""
    boxes = []
    for _ in range(num_clips):
        if doaug == "rctraj":
            boxes += [random_resized_crop_params(height, width, generator=generator)] * 5
        elif doaug == "rc":
            boxes += [random_resized_crop_params(height, width, generator=generator) for _ in range(5)]
        else:
            raise ValueError("crop boxes are drawn for doaug in ('rc', 'rctraj') only")
    return torch.tensor(boxes, dtype=torch.int32)


# ---------------------------------------------------------------------------------------------------------------
# GPU augmentation
# ---------------------------------------------------------------------------------------------------------------
def random_resized_crop(frames_u8, boxes, nhwc=False):
    """frames_u8: CUDA uint8 [N,3,H,W] ([N,H,W,3] with nhwc); boxes: int32 [N,4] (top, left, h, w), host or device.
    -> CUDA float32 [N,3,224,224] in [0,255]: what ``self.aug(frame / 255.0) * 255.0`` yields in the reference."""
    assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.is_contiguous()
    n = frames_u8.shape[0]
    h, w = (frames_u8.shape[1], frames_u8.shape[2]) if nhwc else (frames_u8.shape[2], frames_u8.shape[3])
    boxes = torch.as_tensor(boxes, dtype=torch.int32)
    assert boxes.shape == (n, 4)
    b = boxes.cpu()
    if bool(((b[:, 0] < 0) | (b[:, 1] < 0) | (b[:, 2] < 1) | (b[:, 3] < 1) | (b[:, 0] + b[:, 2] > h)
             | (b[:, 1] + b[:, 3] > w)).any()):
        raise ValueError("crop box outside the frame")
    boxes = boxes.to(frames_u8.device).contiguous()
    out = torch.empty(n, 3, 224, 224, dtype=torch.float32, device=frames_u8.device)
    with torch.cuda.device(frames_u8.device):
        L.check(L.lib.r3m_b200_random_resized_crop(L.ptr(frames_u8), int(nhwc), n, h, w, L.ptr(boxes), L.ptr(out),
                                                   L.current_stream()))
    return out


class GpuAugment:
    """``R3MBuffer.aug`` moved to the device: callable on a uint8 batch [B,5,3,H,W] -> float32 [B,5,3,224,224]
    ("rc"/"rctraj"), or the identity for ``doaug="none"`` (frames must then be 224x224 already and stay uint8)."""

    def __init__(self, doaug="none", generator=None):
        if doaug not in ("none", "rc", "rctraj"):
            raise ValueError(doaug)
        self.doaug, self.generator = doaug, generator

    def __call__(self, batch_u8):
        if self.doaug == "none":
            return batch_u8
        b, five, c, h, w = batch_u8.shape
        boxes = draw_crop_boxes(b, h, w, self.doaug, self.generator)
        return random_resized_crop(batch_u8.reshape(b * five, c, h, w), boxes).reshape(b, five, 3, 224, 224)


# ---------------------------------------------------------------------------------------------------------------
# host -> device feeder
# ---------------------------------------------------------------------------------------------------------------
class FrameFeeder:
    """Wraps an iterable of ``(frames, labels)`` batches (what ``DataLoader(R3MBuffer)`` yields) and delivers them on
    ``device``: every batch is staged in one of ``depth`` pinned host buffers and copied on a dedicated copy stream
    while the previous batch is being consumed, so the step never waits for PCIe (replaces the synchronous
    ``batch_f.cuda()`` of train_representation.py:104).  ``as_uint8`` converts float frames holding integers (the
    reference loader's output without augmentation) to uint8 on the host before the copy: 4x fewer bytes.

    The tensor handed out stays valid until the NEXT batch is requested; the consumer's work must be enqueued on the
    current stream by then (it is: ``Trainer.update`` enqueues everything before returning)."""

    def __init__(self, batches, device, depth=2, as_uint8=True, augment=None):
        self.device = torch.device(device)
        self.depth, self.as_uint8, self.augment = max(2, depth), as_uint8, augment
        self._it = iter(batches)
        self._copy = torch.cuda.Stream(device=self.device)
        self._slots = [None] * self.depth      # (pinned host buffer, device buffer)
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._consumed = [torch.cuda.Event() for _ in range(self.depth)]
        self._labels = [None] * self.depth
        self._head = self._count = 0
        self.bytes_per_batch = 0
        for ev in self._consumed:
            ev.record(torch.cuda.current_stream(self.device))
        self._prefetch()

    def _prefetch(self):
        try:
            frames, labels = next(self._it)
        except StopIteration:
            return False
        slot = (self._head + self._count) % self.depth
        frames = torch.as_tensor(frames)
        if self.as_uint8 and frames.dtype != torch.uint8:
            frames = frames.to(torch.uint8)  # the loader's floats are integers in [0,255] (decoded JPEG bytes)
        buf = self._slots[slot]
        if buf is None or buf[0].shape != frames.shape or buf[0].dtype != frames.dtype:
            buf = (torch.empty(frames.shape, dtype=frames.dtype).pin_memory(),
                   torch.empty(frames.shape, dtype=frames.dtype, device=self.device))
            self._slots[slot] = buf
        src = frames
        if not frames.is_pinned():
            self._ready[slot].synchronize()  # the slot's previous upload has left the pinned staging buffer
            buf[0].copy_(frames)
            src = buf[0]
        self._copy.wait_event(self._consumed[slot])  # the device buffer's last reader has finished
        with torch.cuda.stream(self._copy):
            buf[1].copy_(src, non_blocking=True)
            self._ready[slot].record(self._copy)
        self._keep = src  # a pinned source must outlive its asynchronous copy
        self._labels[slot] = labels
        self.bytes_per_batch = frames.numel() * frames.element_size()
        self._count += 1
        return True

    def __iter__(self):
        return self

    def __next__(self):
        if self._count == 0:
            raise StopIteration
        slot = self._head
        cur = torch.cuda.current_stream(self.device)
        prev = (slot - 1) % self.depth
        self._consumed[prev].record(cur)  # everything enqueued so far has read the previous batch
        cur.wait_event(self._ready[slot])
        self._head = (self._head + 1) % self.depth
        self._count -= 1
        self._prefetch()  # the next batch's copy overlaps the work the caller is about to enqueue
        frames = self._slots[slot][1]
        if self.augment is not None:
            frames = self.augment(frames)
        return frames, self._labels[slot]


# ---------------------------------------------------------------------------------------------------------------
# JPEG decode on the GPU (nvJPEG through torchvision.io.decode_jpeg: a library call, like the reference's
# torchvision.io.read_image on the CPU, data_loaders.py:30-32) — the frames are born in device memory as uint8 and go
# to GpuAugment / the stem kernel without crossing PCIe decoded
# ---------------------------------------------------------------------------------------------------------------
def nvjpeg_batch_decoder(device="cuda"):
    """``batch_decoder`` for R3MBufferU8: the clip's JPEG files are read on the host (compressed bytes only) and decoded
    together by nvJPEG on ``device`` -> uint8 [n,3,H,W] there.  Use in the training process (CUDA in DataLoader workers
    needs the spawn start method)."""
    import torchvision
    from torchvision.io import ImageReadMode

    dev = torch.device(device)

    def decode(paths):
        data = [torchvision.io.read_file(p) for p in paths]
        frames = torchvision.io.decode_jpeg(data, mode=ImageReadMode.RGB, device=dev)
        return torch.stack(frames)

    return decode


# ---------------------------------------------------------------------------------------------------------------
# dataset (host)
# ---------------------------------------------------------------------------------------------------------------
class R3MBufferU8(torch.utils.data.IterableDataset):
    """``R3MBuffer`` (data_loaders.py:38-105) that keeps what it decodes as uint8 and leaves the augmentation to the GPU:
    yields ``(uint8 [5,3,H,W], label)`` with the reference's manifest columns (path, len, txt), index law and label
    rule (``txt[2:]``).  ``decoder(path) -> uint8 [3,H,W]`` defaults to ``torchvision.io.read_image``; ``batch_decoder``
    (``nvjpeg_batch_decoder``) decodes the clip's five files in one nvJPEG call on the GPU instead."""

    def __init__(self, ego4dpath, alpha, datasources=("ego4d",), manifest=None, decoder=None, batch_decoder=None):
        super().__init__()
        self._decode_batch = batch_decoder  # optional: list of paths -> uint8 [n,3,H,W] (e.g. nvjpeg_batch_decoder)
        if "ego4d" not in datasources:
            raise NameError("Invalid Dataset")  # data_loaders.py:61
        self.alpha, self.data_sources = alpha, list(datasources)
        if manifest is None:
            import pandas as pd

            manifest = pd.read_csv(f"{ego4dpath}manifest.csv")
        self.manifest = manifest
        self.ego4dlen = len(manifest)
        if decoder is None:
            import torchvision

            decoder = torchvision.io.read_image
        self._decode = decoder

    def _sample(self):
        random.choice(self.data_sources)
        vidid = np.random.randint(0, self.ego4dlen)
        m = self.manifest.iloc[vidid]
        vidlen, txt, vid = m["len"], m["txt"], m["path"]
        label = txt[2:]
        inds = sample_clip_indices(vidlen, self.alpha)
        paths = [f"{vid}/{i:06}.jpg" for i in inds]
        if self._decode_batch is not None:
            im = self._decode_batch(paths)
        else:
            im = torch.stack([self._decode(p) for p in paths])
        return im, label

    def __iter__(self):
        while True:
            yield self._sample()
