"""Input pipeline on the GPU: uint8 frames consumed directly by the stem kernel, RandomResizedCrop's
arithmetic against torchvision's op on identical boxes, and the pinned double-buffered FrameFeeder."""
import pytest
import torch

from gpu_common import build_model, rel
from oracle import r3m_oracle as O

pytestmark = pytest.mark.gpu


def test_uint8_frames_give_the_same_stem_operand_and_embeddings(lib):
    """obs.float() of models_r3m.py:97 fused into the kernel: uint8 NCHW and uint8 NHWC frames produce bit-identical
    stem operands to the fp32 frames holding the same integers."""
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (7, 3, 224, 224), generator=g, dtype=torch.uint8).cuda()
    xs = [torch.empty(7, 112, 112, 64, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    srcs = [(u8.float().contiguous(), 0), (u8, 1), (u8.permute(0, 2, 3, 1).contiguous(), 2)]
    for (src, fmt), dst in zip(srcs, xs):
        lib.check(lib.lib.r3m_b200_preprocess_stem_format(lib.ptr(src), fmt, lib.ptr(dst), 7, lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(xs[0], xs[1]) and torch.equal(xs[0], xs[2])
    # end to end: R3M.forward on uint8 == on float (eval mode: deterministic kernels)
    params, buffers = O.eval_fixture_state(18)
    m, _ = build_model(18, params, buffers, 0.0)
    m.eval()
    with torch.no_grad():
        a, b = m(u8.float()), m(u8)
        c = m._engine(7).forward(u8.permute(0, 2, 3, 1).contiguous(), False, nhwc=True)
    assert torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.parametrize("H,W", [(224, 224), (240, 320), (480, 360), (1080, 1920), (100, 90)])
def test_random_resized_crop_matches_torchvision_on_identical_boxes(H, W):
    """boxes drawn with the reference's law (bit-exact, tests/test_data_cpu.py); pixels of the antialiased bilinear
    resize within 1e-3 (on the [0, 255] scale) of torchvision.transforms.functional.resized_crop — up-scaling,
    down-scaling by > 8x, and boxes touching the frame border."""
    TF = pytest.importorskip("torchvision.transforms.functional")
    from r3m_b200.data import draw_crop_boxes, random_resized_crop

    g = torch.Generator().manual_seed(H + W)
    n = 10
    frames = torch.randint(0, 256, (n, 3, H, W), generator=g, dtype=torch.uint8)
    torch.manual_seed(H * 7 + W)
    boxes = draw_crop_boxes(2, H, W, "rc")
    boxes[0] = torch.tensor([0, 0, H, W])              # the whole frame
    boxes[1] = torch.tensor([H - min(H, 17), W - min(W, 23), min(H, 17), min(W, 23)])  # a tiny corner box: up-scaling
    got = random_resized_crop(frames.cuda(), boxes)
    got_nhwc = random_resized_crop(frames.permute(0, 2, 3, 1).contiguous().cuda(), boxes, nhwc=True)
    assert got.shape == (n, 3, 224, 224) and got.dtype == torch.float32
    assert torch.equal(got, got_nhwc)
    worst = 0.0
    for i in range(n):
        t, l, h, w = boxes[i].tolist()
        want = TF.resized_crop(frames[i].float() / 255.0, t, l, h, w, [224, 224], antialias=True) * 255.0
        worst = max(worst, float((got[i].cpu() - want).abs().max()))
    assert worst < 1e-3, worst
    assert float(got.min()) >= 0.0 and float(got.max()) <= 255.0 + 1e-3


def test_frame_feeder_delivers_batches_in_order_overlapped():
    from r3m_b200 import FrameFeeder, GpuAugment

    g = torch.Generator().manual_seed(1)
    batches = [(torch.randint(0, 255, (3, 5, 3, 224, 224), generator=g).float(), [f"s{i}"] * 3) for i in range(5)]
    feeder = FrameFeeder(iter(batches), "cuda")
    seen = 0
    for i, (frames, labels) in enumerate(feeder):
        assert frames.is_cuda and frames.dtype == torch.uint8 and labels == [f"s{i}"] * 3
        # consume on the current stream, slowly enough that the next upload overlaps
        torch.randn(2048, 2048, device="cuda").mm(torch.randn(2048, 2048, device="cuda"))
        assert int(frames.long().sum()) == int(batches[i][0].long().sum())
        seen += 1
    assert seen == 5 and feeder.bytes_per_batch == 3 * 5 * 3 * 224 * 224
    # pinned uint8 sources are uploaded without restaging; float sources can be kept as float
    pinned = torch.randint(0, 256, (2, 5, 3, 224, 224), dtype=torch.uint8).pin_memory()
    frames, _ = next(FrameFeeder(iter([(pinned, ["a", "b"])]), "cuda"))
    torch.cuda.synchronize()
    assert torch.equal(frames.cpu(), pinned)
    frames, _ = next(FrameFeeder(iter([batches[0]]), "cuda", as_uint8=False))
    assert frames.dtype == torch.float32 and torch.equal(frames.cpu(), batches[0][0])
    # with the GPU augmentation: uint8 source frames of another size -> float [B,5,3,224,224]
    src = torch.randint(0, 256, (2, 5, 3, 240, 320), dtype=torch.uint8)
    torch.manual_seed(3)
    frames, _ = next(FrameFeeder(iter([(src, ["a", "b"])]), "cuda", augment=GpuAugment("rctraj")))
    assert frames.shape == (2, 5, 3, 224, 224) and frames.dtype == torch.float32


def test_trainer_update_accepts_uint8_frames_from_the_feeder():
    from r3m_b200 import FrameFeeder, Trainer

    params, buffers = O.init_state(18, 3)
    frames = O.synthetic_frames(4, 5)
    perms = O.draw_permutations(4, 6)
    out = []
    for src in (frames, frames.to(torch.uint8)):
        m, model = build_model(18, params, buffers, 0.0)
        batch, _ = next(FrameFeeder(iter([(src, [""] * 4)]), "cuda", as_uint8=False))
        metrics, _ = Trainer(100).update(model, (batch, [""] * 4), 0, perms=perms)
        out.append((metrics, m._any_engine().embeddings().clone()))
    assert rel(out[1][1], out[0][1]) < 2e-2  # same operands; the BatchNorm-statistics atomics reorder run to run
    assert abs(out[0][0]["l2loss"] - out[1][0]["l2loss"]) < 1e-2 * out[0][0]["l2loss"]


def test_nvjpeg_batch_decoder_feeds_the_uint8_path(tmp_path):
    """JPEG decode on the GPU (nvJPEG via torchvision) behind R3MBufferU8 — same clip-index law and
    labels as the CPU decoder, pixels within the decoders' IDCT round-off, frames born on the device as uint8."""
    import numpy as np
    import pandas as pd
    import random
    import torchvision

    from r3m_b200.data import R3MBufferU8, nvjpeg_batch_decoder

    g = torch.Generator().manual_seed(3)
    vid = tmp_path / "vid0"
    vid.mkdir()
    yy, xx = torch.meshgrid(torch.arange(120.0), torch.arange(160.0), indexing="ij")
    for i in range(12):
        # smooth content (no hard chroma edges: the two decoders upsample 4:2:0 chroma differently) plus mild noise
        img = torch.stack([128 + 90 * torch.sin(yy / (17.0 + 5 * c) + 0.3 * i) * torch.cos(xx / (23.0 - 4 * c))
                           for c in range(3)])
        img = (img + 4 * torch.rand(3, 120, 160, generator=g)).clamp(0, 255).to(torch.uint8)
        torchvision.io.write_jpeg(img, str(vid / f"{i:06}.jpg"), quality=92)
    manifest = pd.DataFrame({"path": [str(vid)], "len": [12], "txt": ["C opens the drawer"]})

    def sample(**kw):
        random.seed(5)
        np.random.seed(5)
        return R3MBufferU8("", 0.2, manifest=manifest, **kw)._sample()

    cpu_im, cpu_label = sample()
    gpu_im, gpu_label = sample(batch_decoder=nvjpeg_batch_decoder("cuda"))
    assert gpu_label == cpu_label == "opens the drawer"
    assert gpu_im.is_cuda and gpu_im.dtype == torch.uint8 and tuple(gpu_im.shape) == (5, 3, 120, 160)
    diff = (gpu_im.cpu().int() - cpu_im.int()).abs().float()
    # nvJPEG and libjpeg-turbo differ in IDCT rounding and chroma upsampling: a couple of grey levels, never a wrong frame
    assert diff.mean() < 1.5 and float(diff.flatten().kthvalue(int(0.999 * diff.numel())).values) <= 12, \
        (float(diff.mean()), float(diff.max()))
    other = torch.roll(cpu_im, 1, 0).int()  # a DIFFERENT frame of the clip is far away: the check can fail
    assert (other - cpu_im.int()).abs().float().mean() > 5 * max(float(diff.mean()), 0.2)
