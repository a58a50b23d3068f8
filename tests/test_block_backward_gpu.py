"""Block-level backward WIRING test: after a train-mode forward, a random gradient is written into
the buffer one residual block's backward reads, ONLY that block's backward kernels are run (BatchNorm backward, dgrad,
wgrad, with their accumulate flags, the dual-BatchNorm tail of downsample blocks, the dy ring), and dX, dW, dgamma,
dbeta are compared with torch autograd of the same block (tv resnet.py:89-105 BasicBlock, :143-163 Bottleneck)
evaluated in fp32 on the SAME bf16-rounded block input, rounding where the CUDA path stores bf16.

Tolerance 3e-2 (relative L2).  Measured on the B200 (gpurun_out/r2_block_backward.jsonl -> profiles/): 1.2e-3 .. 2.5e-3 in
layer1, 4e-3 .. 7e-3 in layer2, ~1e-2 in layer3, 1.4e-2 .. 2.3e-2 in layer4 — the deeper the block, the longer its
fp32 reductions (K up to 4608), the more bf16 roundings of an intermediate land on the other side in the two
implementations, and each flipped ReLU-mask bit moves the random-walk sums over the injected gradient (even dbeta of the
LAST BatchNorm, which depends on nothing but the injected gradient and the mask, moves by 1.4e-2 there).  A wiring
error is two orders of magnitude away: dropping the skip-path accumulation gives dX = 0.96 (test_mutation_is_caught)."""
import pytest
import torch
import torch.nn.functional as F

from gpu_common import rel, strict_fp32, well_conditioned_state
from oracle import r3m_oracle as O

pytestmark = pytest.mark.gpu
q = O._RoundBf16.apply
qw = O._RoundBf16Forward.apply


def _block_specs(size):
    """[(prefix, kind, stride, has_ds)] in forward order."""
    kind, layers = O._CFG[size]
    out, inplanes = [], 64
    exp = 1 if kind == "basic" else 4
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
        for b in range(n):
            stride = 2 if (li > 0 and b == 0) else 1
            out.append((f"layer{li + 1}.{b}", kind, stride, b == 0 and (stride != 1 or inplanes != planes * exp)))
            inplanes = planes * exp
    return out


def _bn(x, w, b):
    mean = x.mean((0, 2, 3), keepdim=True)
    var = x.var((0, 2, 3), unbiased=False, keepdim=True)
    return (x - mean) * torch.rsqrt(var + O.BN_EPS) * w[None, :, None, None] + b[None, :, None, None]


def _block_forward(x, P, pre, kind, stride, has_ds):
    conv = lambda t, n, s, p: q(F.conv2d(t, qw(P[f"convnet.{pre}.{n}.weight"]), None, s, p))  # noqa: E731
    bn = lambda t, n: _bn(t, P[f"convnet.{pre}.{n}.weight"], P[f"convnet.{pre}.{n}.bias"])  # noqa: E731
    if kind == "basic":
        out = q(F.relu(bn(conv(x, "conv1", stride, 1), "bn1")))
        out = bn(conv(out, "conv2", 1, 1), "bn2")
    else:
        out = q(F.relu(bn(conv(x, "conv1", 1, 0), "bn1")))
        out = q(F.relu(bn(conv(out, "conv2", stride, 1), "bn2")))
        out = bn(conv(out, "conv3", 1, 0), "bn3")
    identity = x
    if has_ds:
        identity = bn(conv(x, "downsample.0", stride, 0), "downsample.1")
    return F.relu(out + identity)


def check_block(size, block, frames=6, seed=0, tol=3e-2):
    """Returns {name: relative error}; raises AssertionError when any exceeds `tol`."""
    from r3m_b200 import R3M

    strict_fp32()
    params, buffers = well_conditioned_state(size, 40 + seed, False)
    m = R3M("cuda", 1e-4, 1024, size=size, langweight=0.0)
    sd = dict(params)
    sd.update(buffers)
    m.load_state_dict(sd)
    m = m.cuda().train()
    x = O.varied_frames(2, 41 + seed).reshape(-1, 3, 224, 224)[:frames]
    with torch.no_grad():
        m(x.cuda())
    eng = m._engine(frames)
    pre, kind, stride, has_ds = _block_specs(size)[block]
    named = dict(m.named_parameters())
    keys = [k for k in named if k.startswith(f"convnet.{pre}.")]
    P = {k: named[k].detach().clone().float().contiguous().requires_grad_(True) for k in keys}
    x_in = eng.block_buffer(block, 0)
    cin = named[f"convnet.{pre}.conv1.weight"].shape[1]
    hw = int(round((x_in.numel() / (frames * cin)) ** 0.5))
    xr = x_in.view(frames, hw, hw, cin).permute(0, 3, 1, 2).float().contiguous().requires_grad_(True)
    out = _block_forward(xr, P, pre, kind, stride, has_ds)
    a_out = eng.block_buffer(block, 1).view(frames, out.shape[2], out.shape[3], out.shape[1]).permute(0, 3, 1, 2).float()
    errs = {"forward": rel(a_out, out)}
    g = torch.Generator().manual_seed(7 + seed)
    dA = (torch.randn(out.shape, generator=g) * 0.05).bfloat16().cuda()
    eng.block_buffer(block, 2).copy_(dA.permute(0, 2, 3, 1).reshape(-1))
    eng.block_buffer(block, 3).fill_(float("nan"))  # must be fully overwritten
    m._flat(1).zero_()
    eng.run_block_backward(block)
    torch.cuda.synchronize()
    from r3m_b200 import _lib as L

    L.check(L.lib.r3m_b200_check_device_flag())
    out.backward(dA.float())
    d_in = eng.block_buffer(block, 3).view(frames, hw, hw, cin).permute(0, 3, 1, 2).float()
    errs["dX"] = rel(d_in, xr.grad)
    for k in keys:
        errs[k[len("convnet."):]] = rel(named[k].grad, P[k].grad)
    import json
    import os

    rep = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(rep):
        with open(os.path.join(rep, "r2_block_backward.jsonl"), "a") as f:
            f.write(json.dumps({"size": size, "block": pre, "frames": frames, "seed": seed,
                                "mutation": os.environ.get("R3M_TEST_MUTATION", ""), "errors": errs}) + "\n")
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, (size, block, bad, errs)
    return errs


@pytest.mark.parametrize("size,block", [(50, 0), (50, 1), (50, 3), (50, 8), (50, 15), (18, 0), (18, 2), (18, 7),
                                        (34, 7), (34, 15)])
def test_block_backward_matches_autograd(size, block):
    """RN50: layer1.0 (stride-1 downsample), layer1.1 (identity), layer2.0 (stride-2 downsample + parity-class dgrad),
    layer3.1, layer4.2; RN18/34 BasicBlocks: identity and downsample kinds."""
    errs = check_block(size, block)
    assert set(errs) >= {"forward", "dX"} and len(errs) >= 8


def test_mutation_is_caught(monkeypatch):
    """The check above must be able to FAIL: with the skip-path accumulation of identity blocks deliberately dropped
    (R3M_TEST_MUTATION=drop_skip_add, read when an engine is planned) the identity block's dX is wrong by O(1)."""
    monkeypatch.setenv("R3M_TEST_MUTATION", "drop_skip_add")
    with pytest.raises(AssertionError) as info:
        check_block(50, 1, seed=1)
    assert "dX" in str(info.value)
    monkeypatch.delenv("R3M_TEST_MUTATION")
    check_block(50, 1, seed=1)
