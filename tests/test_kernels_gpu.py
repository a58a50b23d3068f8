"""Kernel-level parity on the B200, through the C ABI, against torch fp32 ops on the SAME bf16-rounded operands
(floating-point kernels: the torch fp32 reference of the same op is the checker; tolerances are the bf16 output
rounding 2^-9 / sqrt(3) ~ 1.1e-3 relative L2 for bf16 outputs, fp32 round-off for fp32 outputs)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GEOMS = [  # N, H, Cin, Cout, R, stride, pad — every conv shape class of ResNet-18/34/50
    (2, 56, 64, 64, 1, 1, 0), (2, 56, 64, 64, 3, 1, 1), (3, 28, 128, 128, 3, 1, 1), (2, 14, 256, 256, 3, 1, 1),
    (3, 7, 512, 512, 3, 1, 1), (4, 14, 256, 1024, 1, 1, 0), (2, 56, 128, 128, 3, 2, 1), (2, 56, 256, 512, 1, 2, 0),
    (5, 14, 1024, 256, 1, 1, 0), (3, 28, 64, 128, 3, 2, 1), (1, 7, 2048, 512, 1, 1, 0), (2, 56, 64, 256, 1, 1, 0),
]


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(autouse=True)
def _exact_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _mk(geom, seed=0):
    N, H, Cin, Cout, R, stride, pad = geom
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, H, generator=g).cuda().bfloat16()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).cuda().bfloat16()
    P = (H + 2 * pad - R) // stride + 1
    dy = torch.randn(N, Cout, P, P, generator=g).cuda().bfloat16()
    return x, w, dy, P


@pytest.mark.parametrize("geom", GEOMS)
def test_conv_forward_and_bn_statistics(lib, geom):
    N, H, Cin, Cout, R, stride, pad = geom
    x, w, _, P = _mk(geom)
    xn, wk = x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous()
    y = torch.full((N, P, P, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    ssum, ssq = torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    lib.check(lib.lib.r3m_b200_conv_fwd(lib.ptr(xn), lib.ptr(wk), lib.ptr(y), N, H, H, Cin, Cout, R, R, stride, pad,
                                        lib.ptr(ssum), lib.ptr(ssq), lib.current_stream()))
    lib.check(lib.lib.r3m_b200_check_device_flag())
    ref = F.conv2d(x.float(), w.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    assert rel(y.float(), ref) < 2.5e-3
    yf = y.float()
    assert rel(ssum, yf.sum((0, 1, 2))) < 1e-4 and rel(ssq, (yf * yf).sum((0, 1, 2))) < 1e-5


@pytest.mark.parametrize("geom", GEOMS)
def test_conv_dgrad(lib, geom):
    N, H, Cin, Cout, R, stride, pad = geom
    x, w, dy, P = _mk(geom, 1)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    wm = w.float().permute(0, 2, 3, 1).contiguous()
    wd = torch.empty(Cout * R * R * Cin, device="cuda", dtype=torch.bfloat16)
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_pack_dgrad_filter(lib.ptr(wm), lib.ptr(wd), Cout, R, R, Cin, stride, pad, s))
    dx = torch.full((N, H, H, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.check(lib.lib.r3m_b200_conv_dgrad(lib.ptr(dyn), lib.ptr(wd), lib.ptr(dx), N, H, H, Cin, Cout, R, R, stride,
                                          pad, 0, s))
    ref = torch.nn.grad.conv2d_input(x.shape, w.float(), dy.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    assert rel(dx.float(), ref) < 2.5e-3
    base = torch.randn(N, H, H, Cin, device="cuda").bfloat16()
    acc = base.clone()
    lib.check(lib.lib.r3m_b200_conv_dgrad(lib.ptr(dyn), lib.ptr(wd), lib.ptr(acc), N, H, H, Cin, Cout, R, R, stride,
                                          pad, 1, s))
    lib.check(lib.lib.r3m_b200_check_device_flag())
    assert rel(acc.float(), ref + base.float()) < 4e-3


@pytest.mark.parametrize("geom", GEOMS)
def test_conv_wgrad(lib, geom):
    N, H, Cin, Cout, R, stride, pad = geom
    x, w, dy, P = _mk(geom, 2)
    xn, dyn = x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous()
    dw = torch.zeros(Cout, R, R, Cin, device="cuda")
    lib.check(lib.lib.r3m_b200_conv_wgrad(lib.ptr(dyn), lib.ptr(xn), lib.ptr(dw), N, H, H, Cin, Cout, R, R, stride,
                                          pad, lib.current_stream()))
    lib.check(lib.lib.r3m_b200_check_device_flag())
    ref = torch.nn.grad.conv2d_weight(x.float(), w.shape, dy.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    assert rel(dw, ref) < 2e-5


@pytest.mark.parametrize("N,H,W", [(3, 56, 56), (2, 8, 8), (5, 12, 16), (2, 20, 8), (4, 16, 24)])
def test_conv3x3_tap_reuse_kernel(lib, N, H, W):
    """halo3x3_kernel (3x3, stride 1, 64 -> 64 channels): 16 x 8 output patches, ONE tiled TMA load of the 18 x 10 halo, the
    nine taps as shifted descriptor views of that shared memory, resident filter.  Forward (no statistics: the C-ABI
    statistics outputs are fp32 arrays, the kernel's are the engine's raw accumulators — covered by the block / full-size
    engine tests) and data gradient, including images whose height is not a multiple of the 16-row patch."""
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(N, 64, H, W, generator=g).cuda().bfloat16()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24.0).cuda().bfloat16()
    dy = torch.randn(N, 64, H, W, generator=g).cuda().bfloat16()
    xn, wk = x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous()
    y = torch.full((N, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_conv_fwd(lib.ptr(xn), lib.ptr(wk), lib.ptr(y), N, H, W, 64, 64, 3, 3, 1, 1, None, None, s))
    lib.check(lib.lib.r3m_b200_check_device_flag())
    ref = F.conv2d(x.float(), w.float(), padding=1).permute(0, 2, 3, 1)
    assert rel(y.float(), ref) < 2.5e-3, rel(y.float(), ref)
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    wm = w.float().permute(0, 2, 3, 1).contiguous()
    wd = torch.empty(64 * 9 * 64, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.lib.r3m_b200_pack_dgrad_filter(lib.ptr(wm), lib.ptr(wd), 64, 3, 3, 64, 1, 1, s))
    dx = torch.full((N, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.check(lib.lib.r3m_b200_conv_dgrad(lib.ptr(dyn), lib.ptr(wd), lib.ptr(dx), N, H, W, 64, 64, 3, 3, 1, 1, 0, s))
    lib.check(lib.lib.r3m_b200_check_device_flag())
    refd = torch.nn.grad.conv2d_input(x.shape, w.float(), dy.float(), padding=1).permute(0, 2, 3, 1)
    assert rel(dx.float(), refd) < 2.5e-3, rel(dx.float(), refd)


PAIR_GEOMS = [  # filter gradients that run on CTA pairs (cta_group::2): Cout % 256 == 0, an even number of (tap, 64-ch) items
    (20, 14, 256, 256, 3, 1, 1),   # 36 items -> groups of 8 (last one 4), 31 pixel blocks: several pipeline rounds
    (16, 28, 128, 512, 1, 1, 0),   # 2 items: one per CTA, N = 128 instructions
    (9, 28, 512, 1024, 1, 2, 0),   # stride-2 downsample, 8 items
    (7, 28, 256, 256, 3, 2, 1),    # stride-2 3x3
    (11, 7, 512, 2048, 1, 1, 0),   # 16 k-tiles, ragged pixel tail (539 pixels)
    (6, 14, 1024, 256, 1, 1, 0),   # 16 items -> two groups of 8
    (5, 7, 512, 512, 3, 1, 1),     # 72 items -> nine groups
]


@pytest.mark.parametrize("geom", PAIR_GEOMS)
def test_conv_wgrad_cta_pairs(lib, geom):
    """wgrad_pair_kernel: ONE tcgen05.mma.cta_group::2 of M = 256 per CTA pair, each CTA staging its own dY tile and half of
    the activation items (TMA loads of both CTAs signal the leader's mbarrier; commits multicast to both).  Same checker
    and tolerance as the single-CTA kernel; run twice into the same buffer to check accumulation and determinism."""
    N, H, Cin, Cout, R, stride, pad = geom
    x, w, dy, P = _mk(geom, 3)
    xn, dyn = x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous()
    ref = torch.nn.grad.conv2d_weight(x.float(), w.shape, dy.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
    outs = []
    for _ in range(2):
        dw = torch.zeros(Cout, R, R, Cin, device="cuda")
        lib.check(lib.lib.r3m_b200_conv_wgrad(lib.ptr(dyn), lib.ptr(xn), lib.ptr(dw), N, H, H, Cin, Cout, R, R, stride,
                                              pad, lib.current_stream()))
        lib.check(lib.lib.r3m_b200_check_device_flag())
        outs.append(dw)
    assert rel(outs[0], ref) < 2e-5, rel(outs[0], ref)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("geom,res,relu", [((2, 56, 64, 256, 1, 1, 0), True, True), ((3, 14, 256, 256, 3, 1, 1), False, True),
                                            ((2, 28, 256, 512, 1, 2, 0), False, False), ((1, 7, 512, 2048, 1, 1, 0), True, True)])
def test_conv_forward_folded_bn_epilogue(lib, geom, res, relu):
    """Inference path: y = relu(conv * scale + shift + residual) in the conv epilogue == conv -> eval BN -> add -> relu."""
    N, H, Cin, Cout, R, stride, pad = geom
    x, w, _, P = _mk(geom, 7)
    g = torch.Generator().manual_seed(3)
    scale = (0.5 + torch.rand(Cout, generator=g)).cuda()
    shift = torch.randn(Cout, generator=g).cuda()
    r = torch.randn(N, P, P, Cout, generator=g).cuda().bfloat16() if res else None
    xn, wk = x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous()
    y = torch.full((N, P, P, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.check(lib.lib.r3m_b200_conv_fwd_affine(lib.ptr(xn), lib.ptr(wk), lib.ptr(y), N, H, H, Cin, Cout, R, R, stride,
                                               pad, lib.ptr(scale), lib.ptr(shift), lib.ptr(r), int(relu),
                                               lib.current_stream()))
    lib.check(lib.lib.r3m_b200_check_device_flag())
    ref = F.conv2d(x.float(), w.float(), stride=stride, padding=pad).permute(0, 2, 3, 1) * scale + shift
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    assert rel(y.float(), ref) < 2.5e-3


def test_conv_linearity_at_full_size(lib):
    """Size-independent property at BASELINE c2/c3 size (320 frames, layer3 3x3): conv(x1 + x2) == conv(x1) + conv(x2)
    up to output rounding — no oracle needed."""
    N, H, C = 320, 14, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    x1 = torch.randn(N, H, H, C, device="cuda", generator=g).bfloat16()
    x2 = torch.randn(N, H, H, C, device="cuda", generator=g).bfloat16()
    x12 = (x1.float() + x2.float()).bfloat16()
    w = (torch.randn(C, 3, 3, C, device="cuda", generator=g) / (C * 9) ** 0.5).bfloat16()
    outs = []
    for x in (x1, x2, x12):
        y = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
        lib.check(lib.lib.r3m_b200_conv_fwd(lib.ptr(x), lib.ptr(w), lib.ptr(y), N, H, H, C, C, 3, 3, 1, 1, None, None,
                                            lib.current_stream()))
        outs.append(y.float())
    lib.check(lib.lib.r3m_b200_check_device_flag())
    assert rel(outs[2], outs[0] + outs[1]) < 6e-3  # x12 itself is rounded to bf16 (2^-9) + three output roundings


def test_preprocess_matches_normalize(lib):
    """models_r3m.py:97-98: (obs/255 - mean)/std, laid out as the stem operand (include/r3m_b200.h)."""
    N = 3
    g = torch.Generator().manual_seed(0)
    obs = torch.randint(0, 255, (N, 3, 224, 224), generator=g).float().cuda()
    xs = torch.empty(N, 112, 112, 64, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.lib.r3m_b200_preprocess_stem(lib.ptr(obs), lib.ptr(xs), N, lib.current_stream()))
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda")[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda")[None, :, None, None]
    x = ((obs / 255.0 - mean) / std)
    xp = F.pad(x, (4, 4, 0, 0))  # columns -4..227
    ref = torch.zeros(N, 112, 112, 64, device="cuda")
    for kw in range(4):
        for dy in range(2):
            for dx in range(2):
                for c in range(3):
                    j = kw * 16 + (dy * 2 + dx) * 4 + c
                    cols = 2 * (torch.arange(112, device="cuda") - 2 + kw) + dx + 4
                    ref[:, :, :, j] = xp[:, c, dy::2, :][:, :, cols]
    assert rel(xs.float(), ref) < 2.5e-3
    assert float(xs.float()[..., 3::4].abs().max()) == 0.0  # padded 4th channel is exactly zero


NULL9 = [None] * 9


def _unpack_mask(mask, C):
    bits = (mask.long()[..., None] >> torch.arange(8, device=mask.device)) & 1
    return bits.reshape(mask.shape[0], C).bool()


@pytest.mark.parametrize("C,M,res,relu", [(64, 1000, False, True), (256, 777, True, True), (2048, 98, True, True),
                                           (512, 300, False, False)])
def test_bn_apply_train_and_eval(lib, C, M, res, relu):
    g = torch.Generator().manual_seed(C + M)
    y = (torch.randn(M, C, generator=g) * 2 + 0.5).cuda().bfloat16()
    r = torch.randn(M, C, generator=g).cuda().bfloat16() if res else None
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    beta = (0.1 * torch.randn(C, generator=g)).cuda()
    rm, rv = torch.zeros(C).cuda(), torch.ones(C).cuda()
    yf = y.float()
    ssum, ssq = yf.sum(0), (yf * yf).sum(0)
    sm, sr = torch.empty(C).cuda(), torch.empty(C).cuda()
    a = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    mask = torch.zeros(M, C // 8, device="cuda", dtype=torch.uint8)
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_bn_apply(lib.ptr(y), lib.ptr(a), lib.ptr(r), M, C, int(relu), 1, lib.ptr(ssum),
                                        lib.ptr(ssq), lib.ptr(gamma), lib.ptr(beta), lib.ptr(rm), lib.ptr(rv),
                                        lib.ptr(sm), lib.ptr(sr), lib.ptr(mask), *NULL9, s))
    rm0, rv0 = torch.zeros(C).cuda(), torch.ones(C).cuda()
    ref = F.batch_norm(yf, rm0, rv0, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    assert rel(a.float(), ref) < 2.5e-3
    assert torch.equal(_unpack_mask(mask, C), a.float() > 0)       # the bit mask describes the stored activation
    assert rel(rm, rm0) < 1e-4 and rel(rv, rv0) < 1e-4            # running stats: momentum 0.1, unbiased variance
    assert rel(sm, yf.mean(0)) < 1e-4 and rel(sr, 1 / (yf.var(0, unbiased=False) + 1e-5).sqrt()) < 1e-4
    # eval mode uses the running statistics
    lib.check(lib.lib.r3m_b200_bn_apply(lib.ptr(y), lib.ptr(a), lib.ptr(r), M, C, int(relu), 0, None, None,
                                        lib.ptr(gamma), lib.ptr(beta), lib.ptr(rm), lib.ptr(rv), None, None, None,
                                        *NULL9, s))
    ref = F.batch_norm(yf, rm, rv, gamma, beta, training=False, eps=1e-5)
    if res:
        ref = ref + r.float()
    if relu:
        ref = ref.relu()
    assert rel(a.float(), ref) < 2.5e-3


@pytest.mark.parametrize("C,M", [(256, 1500), (2048, 147)])
def test_dual_bn_forward_backward(lib, C, M):
    """Downsample block tail: a = relu(bn3(y3) + bn_ds(y_ds)) in one pass, and both BatchNorm backwards fed by the same
    masked gradient in one reduce + one apply pass (tv resnet.py:155-161)."""
    g = torch.Generator().manual_seed(C)
    y1 = (torch.randn(M, C, generator=g) * 1.3 + 0.2).cuda().bfloat16()
    y2 = (torch.randn(M, C, generator=g) * 0.7 - 0.1).cuda().bfloat16()
    g1 = (1 + 0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    b1 = (0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    g2 = (1 + 0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    b2 = (0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    dA = torch.randn(M, C, generator=g).cuda().bfloat16()
    l1, l2 = y1.float().requires_grad_(True), y2.float().requires_grad_(True)
    ref = (F.batch_norm(l1, None, None, g1, b1, training=True, eps=1e-5)
           + F.batch_norm(l2, None, None, g2, b2, training=True, eps=1e-5)).relu()
    ref.backward(dA.float())
    stats = lambda t: (t.float().sum(0), (t.float() ** 2).sum(0))  # noqa: E731
    (s1, q1), (s2, q2) = stats(y1), stats(y2)
    bufs = [torch.zeros(C).cuda() for _ in range(8)]
    rm1, rv1, sm1, sr1, rm2, rv2, sm2, sr2 = bufs
    rv1.fill_(1), rv2.fill_(1)
    a = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    mask = torch.zeros(M, C // 8, device="cuda", dtype=torch.uint8)
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_bn_apply(lib.ptr(y1), lib.ptr(a), None, M, C, 1, 1, lib.ptr(s1), lib.ptr(q1),
                                        lib.ptr(g1.detach()), lib.ptr(b1.detach()), lib.ptr(rm1), lib.ptr(rv1),
                                        lib.ptr(sm1), lib.ptr(sr1), lib.ptr(mask), lib.ptr(y2), lib.ptr(s2), lib.ptr(q2),
                                        lib.ptr(g2.detach()), lib.ptr(b2.detach()), lib.ptr(rm2), lib.ptr(rv2),
                                        lib.ptr(sm2), lib.ptr(sr2), s))
    assert rel(a.float(), ref.detach()) < 2.5e-3
    assert rel(sm2, y2.float().mean(0)) < 1e-4 and rel(rm2, 0.1 * y2.float().mean(0)) < 1e-4
    sums, sums2 = torch.zeros(2 * C, device="cuda"), torch.zeros(C, device="cuda")
    dy1 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    dy2 = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    dg1, db1, dg2, db2 = (torch.empty(C).cuda() for _ in range(4))
    lib.check(lib.lib.r3m_b200_bn_backward(lib.ptr(dA), None, lib.ptr(mask), lib.ptr(y1), M, C, lib.ptr(sm1),
                                           lib.ptr(sr1), lib.ptr(g1.detach()), lib.ptr(sums), lib.ptr(dy1), None,
                                           lib.ptr(dg1), lib.ptr(db1), lib.ptr(y2), lib.ptr(sm2), lib.ptr(sr2),
                                           lib.ptr(g2.detach()), lib.ptr(sums2), lib.ptr(dy2), lib.ptr(dg2),
                                           lib.ptr(db2), s))
    # the reference mask comes from the fp32 sum; ours from the bf16-stored activation: identical except at exact ties
    assert rel(dy1.float(), l1.grad) < 6e-3 and rel(dy2.float(), l2.grad) < 6e-3
    assert rel(dg1, g1.grad) < 2e-3 and rel(db1, b1.grad) < 2e-3
    assert rel(dg2, g2.grad) < 2e-3 and rel(db2, b2.grad) < 2e-3


@pytest.mark.parametrize("C,M,mask_kind,want_dz", [(64, 2000, "act", False), (256, 999, "bits", True),
                                                   (1024, 196, "none", False)])
def test_bn_backward_matches_autograd(lib, C, M, mask_kind, want_dz):
    g = torch.Generator().manual_seed(C)
    y = (torch.randn(M, C, generator=g) * 1.5 + 0.3).cuda().bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    beta = (0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    dA = torch.randn(M, C, generator=g).cuda().bfloat16()
    yl = y.float().requires_grad_(True)
    z = F.batch_norm(yl, None, None, gamma, beta, training=True, eps=1e-5)
    a = z.relu() if mask_kind != "none" else z
    a.backward(dA.float())
    a_b = a.detach().bfloat16()
    bits = None
    if mask_kind == "bits":
        on = (a_b.float() > 0).reshape(M, C // 8, 8).long()
        bits = (on << torch.arange(8, device="cuda")).sum(-1).to(torch.uint8).contiguous()
    mean = y.float().mean(0)
    rstd = 1 / (y.float().var(0, unbiased=False) + 1e-5).sqrt()
    sums = torch.zeros(2 * C, device="cuda")
    dy = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    dz = torch.empty(M, C, device="cuda", dtype=torch.bfloat16) if want_dz else None
    dg, db = torch.empty(C).cuda(), torch.empty(C).cuda()
    lib.check(lib.lib.r3m_b200_bn_backward(lib.ptr(dA), lib.ptr(a_b) if mask_kind == "act" else None, lib.ptr(bits),
                                           lib.ptr(y), M, C, lib.ptr(mean), lib.ptr(rstd), lib.ptr(gamma.detach()),
                                           lib.ptr(sums), lib.ptr(dy), lib.ptr(dz), lib.ptr(dg), lib.ptr(db),
                                           *([None] * 8), lib.current_stream()))
    assert rel(dy.float(), yl.grad) < 4e-3
    assert rel(dg, gamma.grad) < 1e-4 and rel(db, beta.grad) < 1e-4
    if want_dz:
        assert rel(dz.float(), dA.float() * (a.detach() > 0)) < 1e-6


def test_stem_pool_forward_backward(lib):
    N, H, C = 2, 16, 64
    g = torch.Generator().manual_seed(3)
    y = torch.randn(N, H, H, C, generator=g).cuda().bfloat16()
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    beta = (0.1 * torch.randn(C, generator=g)).cuda()
    yf = y.float()
    ssum, ssq = yf.sum((0, 1, 2)), (yf * yf).sum((0, 1, 2))
    rm, rv, sm, sr = torch.zeros(C).cuda(), torch.ones(C).cuda(), torch.empty(C).cuda(), torch.empty(C).cuda()
    a = torch.empty(N, H // 2, H // 2, C, device="cuda", dtype=torch.bfloat16)
    am = torch.empty(N, H // 2, H // 2, C, device="cuda", dtype=torch.uint8)
    ymax = torch.empty(N, H // 2, H // 2, C, device="cuda", dtype=torch.bfloat16)
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_stem_bn_relu_maxpool(lib.ptr(y), lib.ptr(a), lib.ptr(am), lib.ptr(ymax), N, H, H, C, 1, lib.ptr(ssum),
                                                    lib.ptr(ssq), lib.ptr(gamma), lib.ptr(beta), lib.ptr(rm),
                                                    lib.ptr(rv), lib.ptr(sm), lib.ptr(sr), s))
    x = yf.permute(0, 3, 1, 2).requires_grad_(True)
    z = F.batch_norm(x, None, None, gamma, beta, training=True, eps=1e-5).relu()
    z.retain_grad()
    p = F.max_pool2d(z, 3, 2, 1)
    assert rel(a.float(), p.permute(0, 2, 3, 1)) < 2.5e-3
    dA = torch.randn(N, H // 2, H // 2, C, generator=g).cuda().bfloat16()
    p.backward(dA.float().permute(0, 3, 1, 2))
    dz = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.lib.r3m_b200_maxpool_backward(lib.ptr(dA), lib.ptr(a), lib.ptr(am), lib.ptr(dz), N, H, H, C, s))
    # gradient w.r.t. the BN output (ReLU mask applied): compare with autograd's grad at z, masked
    want = (z.grad * (z.detach() > 0)).permute(0, 2, 3, 1)
    assert rel(dz.float(), want) < 4e-3
    # argmax codes: 0..8 = window position of the maximum, 15 = maximum clipped by the ReLU (no gradient)
    codes = am.cpu()
    assert bool(((codes <= 8) | (codes == 15)).all())
    assert bool(((codes == 15) == (a.float().cpu() == 0)).all())

    # fused form the engine runs: maxpool backward + ReLU mask + BatchNorm backward in two passes -> dy, dgamma, dbeta
    # ymax = the raw conv output at each window's argmax
    zmax = F.max_pool2d(F.batch_norm(yf.permute(0, 3, 1, 2), None, None, gamma, beta, training=True, eps=1e-5).relu(), 3, 2, 1)
    live = (zmax > 0).permute(0, 2, 3, 1)
    recon = (ymax.float() - sm) * sr * gamma + beta  # BN of ymax must reproduce the pooled activation where it is live
    assert rel(recon[live], a.float()[live]) < 5e-3
    for use_ymax in (True, False):  # reduce pass over the pooled elements / over 2x2 quads of conv-output pixels
        sums = torch.zeros(2 * C, device="cuda")
        dy = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
        dgamma, dbeta = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        lib.check(lib.lib.r3m_b200_stem_backward(lib.ptr(dA), lib.ptr(am), lib.ptr(ymax) if use_ymax else None, lib.ptr(y),
                                                 N, H, H, C, lib.ptr(sm), lib.ptr(sr), lib.ptr(gamma), lib.ptr(sums),
                                                 lib.ptr(dy), lib.ptr(dgamma), lib.ptr(dbeta), s))
        _check_stem_backward(dy, dgamma, dbeta, yf, gamma, beta, want)


def _check_stem_backward(dy, dgamma, dbeta, yf, gamma, beta, want):
    # reference: autograd through BatchNorm, ReLU and the pooling (x.grad), and BatchNorm's parameter gradients
    # driven by the exact fp32 masked gradient (the fused kernel never rounds it to bf16)
    x2 = yf.permute(0, 3, 1, 2).detach().requires_grad_(True)
    g2, b2 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.batch_norm(x2, None, None, g2, b2, training=True, eps=1e-5).backward(want.permute(0, 3, 1, 2).contiguous())
    assert rel(dy.float(), x2.grad.permute(0, 2, 3, 1)) < 4e-3
    assert rel(dgamma, g2.grad) < 1e-4 and rel(dbeta, b2.grad) < 1e-4


def test_avgpool_forward_backward(lib):
    N, HW, C = 7, 49, 2048
    g = torch.Generator().manual_seed(4)
    a = torch.randn(N, HW, C, generator=g).cuda().bfloat16()
    out = torch.empty(N, C, device="cuda")
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_avgpool_forward(lib.ptr(a), lib.ptr(out), N, HW, C, s))
    assert rel(out, a.float().mean(1)) < 1e-6
    dE = torch.randn(N, C, generator=g).cuda()
    dA = torch.empty(N, HW, C, device="cuda", dtype=torch.bfloat16)
    lib.check(lib.lib.r3m_b200_avgpool_backward(lib.ptr(dE), lib.ptr(dA), N, HW, C, s))
    assert rel(dA.float(), (dE / HW)[:, None, :].expand(N, HW, C)) < 2.5e-3


@pytest.mark.parametrize("B,D", [(2, 512), (64, 2048), (7, 512)])
def test_loss_heads_match_oracle(lib, B, D):
    """Loss tolerance of the north star: <= 1e-4 relative on identical embeddings (here ~1e-6)."""
    from oracle import r3m_oracle as O

    g = torch.Generator().manual_seed(B)
    E = torch.randn(5 * B, D, generator=g).relu() * 0.05
    E[3] = 0.0                                   # zero row: sub-gradients must vanish
    perms = O.draw_permutations(B, 5)
    perms[9] = torch.arange(B)                   # fixed points: zero distance -> zero gradient, exp(0) in the sum
    hyper = dict(l2weight=1e-3, l1weight=1e-4, tcnweight=1.0, langweight=0.0)
    e = E.clone().requires_grad_(True)
    full, m = O.losses({}, e, perms, hyper)
    full.backward()
    Ed = E.cuda()
    dE = torch.full((5 * B, D), float("nan"), device="cuda")
    metrics = torch.zeros(16, device="cuda")
    pd = perms.to(torch.int32).cuda()
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_loss_lp(lib.ptr(Ed), lib.ptr(dE), 5 * B, D, 1e-3, 1e-4, lib.ptr(metrics), s))
    lib.check(lib.lib.r3m_b200_loss_tcn(lib.ptr(Ed), lib.ptr(dE), lib.ptr(pd), B, D, 1.0, lib.ptr(metrics), s))
    got = metrics.cpu().tolist()
    for k, i in (("l2loss", 0), ("l1loss", 1), ("l0loss", 2), ("tcnloss", 7), ("aligned", 8), ("full_loss", 9)):
        assert abs(got[i] - m[k]) <= 1e-4 * max(abs(m[k]), 1e-6), (k, got[i], m[k])
    assert rel(dE.cpu(), e.grad) < 1e-4
    assert torch.isfinite(dE).all()


@pytest.mark.parametrize("B,D", [(6, 512), (16, 2048)])
def test_tcn_head_with_cosine_similarity(lib, B, D):
    """R3M(l2dist=False): sim = nn.CosineSimilarity(dim=1) (models_r3m.py:105-107) in the fused TCN head, against the
    oracle (torch autograd through F.cosine_similarity) on identical embeddings."""
    from oracle import r3m_oracle as O

    g = torch.Generator().manual_seed(100 + B)
    E = torch.randn(5 * B, D, generator=g).relu() * 0.05
    perms = O.draw_permutations(B, 6)
    perms[10] = torch.arange(B)                  # es2 against itself: cosine 1, gradient exactly zero
    hyper = dict(l2weight=0.0, l1weight=0.0, tcnweight=0.7, langweight=0.0, l2dist=False)
    e = E.clone().requires_grad_(True)
    full, m = O.losses({}, e, perms, hyper)
    full.backward()
    Ed = E.cuda()
    dE = torch.zeros(5 * B, D, device="cuda")
    metrics = torch.zeros(16, device="cuda")
    pd = perms.to(torch.int32).cuda()
    s = lib.current_stream()
    lib.check(lib.lib.r3m_b200_loss_tcn_sim(lib.ptr(Ed), lib.ptr(dE), lib.ptr(pd), B, D, 0.7, 0, lib.ptr(metrics), s))
    got = metrics.cpu().tolist()
    assert abs(got[7] - m["tcnloss"]) <= 1e-4 * abs(m["tcnloss"]), (got[7], m["tcnloss"])
    assert abs(got[8] - m["aligned"]) <= 1.0 / B + 1e-6
    assert abs(got[9] - 0.7 * m["tcnloss"]) <= 1e-4 * abs(0.7 * m["tcnloss"])
    assert rel(dE.cpu(), e.grad) < 1e-4
    # and the L2 flavour through the same entry point equals r3m_b200_loss_tcn
    d1, d2 = torch.zeros_like(dE), torch.zeros_like(dE)
    m1, m2 = torch.zeros(16, device="cuda"), torch.zeros(16, device="cuda")
    lib.check(lib.lib.r3m_b200_loss_tcn_sim(lib.ptr(Ed), lib.ptr(d1), lib.ptr(pd), B, D, 0.7, 1, lib.ptr(m1), s))
    lib.check(lib.lib.r3m_b200_loss_tcn(lib.ptr(Ed), lib.ptr(d2), lib.ptr(pd), B, D, 0.7, lib.ptr(m2), s))
    assert rel(d1, d2) < 1e-6 and abs(float(m1[7]) - float(m2[7])) < 1e-6


def test_adam_matches_torch(lib):
    n = 100003
    g = torch.Generator().manual_seed(9)
    p0, grad = torch.randn(n, generator=g), torch.randn(n, generator=g) * 1e-3
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4)
    p, m, v = p0.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        ref.grad = grad * step
        opt.step()
        gd = (grad * step * 2.0).cuda()  # grad_scale 0.5 undoes the factor 2 (the all-reduce SUM -> mean path)
        lib.check(lib.lib.r3m_b200_adam(lib.ptr(p), lib.ptr(gd), lib.ptr(m), lib.ptr(v), lib.ptr(pb), n, 1e-4, step,
                                        0.5, lib.current_stream()))
    assert rel(p.cpu(), ref.detach()) < 1e-6
    assert torch.equal(pb.cpu(), p.cpu().bfloat16())
