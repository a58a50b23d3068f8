"""Per-launch profile of one RN50 update step (in-stream CUDA events) -> gpurun_out/step_ops.csv, and a few
stand-alone conv launches for ncu (`--convs`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def step_ops(size=50, clips=64, lang=1):
    import bench
    import r3m_b200
    from r3m_b200 import R3M
    from r3m_b200.trainer import draw_permutations

    r3m_b200.set_lang_encoder_factory(bench.StubLangEncoder)
    m = R3M("cuda", 1e-4, 1024, size=size, l2weight=1e-5, l1weight=1e-5, langweight=float(lang), tcnweight=1.0).cuda()
    frames = torch.randint(0, 255, (clips * 5, 3, 224, 224), device="cuda").float()
    eng = m._engine(clips * 5)
    perms = draw_permutations(clips, m.langweight, m.tcnweight).cuda()
    emb = mask = None
    if lang:
        emb = m.lang_enc([""] * clips).cuda().float().contiguous()
        mask = torch.ones(clips, device="cuda")
    ncu = os.environ.get("R3M_NCU") == "1"  # under `ncu --profile-from-start off`: capture exactly the third step
    for i in range(3):
        if ncu and i == 2:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        fam = eng.profile_update(frames, perms, emb, mask, 1e-5, 1e-5, float(lang), 1.0, 1e-4, i + 1)
        if ncu and i == 2:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
    ops = eng.profile_ops()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"step_ops_rn{size}.csv"), "w") as f:
        f.write("idx,family,ms,gflop,mbytes,tflops,gbs,label\n")
        for i, (fm, ms, fl, by, label) in enumerate(ops):
            f.write(f"{i},{fm},{ms:.4f},{fl / 1e9:.3f},{by / 1e6:.2f},{fl / ms / 1e9 if ms > 0 else 0:.1f},"
                    f"{by / ms / 1e6 if ms > 0 else 0:.0f},{label}\n")
    print({k: round(v["ms"], 3) for k, v in fam.items()})


def convs():
    from r3m_b200 import _lib as L

    N = 320
    shapes = [(56, 64, 64, 3, 1, 1), (14, 256, 256, 3, 1, 1), (56, 64, 256, 1, 1, 0), (56, 256, 64, 1, 1, 0),
              (14, 1024, 256, 1, 1, 0), (7, 512, 2048, 1, 1, 0)]
    s = L.current_stream()
    for (H, Cin, Cout, R, stride, pad) in shapes:
        x = torch.randn(N, H, H, Cin, device="cuda").bfloat16()
        w = (torch.randn(Cout, R, R, Cin, device="cuda") / (Cin * R * R) ** 0.5).bfloat16()
        P = (H + 2 * pad - R) // stride + 1
        y = torch.empty(N, P, P, Cout, device="cuda", dtype=torch.bfloat16)
        ssum = torch.zeros(Cout, device="cuda")
        ssq = torch.zeros(Cout, device="cuda")
        for _ in range(3):
            L.check(L.lib.r3m_b200_conv_fwd(L.ptr(x), L.ptr(w), L.ptr(y), N, H, H, Cin, Cout, R, R, stride, pad,
                                            L.ptr(ssum), L.ptr(ssq), s))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            L.check(L.lib.r3m_b200_conv_fwd(L.ptr(x), L.ptr(w), L.ptr(y), N, H, H, Cin, Cout, R, R, stride, pad,
                                            L.ptr(ssum), L.ptr(ssq), s))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * N * P * P * Cout * Cin * R * R
        by = 2.0 * (N * H * H * Cin + N * P * P * Cout)
        print(f"conv H={H} Cin={Cin} Cout={Cout} R={R}: {ms:.4f} ms  {fl / ms / 1e9:.0f} TFLOP/s  {by / ms / 1e6:.0f} GB/s")
        dy = torch.randn(N, P, P, Cout, device="cuda").bfloat16()
        dw = torch.zeros(Cout, R, R, Cin, device="cuda")
        for _ in range(2):
            L.check(L.lib.r3m_b200_conv_wgrad(L.ptr(dy), L.ptr(x), L.ptr(dw), N, H, H, Cin, Cout, R, R, stride, pad, s))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            L.check(L.lib.r3m_b200_conv_wgrad(L.ptr(dy), L.ptr(x), L.ptr(dw), N, H, H, Cin, Cout, R, R, stride, pad, s))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  wgrad: {ms:.4f} ms  {fl / ms / 1e9:.0f} TFLOP/s")


def wgrads():
    from r3m_b200 import _lib as L

    N = 320
    shapes = [(14, 256, 256, 3, 1, 1), (56, 64, 64, 3, 1, 1), (56, 64, 256, 1, 1, 0), (7, 512, 512, 3, 1, 1),
              (28, 128, 128, 3, 1, 1), (14, 1024, 256, 1, 1, 0)]
    s = L.current_stream()
    out = []
    for (H, Cin, Cout, R, stride, pad) in shapes:
        x = torch.randn(N, H, H, Cin, device="cuda").bfloat16()
        P = (H + 2 * pad - R) // stride + 1
        dy = torch.randn(N, P, P, Cout, device="cuda").bfloat16()
        dw = torch.zeros(Cout, R, R, Cin, device="cuda")
        for _ in range(2):
            L.check(L.lib.r3m_b200_conv_wgrad(L.ptr(dy), L.ptr(x), L.ptr(dw), N, H, H, Cin, Cout, R, R, stride, pad, s))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            L.check(L.lib.r3m_b200_conv_wgrad(L.ptr(dy), L.ptr(x), L.ptr(dw), N, H, H, Cin, Cout, R, R, stride, pad, s))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * N * P * P * Cout * Cin * R * R
        out.append(f"{H}/{Cin}/{Cout}/{R}: {ms:.4f}ms {fl / ms / 1e9:.0f}TF")
    print(os.environ.get("R3M_WGRAD_GROUP"), os.environ.get("R3M_WGRAD_STAGES"), os.environ.get("R3M_WGRAD_WAVES"),
          " | ".join(out))


if __name__ == "__main__":
    if "--wgrads" in sys.argv:
        wgrads()
    elif "--convs" in sys.argv:
        convs()
    else:
        step_ops()
