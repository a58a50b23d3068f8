"""Diagnostic sweep of the tcgen05 kernels against torch fp32 on the same bf16-rounded operands.

Each case runs in its own subprocess (a device fault or watchdog in one case must not poison the others) and appends
one JSON line to gpurun_out/kernel_check.jsonl.  Usage: python tools/gpu_kernel_check.py [--only fwd,wgrad,dgrad]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GEOMS = [
    # N, H, W, Cin, Cout, R, stride, pad
    (2, 56, 56, 64, 64, 1, 1, 0),
    (2, 56, 56, 64, 64, 3, 1, 1),
    (3, 28, 28, 128, 128, 3, 1, 1),
    (2, 14, 14, 256, 256, 3, 1, 1),
    (3, 7, 7, 512, 512, 3, 1, 1),
    (4, 14, 14, 256, 1024, 1, 1, 0),
    (2, 56, 56, 128, 128, 3, 2, 1),
    (2, 56, 56, 256, 512, 1, 2, 0),
    (5, 14, 14, 1024, 256, 1, 1, 0),
    (2, 112, 112, 64, 64, 4, 1, 2),  # stand-in for the packed stem geometry (taps in both axes here)
]


def report(got, ref, extra):
    import torch

    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    rel = (diff.norm() / (ref.norm() + 1e-30)).item()
    out = dict(rel_l2=rel, max_abs=diff.max().item(), ref_absmax=ref.abs().max().item(),
               finite=bool(torch.isfinite(got).all().item()))
    bad = (diff > 0.02 * ref.abs().max() + 1e-3).nonzero()
    out["n_bad"] = int(bad.shape[0])
    out["n_total"] = int(got.numel())
    samples = []
    for idx in bad[:6].tolist():
        samples.append((idx, float(got[tuple(idx)]), float(ref[tuple(idx)])))
    out["bad_samples"] = samples
    out.update(extra)
    return out


def run_case(kind, geom):
    import torch
    import torch.nn.functional as F
    from r3m_b200 import _lib as L

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    N, H, W, Cin, Cout, R, stride, pad = geom
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.randn(N, Cin, H, W, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(dev)
    wb = w.bfloat16()
    P = (H + 2 * pad - R) // stride + 1
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_krsc = wb.permute(0, 2, 3, 1).contiguous()
    s = L.current_stream()
    if kind == "fwd":
        y = torch.full((N, P, P, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
        ssum = torch.zeros(Cout, device=dev)
        ssq = torch.zeros(Cout, device=dev)
        L.check(L.lib.r3m_b200_conv_fwd(L.ptr(x_nhwc), L.ptr(w_krsc), L.ptr(y), N, H, W, Cin, Cout, R, R, stride, pad,
                                        L.ptr(ssum), L.ptr(ssq), s))
        L.check(L.lib.r3m_b200_check_device_flag())
        ref = F.conv2d(x.float(), wb.float(), stride=stride, padding=pad).permute(0, 2, 3, 1)
        out = report(y, ref, {})
        yb = y.float()
        out["stat_sum_rel"] = ((ssum - yb.sum((0, 1, 2))).norm() / (yb.sum((0, 1, 2)).norm() + 1e-20)).item()
        out["stat_sq_rel"] = ((ssq - (yb * yb).sum((0, 1, 2))).norm() / (yb * yb).sum((0, 1, 2)).norm()).item()
        return out
    dy = torch.randn(N, Cout, P, P, generator=g).to(dev).bfloat16()
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous()
    if kind == "wgrad":
        dw = torch.zeros(Cout, R, R, Cin, device=dev)
        L.check(L.lib.r3m_b200_conv_wgrad(L.ptr(dy_nhwc), L.ptr(x_nhwc), L.ptr(dw), N, H, W, Cin, Cout, R, R, stride,
                                          pad, s))
        L.check(L.lib.r3m_b200_check_device_flag())
        ref = torch.nn.grad.conv2d_weight(x.float(), w.shape, dy.float(), stride=stride, padding=pad)
        return report(dw, ref.permute(0, 2, 3, 1), {})
    if kind == "dgrad":
        wd = torch.empty(Cout * R * R * Cin, device=dev, dtype=torch.bfloat16)
        wm = wb.float().permute(0, 2, 3, 1).contiguous()
        L.check(L.lib.r3m_b200_pack_dgrad_filter(L.ptr(wm), L.ptr(wd), Cout, R, R, Cin, stride, pad, s))
        dx = torch.full((N, H, W, Cin), float("nan"), device=dev, dtype=torch.bfloat16)
        L.check(L.lib.r3m_b200_conv_dgrad(L.ptr(dy_nhwc), L.ptr(wd), L.ptr(dx), N, H, W, Cin, Cout, R, R, stride, pad,
                                          0, s))
        L.check(L.lib.r3m_b200_check_device_flag())
        ref = torch.nn.grad.conv2d_input(x.shape, wb.float(), dy.float(), stride=stride, padding=pad)
        out = report(dx, ref.permute(0, 2, 3, 1), {})
        # accumulate mode
        base = torch.randn(N, H, W, Cin, generator=g).to(dev).bfloat16()
        dx2 = base.clone()
        L.check(L.lib.r3m_b200_conv_dgrad(L.ptr(dy_nhwc), L.ptr(wd), L.ptr(dx2), N, H, W, Cin, Cout, R, R, stride, pad,
                                          1, s))
        L.check(L.lib.r3m_b200_check_device_flag())
        ref2 = ref.permute(0, 2, 3, 1) + base.float()
        out["accumulate_rel_l2"] = ((dx2.float() - ref2).norm() / ref2.norm()).item()
        return out
    raise ValueError(kind)


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "--case":
        kind = sys.argv[2]
        geom = tuple(json.loads(sys.argv[3]))
        try:
            res = run_case(kind, geom)
            res["status"] = "ok"
        except Exception as e:  # noqa: BLE001 - diagnostics only
            res = {"status": "error", "error": repr(e)[:600]}
        res.update(kind=kind, geom=geom)
        print("RESULT " + json.dumps(res))
        return
    only = None
    if len(sys.argv) >= 3 and sys.argv[1] == "--only":
        only = sys.argv[2].split(",")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "kernel_check.jsonl")
    with open(path, "w") as f:
        for kind in ("fwd", "wgrad", "dgrad"):
            if only and kind not in only:
                continue
            for geom in GEOMS:
                try:
                    p = subprocess.run([sys.executable, __file__, "--case", kind, json.dumps(geom)],
                                       capture_output=True, text=True, timeout=180)
                    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                    if line:
                        res = json.loads(line[-1][7:])
                    else:
                        res = {"status": "crash", "kind": kind, "geom": geom, "rc": p.returncode,
                               "stderr": p.stderr[-800:]}
                except subprocess.TimeoutExpired:
                    res = {"status": "timeout", "kind": kind, "geom": geom}
                f.write(json.dumps(res) + "\n")
                f.flush()
                brief = {k: res.get(k) for k in ("status", "kind", "geom", "rel_l2", "n_bad", "stat_sum_rel",
                                                 "accumulate_rel_l2", "error")}
                print(json.dumps(brief), flush=True)


if __name__ == "__main__":
    main()
