"""Does chunking a layer so that a producer's output is still L2 resident (126 MB) pay?  wgrad / dgrad / bn_apply on a
40-frame chunk right after the kernel that wrote their input (hot) vs after an L2 flush (cold) vs the full 320 frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from r3m_b200 import _lib as L
s = L.current_stream()
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def timeit(fn, pre, iters=8):
    tot = 0.0
    for _ in range(iters):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters

for (H, Cin, Cout, R, pad) in [(56, 64, 256, 1, 0), (56, 64, 64, 3, 1), (28, 128, 512, 1, 0)]:
    for N in (40, 320):
        x = torch.randn(N, H, H, Cin, device="cuda").bfloat16()
        dy = torch.randn(N, H, H, Cout, device="cuda").bfloat16()
        dw = torch.zeros(Cout, R, R, Cin, device="cuda")
        w = torch.randn(Cout, R, R, Cin, device="cuda").bfloat16()
        y = torch.empty(N, H, H, Cout, device="cuda", dtype=torch.bfloat16)
        def wgrad():
            L.check(L.lib.r3m_b200_conv_wgrad(L.ptr(dy), L.ptr(x), L.ptr(dw), N, H, H, Cin, Cout, R, R, 1, pad, s))
        def fwd():
            L.check(L.lib.r3m_b200_conv_fwd(L.ptr(x), L.ptr(w), L.ptr(y), N, H, H, Cin, Cout, R, R, 1, pad, None, None, s))
        def hot():   # rewrite the inputs (stay in L2)
            dy.mul_(1.0); x.mul_(1.0)
        def cold():
            dy.mul_(1.0); x.mul_(1.0); flush.zero_()
        for _ in range(2): wgrad(); fwd()
        torch.cuda.synchronize()
        mb = (dy.numel() + x.numel()) * 2 / 1e6
        print(f"H={H} {Cin}->{Cout} R={R} N={N} ({mb:.0f} MB in): wgrad hot {timeit(wgrad, hot)*320/N:.3f} cold {timeit(wgrad, cold)*320/N:.3f} | "
              f"fwd hot {timeit(fwd, hot)*320/N:.3f} cold {timeit(fwd, cold)*320/N:.3f}  (ms, scaled to 320 frames)")
