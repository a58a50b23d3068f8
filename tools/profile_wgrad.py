import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from r3m_b200 import _lib as L
N = 320
shapes = [(14, 256, 256, 3, 1, 1), (56, 64, 64, 3, 1, 1), (56, 64, 256, 1, 1, 0), (7, 512, 512, 3, 1, 1)]
s = L.current_stream()
for (H, Cin, Cout, R, stride, pad) in shapes:
    x = torch.randn(N, H, H, Cin, device="cuda").bfloat16()
    P = (H + 2 * pad - R) // stride + 1
    dy = torch.randn(N, P, P, Cout, device="cuda").bfloat16()
    dw = torch.zeros(Cout, R, R, Cin, device="cuda")
    for _ in range(2):
        L.check(L.lib.r3m_b200_conv_wgrad(L.ptr(dy), L.ptr(x), L.ptr(dw), N, H, H, Cin, Cout, R, R, stride, pad, s))
    torch.cuda.synchronize()
