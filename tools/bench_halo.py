"""3x3 / 64 -> 64 / 56 x 56 / 320 frames (ResNet-50 layer1 conv2 at c3 size): halo3x3_kernel (R3M_HALO=1, default) vs the
im2col kernel (R3M_HALO=0), device-timed, forward without statistics."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from r3m_b200 import _lib as L  # noqa: E402

N, H, C = 320, 56, 64
x = torch.randn(N, H, H, C, device="cuda").bfloat16()
w = (torch.randn(C, 3, 3, C, device="cuda") / 24).bfloat16()
y = torch.empty(N, H, H, C, device="cuda", dtype=torch.bfloat16)
s = L.current_stream()


def run():
    L.check(L.lib.r3m_b200_conv_fwd(L.ptr(x), L.ptr(w), L.ptr(y), N, H, H, C, C, 3, 3, 1, 1, None, None, s))


for _ in range(5):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"R3M_HALO={os.environ.get('R3M_HALO', '1')}: {ms * 1e3:.1f} us, {2 * N * H * H * C * C * 9 / ms / 1e9:.0f} TFLOP/s")
