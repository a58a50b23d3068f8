"""Forward-only timings for BASELINE configs c1 / c2 (eval = load_r3m inference path, train = update()'s encode)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from r3m_b200 import R3M

for size, B in ((18, 4), (50, 256), (34, 256), (18, 256)):
    m = R3M("cuda", 1e-4, 1024, size=size, langweight=0.0).cuda()
    x = torch.randint(0, 255, (B, 3, 224, 224), device="cuda").float()
    for mode in ("eval", "train"):
        m.eval() if mode == "eval" else m.train()
        with torch.no_grad():
            for _ in range(5):
                m(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                out = m(x)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"ResNet-{size} forward {mode:5s} batch {B}: {ms:.3f} ms  {B / ms * 1e3:.0f} frames/s  (launches {m._any_engine().launches()})")
