"""How much of a step is GPU idle time between update() calls?  Times K updates with the per-step metrics read-back
(the reference's contract: update() returns python floats) and with the read-back deferred to the end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import r3m_b200  # noqa: E402
from r3m_b200 import R3M, Trainer  # noqa: E402

r3m_b200.set_lang_encoder_factory(bench.StubLangEncoder)
m = R3M("cuda", 1e-4, 1024, size=50, l2weight=1e-5, l1weight=1e-5, langweight=1.0, tcnweight=1.0)
model = torch.nn.DataParallel(m.cuda(), device_ids=[0])
tr = Trainer(10 ** 9)
B = 64
frames = torch.randint(0, 255, (B, 5, 3, 224, 224), device="cuda").float()
lang = bench.sentences_for(B)
for _ in range(5):
    tr.update(model, (frames, lang), 0)


def timed(n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tr.update(model, (frames, lang), 0)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = timed(20)
eng = m._engine(B * 5)
orig = eng.read_metrics
eng.read_metrics = lambda: [0.0] * 16  # no sync, no read-back
b = timed(20)
eng.read_metrics = orig
print(f"per-step sync: {a:.3f} ms/step   deferred sync: {b:.3f} ms/step   gap {a - b:.3f} ms")
