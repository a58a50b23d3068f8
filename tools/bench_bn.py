import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from r3m_b200 import _lib as L
M, C = 1003520, 256
y = torch.randn(M, C, device="cuda").bfloat16()
dA = torch.randn(M, C, device="cuda").bfloat16()
a = torch.randn(M, C, device="cuda").relu().bfloat16()
bits = torch.randint(0, 255, (M, C // 8), device="cuda", dtype=torch.uint8)
mean, rstd, gamma = torch.zeros(C).cuda(), torch.ones(C).cuda(), torch.ones(C).cuda()
sums = torch.zeros(2 * C).cuda()
dy = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
dg, db = torch.empty(C).cuda(), torch.empty(C).cuda()
s = L.current_stream()
E = M * C * 2 / 1e9
for kind in ("act", "bits", "none"):
    def run():
        L.check(L.lib.r3m_b200_bn_backward(L.ptr(dA), L.ptr(a) if kind == "act" else None, L.ptr(bits) if kind == "bits" else None,
                                           L.ptr(y), M, C, L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(sums), L.ptr(dy), None,
                                           L.ptr(dg), L.ptr(db), *([None] * 8), s))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nb = {"act": 3 + 4, "bits": 2.06 + 3.06, "none": 2 + 3}[kind] * E
    print(f"bn_backward mask={kind}: {ms:.3f} ms (reduce+apply), {nb / ms * 1e3:.0f} GB/s algorithmic")
# bn_apply with / without mask_out
ssum, ssq = torch.zeros(C).cuda(), torch.ones(C).cuda() * M
beta = torch.zeros(C).cuda(); rm, rv, sm, sr = (torch.zeros(C).cuda() for _ in range(4))
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
for with_mask in (False, True):
    def run():
        L.check(L.lib.r3m_b200_bn_apply(L.ptr(y), L.ptr(out), None, M, C, 1, 1, L.ptr(ssum), L.ptr(ssq), L.ptr(gamma), L.ptr(beta),
                                        L.ptr(rm), L.ptr(rv), L.ptr(sm), L.ptr(sr), L.ptr(bits) if with_mask else None, *([None] * 9), s))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"bn_apply mask_out={with_mask}: {ms:.3f} ms, {(2 + with_mask / 16) * E / ms * 1e3:.0f} GB/s")
