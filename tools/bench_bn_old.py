import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
lib = ctypes.CDLL(os.path.join(ROOT, "tools", "libold_tmp.so"))
vp, ci = ctypes.c_void_p, ctypes.c_int
lib.r3m_b200_bn_backward.argtypes = [vp, vp, vp, ci, ci] + [vp] * 9
lib.r3m_b200_bn_apply.argtypes = [vp, vp, vp, ci, ci, ci, ci] + [vp] * 9
P = lambda t: None if t is None else vp(t.data_ptr())
M, C = 1003520, 256
y = torch.randn(M, C, device="cuda").bfloat16(); dA = torch.randn(M, C, device="cuda").bfloat16()
a = torch.randn(M, C, device="cuda").relu().bfloat16()
mean, rstd, gamma = torch.zeros(C).cuda(), torch.ones(C).cuda(), torch.ones(C).cuda()
sums = torch.zeros(2 * C).cuda(); dy = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
dg, db = torch.empty(C).cuda(), torch.empty(C).cuda()
s = vp(torch.cuda.current_stream().cuda_stream)
E = M * C * 2 / 1e9
for kind in ("act", "none"):
    def run():
        assert lib.r3m_b200_bn_backward(P(dA), P(a) if kind == "act" else None, P(y), M, C, P(mean), P(rstd), P(gamma), P(sums), P(dy), None, P(dg), P(db), s) == 0
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nb = {"act": 7, "none": 5}[kind] * E
    print(f"OLD bn_backward mask={kind}: {ms:.3f} ms, {nb / ms * 1e3:.0f} GB/s")
