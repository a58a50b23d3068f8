"""How much of a streaming consumer's DRAM read does the 126 MB L2 save when it walks its input in the OPPOSITE order
of the producer that has just written it?  Producer: y = x * 2 over S bytes in address order.  Consumer: z = y + 1 in
`chunks` pieces launched first-to-last (same order: what the step does today) or last-to-first (reversed)."""
import json
import os
import sys

import torch

out = {}
for mb in (32, 64, 96, 128, 192, 256, 512, 1024):
    n = mb * 1024 * 1024 // 2
    x = torch.randn(n, device="cuda").bfloat16()
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    chunks = 16
    step = n // chunks
    res = {}
    for order in ("same", "reversed"):
        idx = list(range(chunks)) if order == "same" else list(range(chunks - 1, -1, -1))
        tot = 0.0
        for it in range(6):
            torch.mul(x, 2, out=y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for c in idx:
                torch.add(y[c * step:(c + 1) * step], 1, out=z[c * step:(c + 1) * step])
            e1.record()
            torch.cuda.synchronize()
            if it > 0:
                tot += e0.elapsed_time(e1)
        res[order] = tot / 5
    res["gbs_same"] = 2 * mb / 1e3 / (res["same"] / 1e3) * 1.048576
    res["gbs_reversed"] = 2 * mb / 1e3 / (res["reversed"] / 1e3) * 1.048576
    out[f"{mb}MB"] = res
    print(mb, res)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                                 "s4_l2_order.json"), "w"), indent=1)
