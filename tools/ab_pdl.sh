#!/bin/bash
# A/B of the programmatic-dependent-launch variants (run on the GPU box; rebuilds the library per variant).
set -u
mkdir -p gpurun_out
summ() { python - "$@" <<'P'
import json,sys
for n in sys.argv[1:]:
    try:
        d=json.load(open(f"gpurun_out/ab_{n}.json")); f=d["roofline"]["families"]
        print(n, "step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k:round(v["ms"],2) for k,v in f.items()})
    except Exception as ex: print(n,"ERR",ex)
P
}
R3M_PDL=0 python bench.py > gpurun_out/ab_pdl0.json 2>/dev/null
R3M_PDL=1 python bench.py > gpurun_out/ab_trig1.json 2>/dev/null
for t in 2 0; do
  make -C r3m_b200/csrc clean >/dev/null; make -C r3m_b200/csrc -j16 EXTRA=-DR3M_PDL_TRIGGER=$t >/dev/null 2>&1
  R3M_PDL=1 python bench.py > gpurun_out/ab_trig$t.json 2>/dev/null
done
summ pdl0 trig1 trig2 trig0
