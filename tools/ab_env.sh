#!/bin/bash
# In-box A/B of environment-selected variants: tools/ab_env.sh "<name>=<env assignments>" ... ; each variant runs the
# N=1 bench (no CPU / cuDNN reference legs) `reps` times, interleaved, and prints ms per step.
reps=${REPS:-2}
for r in $(seq $reps); do
  for v in "$@"; do
    name="${v%%=*}"; envs="${v#*=}"
    out=$(env $envs python bench.py --no-cpu-baseline --no-gpu-reference --no-other-configs --steps 20 --warmup 5 2>/dev/null)
    echo "$out" > gpurun_out/ab_${name}.json
    python - "$name" <<PY
import json,sys
d=json.loads('''$out''')
f=d["roofline"]["families"]
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], {k: round(v["ms"],3) for k,v in f.items()})
PY
  done
done
