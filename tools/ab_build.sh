#!/bin/bash
# A/B of a compile-time variant on the GPU box: tools/ab_build.sh "<EXTRA nvcc flags>" <tag>  -> gpurun_out/ab_<tag>_ops.csv
set -u
make -C r3m_b200/csrc clean >/dev/null; make -C r3m_b200/csrc -j16 EXTRA="$1" >/dev/null 2>&1
python tools/profile_step.py 2>&1 | tail -1
cp gpurun_out/step_ops_rn50.csv gpurun_out/ab_$2_ops.csv
python bench.py --no-cpu-baseline > gpurun_out/ab_$2.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/ab_$2.json'));print('$2',d['value'],d['ms_per_step'],d['e2e']['ms_per_step'])"
