#!/bin/bash
# Round-end evidence run on ONE B200 box (gpurun): GPU tests, smoke, both bench arms, and the ncu launch list of the bench
# command.  Everything lands in gpurun_out/ with the given prefix.
p=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${p}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${p}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${p}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${p}_smoke.log
[ -n "${SKIP_REF:-}" ] || python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${p}_bench_ref.json 2> gpurun_out/${p}_bench_ref.err
python bench.py > gpurun_out/${p}_bench_n1.json 2> gpurun_out/${p}_bench_n1.err; echo "bench rc=$?"
timeout ${NCU_TIMEOUT:-600} ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-2000} --csv --log-file gpurun_out/${p}_launches.csv \
  python bench.py --steps ${NCU_STEPS:-2} --warmup ${NCU_WARMUP:-3} --no-cpu-baseline --no-gpu-reference --no-other-configs > gpurun_out/${p}_launches_bench.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/${p}_pytest.log; tail -1 gpurun_out/${p}_smoke.log
python - <<PY
import json
d = json.load(open("gpurun_out/${p}_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v["ms"], 3) for k, v in d["roofline"]["families"].items()})
PY
