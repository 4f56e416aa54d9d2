#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
for w in "C2" "C1 --steps 50"; do
python bench.py --workload $w --no-cpu-baseline > gpurun_out/k3.json 2>> gpurun_out/k3.err
python - gpurun_out/k3.json <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline"]
print(j["config"]["workload"][:50], "| ms", round(j["ms_per_step"],4), "sync", round(j["config"]["ms_per_step_one_at_a_time"],4), "e2e", round(j["e2e"]["ms_per_step"],4), "kern", round(r["kernel_ms"],4), {k:round(v,4) for k,v in r["other_kernels_ms"].items()})
P
done
tail -n 3 gpurun_out/k3.err
