#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8
python bench.py --no-cpu-baseline > gpurun_out/k3.json 2>> gpurun_out/k3.err
python - gpurun_out/k3.json <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline"]
print(j["config"]["workload"][:50], "| ms", round(j["ms_per_step"],4), "e2e", j["e2e"]["ms_per_step"], "kern", round(r["kernel_ms"],4), {k:round(v,4) for k,v in r["other_kernels_ms"].items()})
P
tail -n 3 gpurun_out/k3.err
