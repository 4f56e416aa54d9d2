#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r1e_pytest.log
python bench.py > gpurun_out/r1e_bench_c2.json 2> gpurun_out/r1e_bench.err
python bench.py --workload C4 --n 1250000 --no-cpu-baseline --steps 50 > gpurun_out/r1e_bench_c4shard.json 2>> gpurun_out/r1e_bench.err
python bench.py --q 1250 --no-cpu-baseline --steps 50 > gpurun_out/r1e_bench_c2q1250.json 2>> gpurun_out/r1e_bench.err
for f in gpurun_out/r1e_bench_*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=j["roofline"]; e=j.get("e2e") or {}
    print(j["config"]["workload"][:60], "| ms", round(j["ms_per_step"],4), "sync", round(j["config"]["ms_per_step_one_at_a_time"],4), "qps", int(j["value"]), "e2e", int(e.get("value",0)),
          "| roof", r["bound"], round(r["frac"],3), "kern_ms", round(r["kernel_ms"],4), {k:round(v,4) for k,v in r["other_kernels_ms"].items()}, j["gpu_launches"])
except Exception as ex: print("ERR", ex)
P
done
tail -n 5 gpurun_out/r1e_bench.err
