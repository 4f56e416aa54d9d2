#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r1b_pytest.log
python bench.py --workload C5 --n 1250000 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r1b_bench_c5shard.json 2> gpurun_out/r1b_bench.err
python - gpurun_out/r1b_bench_c5shard.json <<'P'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=j["roofline"]; print(j["config"]["workload"][:60], "| ms", round(j["ms_per_step"],4), "qps", int(j["value"]), r["bound"], round(r["frac"],3), r["kernel_ms"], r["other_kernels_ms"], j["config"]["path"])
P
tail -5 gpurun_out/r1b_bench.err
