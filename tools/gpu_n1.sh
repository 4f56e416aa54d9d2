#!/bin/bash
# one-GPU evidence run: parity tests, bench lines of the named workloads, ncu launch lists and one
# `ncu --set full` capture of the three main kernels (each ncu pass only after its command ran clean)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 | tee gpurun_out/r1_pytest.log
python bench.py > gpurun_out/r1_bench_c2.json 2> gpurun_out/r1_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_c2_reference.json 2>> gpurun_out/r1_bench_c2.err
python bench.py --workload C1 --steps 50 > gpurun_out/r1_bench_c1.json 2>> gpurun_out/r1_bench.err
python bench.py --workload C3 --no-cpu-baseline --steps 10 > gpurun_out/r1_bench_c3.json 2>> gpurun_out/r1_bench.err
python bench.py --workload C4 --no-cpu-baseline --steps 20 > gpurun_out/r1_bench_c4.json 2>> gpurun_out/r1_bench.err
python bench.py --workload C4 --n 1250000 --no-cpu-baseline --steps 50 > gpurun_out/r1_bench_c4shard.json 2>> gpurun_out/r1_bench.err
python bench.py --workload C5 --n 1250000 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r1_bench_c5shard.json 2>> gpurun_out/r1_bench.err
NCUCMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph"
$NCUCMD > gpurun_out/r1_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_c2.csv $NCUCMD > gpurun_out/r1_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_c4shard.csv \
  python bench.py --workload C4 --n 1250000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/r1_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"simtopk_kernel|select_rescore" -s 9 -c 3 -f -o gpurun_out/r1_prof_c2 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/r1_ncu_full.log 2>&1
for f in gpurun_out/r1_bench_*.json; do echo "== $f"; python - "$f" <<'P'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    if j.get("impl") == "reference":
        print("reference arm:", int(j["value"]), j["unit"], j["cpu_baseline"]["cores"], "cores |", j["config"]["workload"][-60:])
    else:
        r=j["roofline"]; e=j.get("e2e") or {}; c=j.get("cpu_baseline") or {}
        print(j["config"]["workload"][:60], "| ms", round(j["ms_per_step"],4), "sync", round(j["config"]["ms_per_step_one_at_a_time"],4), "qps", int(j["value"]), "e2e", int(e.get("value",0)),
              "| roof", r["bound"], round(r["frac"],3), "kern_ms", round(r["kernel_ms"],4), {k:round(v,4) for k,v in r["other_kernels_ms"].items()}, j["config"]["path"].get("uncertified"), "cpu", c.get("value"))
except Exception as ex: print("ERR", ex)
P
done
tail -n 5 gpurun_out/r1_bench.err gpurun_out/r1_bench_c2.err
