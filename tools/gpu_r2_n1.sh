#!/bin/bash
# round-2 one-GPU evidence run: parity tests, the driver's default bench line (+ reference arm), ncu launch lists
# and `ncu --set full` captures of the headline (C3) and of the streaming-regime shard (C4, 1.25M rows)
set -u
mkdir -p gpurun_out
T=${1:-r2a}
STAGE=${2:-all}
if [ "$STAGE" = all ] || [ "$STAGE" = test ]; then
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -n 25 | tee gpurun_out/${T}_pytest_round2.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -n 25 | tee gpurun_out/${T}_pytest.log
fi
if [ "$STAGE" = all ] || [ "$STAGE" = bench ]; then
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --workload C4 --gallery-rows 1250000 --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/${T}_bench_c4shard.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --workload C4 --gallery-rows 1250000 --no-cpu-baseline --no-e2e --steps 50 --k3-width 2 > gpurun_out/${T}_bench_c4shard_k3w256.json 2>> gpurun_out/${T}_bench.err
HCIR_MAIN_FLAGS=16 timeout 300 python bench.py --also none --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_c3_pairs.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --workload C5 --gallery-rows 1250000 --no-cpu-baseline --no-e2e --steps 5 --warmup 3 > gpurun_out/${T}_bench_c5shard.json 2>> gpurun_out/${T}_bench.err
fi
if [ "$STAGE" = all ] || [ "$STAGE" = ncu ]; then
C3CMD="python bench.py --also none --no-cpu-baseline --no-e2e --steps 2 --warmup 3"
C4CMD="python bench.py --workload C4 --gallery-rows 1250000 --also none --no-cpu-baseline --no-e2e --steps 2 --warmup 3"
$C3CMD > gpurun_out/${T}_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c3.csv $C3CMD > gpurun_out/${T}_ncu_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c4shard.csv $C4CMD > gpurun_out/${T}_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"simtopk_kernel|select_rescore" -s 12 -c 4 -f -o gpurun_out/${T}_prof_c3 $C3CMD > gpurun_out/${T}_ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"simtopk_kernel|select_rescore|l2norm|threshold" -s 20 -c 6 -f -o gpurun_out/${T}_prof_c4shard $C4CMD > gpurun_out/${T}_ncu_full_c4.log 2>&1
fi
for f in gpurun_out/${T}_bench*.json; do echo "== $f"; python - "$f" <<'P'
import json,sys
def show(j, pre=""):
    r=j.get("roofline") or {}; e=j.get("e2e") or {}; c=j.get("cpu_baseline") or {}
    print(pre, j["config"]["workload"][:70], "| ms", round(j["ms_per_step"],4), "sync", round(j["config"]["ms_per_step_one_at_a_time"],4), "qps", int(j["value"]), "e2e", int(e.get("value",0)),
          "| roof", r.get("bound"), round(r.get("frac",0),3), "kern_ms", round(r.get("kernel_ms",0),4), {k:round(v,4) for k,v in (r.get("other_kernels_ms") or {}).items()}, "unc", j["config"]["path"].get("uncertified"), "probe", j.get("probe"), "cpu", c.get("value"))
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    if j.get("impl") == "reference":
        print("reference arm:", int(j["value"]), j["unit"], j["cpu_baseline"]["cores"], "cores |", j["config"]["workload"][:100])
    else:
        show(j)
        for a in j.get("also") or []:
            if "value" in a: show(a, "   also[%s]" % a["label"]); print("      ", {k:a[k] for k in a if k.startswith("value_") or k.startswith("eff")})
            else: print("   also", a)
        if j.get("cpu_baseline"): print("   cpu:", {k:(v if not isinstance(v,dict) else v.get("value")) for k,v in j["cpu_baseline"].items() if k not in ("host","sample")})
except Exception as ex: print("ERR", ex)
P
done
tail -n 12 gpurun_out/${T}_bench.err 2>/dev/null
