#!/bin/bash
# N-GPU evidence run: sharded parity worker, the driver's default bench (headline + also), optional extras
set -u
mkdir -p gpurun_out
N=${1:-2}
T=${2:-r2n}
MODE=${3:-full}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
if [ "$MODE" = full ] || [ "$MODE" = test ]; then
run tests/mgpu_worker.py > gpurun_out/${T}_mgpu_n$N.log 2>&1; echo "worker rc=$?"; grep -E "MGPU_OK|Error|error|assert|warn" gpurun_out/${T}_mgpu_n$N.log | head -10; tail -n 3 gpurun_out/${T}_mgpu_n$N.log
fi
if [ "$MODE" = full ] || [ "$MODE" = bench ]; then
run bench.py --gpus $N > gpurun_out/${T}_n${N}_default.json 2> gpurun_out/${T}_n${N}_default.err; echo "default rc=$?"
run bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_n${N}_reference.json 2>> gpurun_out/${T}_n${N}_default.err; echo "reference rc=$?"
fi
if [ "$MODE" = extra ]; then
run bench.py --gpus $N --workload C5 --also none --no-e2e --steps 4 --warmup 2 > gpurun_out/${T}_n${N}_c5.json 2> gpurun_out/${T}_n${N}_c5.err; echo "c5 rc=$?"
run bench.py --gpus $N --workload C3 --gallery-rows 10000000 --also none --no-e2e --steps 8 --warmup 3 > gpurun_out/${T}_n${N}_c3x10.json 2> gpurun_out/${T}_n${N}_c3x10.err; echo "c3x10 rc=$?"
fi
for f in gpurun_out/${T}_n${N}_*.json; do python - "$f" <<'P'
import json,sys
def show(j, pre=""):
    r=j.get("roofline") or {}; e=j.get("e2e") or {}
    print(pre, j["config"]["workload"][:90], "| ms", round(j["ms_per_step"],4), "sync", round(j["config"]["ms_per_step_one_at_a_time"],4), "qps", int(j["value"]), "e2e", int(e.get("value",0)),
          "| roof", r.get("bound"), round(r.get("frac",0),3), "kern_ms", round(r.get("kernel_ms",0),4), "by_rank", r.get("kernel_ms_by_rank"), {k:round(v,4) for k,v in (r.get("other_kernels_ms") or {}).items()}, "unc", j["config"]["path"].get("uncertified"), j["config"]["path"].get("exchange"), "probe", (j.get("probe") or {}).get("rank1"), "kps", j["config"].get("kernels_per_step"))
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    if j.get("impl") == "reference":
        print("reference arm:", int(j["value"]), j["unit"], j["cpu_baseline"]["cores"], "cores |", j["config"]["workload"][:100], j["config"]["host"].get("blas"))
    else:
        show(j)
        for a in j.get("also") or []:
            if "value" in a: show(a, "   also[%s]" % a["label"]); print("      ", {k:a[k] for k in a if k.startswith("value_") or k.startswith("eff") or k.startswith("n1_v")})
            else: print("   also", a)
except Exception as ex: print(sys.argv[1], "ERR", ex)
P
done
for f in gpurun_out/${T}_n${N}_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version\|^$" $f | tail -n 6; done
