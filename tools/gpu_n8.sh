#!/bin/bash
# N-GPU evidence run (default 8): sharded parity worker, the driver's default bench, the named multi-GPU configs
set -u
mkdir -p gpurun_out
N=${1:-8}
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
QUICK=${2:-full}
if [ "$QUICK" = full ]; then
run tests/mgpu_worker.py > gpurun_out/r1_mgpu_n$N.log 2>&1; echo "worker rc=$?"; grep -E "MGPU_OK|Error|error|assert|warn" gpurun_out/r1_mgpu_n$N.log | head -10
fi
b() { name=$1; shift; run bench.py --gpus $N --no-e2e "$@" > gpurun_out/r1_n${N}_$name.json 2> gpurun_out/r1_n${N}_$name.err; echo "$name rc=$?"; }
run bench.py --gpus $N > gpurun_out/r1_n${N}_c2_default.json 2> gpurun_out/r1_n${N}_c2_default.err; echo "default rc=$?"
b c4_peer --workload C4 --steps 50
if [ "$QUICK" = full ]; then
b c2_nccl --exchange nccl
b c4_nccl --workload C4 --steps 50 --exchange nccl
b c3x10_peer --workload C3 --gallery-rows 10000000 --steps 10
b c3_default --workload C3 --steps 20
b c5_peer --workload C5 --steps 5 --warmup 3
fi
for f in gpurun_out/r1_n${N}_*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    r=j["roofline"] or {}
    print(sys.argv[1].split("/")[-1], "| ms", round(j["ms_per_step"],4), "sync", round(j["config"].get("ms_per_step_one_at_a_time",0),4), "by_rank", r.get("kernel_ms_by_rank"), "qps", int(j["value"]), "| kern_ms", round(r.get("kernel_ms",0),4), "frac", round(r.get("frac",0),3), {k:round(v,4) for k,v in r.get("other_kernels_ms",{}).items()}, j["config"]["path"].get("exchange"), j["config"]["path"].get("uncertified"), (j.get("e2e") or {}).get("value"))
except Exception as ex: print(sys.argv[1], "ERR", ex)
P
done
for f in gpurun_out/r1_n${N}_*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|NCCL version\|^$" $f | tail -n 4; done
