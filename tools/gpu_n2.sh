#!/bin/bash
# two-GPU run: sharded parity worker, then bench lines for both partitions and both exchanges
set -u
mkdir -p gpurun_out
N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run tests/mgpu_worker.py > gpurun_out/r1_mgpu_n$N.log 2>&1; echo "worker rc=$?"; grep -E "MGPU_OK|Error|error|assert" gpurun_out/r1_mgpu_n$N.log | head -20
for ex in nccl peer; do
  run bench.py --gpus $N --steps 20 --no-e2e --exchange $ex > gpurun_out/r1_n${N}_c2_query_$ex.json 2> gpurun_out/r1_n${N}_$ex.err; echo "rc=$?"
  run bench.py --gpus $N --steps 20 --no-e2e --shard gallery --exchange $ex > gpurun_out/r1_n${N}_c2_gallery_$ex.json 2>> gpurun_out/r1_n${N}_$ex.err; echo "rc=$?"
  run bench.py --gpus $N --steps 30 --no-e2e --workload C4 --gallery-rows $((1250000 * N)) --exchange $ex > gpurun_out/r1_n${N}_c4_$ex.json 2>> gpurun_out/r1_n${N}_$ex.err; echo "rc=$?"
done
for f in gpurun_out/r1_n${N}_*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    r=j["roofline"] or {}
    print(sys.argv[1].split("/")[-1], "| ms", round(j["ms_per_step"],4), "qps", int(j["value"]), "| kern_ms", round(r.get("kernel_ms",0),4), {k:round(v,4) for k,v in r.get("other_kernels_ms",{}).items()}, j["config"]["path"].get("exchange"))
except Exception as ex: print(sys.argv[1], "ERR", ex)
P
done
tail -5 gpurun_out/r1_n${N}_*.err
