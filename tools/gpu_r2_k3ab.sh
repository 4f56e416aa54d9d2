#!/bin/bash
# K3 A/B: list-walk / staging variants (HCIR_K3_VARIANT bit 0 = per-list walk, bit 1 = per-key atomics) and CTA widths
set -u
mkdir -p gpurun_out
T=${1:-r2b}
for v in 0 1 2 3; do
  HCIR_K3_VARIANT=$v timeout 200 python bench.py --also none --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/${T}_c3_v$v.json 2>> gpurun_out/${T}.err
  HCIR_K3_VARIANT=$v timeout 200 python bench.py --workload C2 --also none --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/${T}_c2_v$v.json 2>> gpurun_out/${T}.err
done
for w in 2 3; do
  timeout 200 python bench.py --q 512 --k3-width $w --also none --no-cpu-baseline --no-e2e --steps 30 > gpurun_out/${T}_c3q512_w$w.json 2>> gpurun_out/${T}.err
done
timeout 200 python bench.py --workload C2 --k3-width 2 --also none --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/${T}_c2_w2.json 2>> gpurun_out/${T}.err
for f in gpurun_out/${T}_*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    r=j["roofline"]
    print(sys.argv[1].split("/")[-1], "| ms", round(j["ms_per_step"],4), "main", round(r["kernel_ms"],4), {k:round(v,4) for k,v in r["other_kernels_ms"].items()})
except Exception as ex: print(sys.argv[1], "ERR", ex)
P
done
tail -n 5 gpurun_out/${T}.err
