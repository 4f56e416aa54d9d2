#!/bin/bash
# A/B of K3 build options on the GPU box (rebuilds the library there through HCIR_NVCC_EXTRA): register budget
# (-DHCIR_K3_THREADS_PER_SM) and dot-product load grouping (-DHCIR_DOT_GROUP); results: profiles/README.md r2d
set -u
mkdir -p gpurun_out
T=${1:-r2d}
run() { tag=$1; shift
  for wl in "--workload C2" "" "--q 512" "--workload C4 --gallery-rows 1250000 --steps 50" "--workload C5 --gallery-rows 1250000 --steps 4 --warmup 2"; do
    n=$(echo "$wl" | tr -d ' -' | cut -c1-24); [ -z "$n" ] && n=C3
    timeout 300 python bench.py $wl --also none --no-cpu-baseline --no-e2e > gpurun_out/${T}_${tag}_$n.json 2>> gpurun_out/${T}.err
  done; }
run base
HCIR_NVCC_EXTRA="-DHCIR_K3_THREADS_PER_SM=1280" run occ1280
HCIR_NVCC_EXTRA="-DHCIR_DOT_GROUP=2 -DHCIR_K3_THREADS_PER_SM=1280" run occ1280_g2
HCIR_NVCC_EXTRA="-DHCIR_DOT_GROUP=4" run g4
for f in gpurun_out/${T}_*.json; do python - "$f" <<'P'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    r=j["roofline"]
    print(sys.argv[1].split("/")[-1], "| ms", round(j["ms_per_step"],4), "sync", round(j["config"]["ms_per_step_one_at_a_time"],4), "main", round(r["kernel_ms"],4), {k:round(v,4) for k,v in r["other_kernels_ms"].items()}, j["config"]["path"].get("uncertified"))
except Exception as ex: print(sys.argv[1], "ERR", ex)
P
done
tail -n 5 gpurun_out/${T}.err
