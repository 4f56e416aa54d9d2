"""Summarise an `ncu --set full` report into the JSON kept under profiles/:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_summary.json
(reads the report with `ncu -i ... --page raw --csv`; values keep ncu's units in a sibling key)."""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.max.per_second", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["Kernel Name"][:60], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"),
              d.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
