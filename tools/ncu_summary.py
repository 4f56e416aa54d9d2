"""Summarise an `ncu --set full` report into the JSON kept under profiles/:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_ncu_full_<workload>.json
Reads the report with `ncu -i ... --page raw --csv` (one record per profiled kernel launch; values keep
ncu's units) and `--page source --csv` (warp-stall sampling summed over the kernel's SASS: the share of
every stall reason, which says WHY a pipe idles -- e.g. whether the tensor pipe of the main pass waits on
MMA issue, on the shared-memory operands or on the accumulator-empty barrier)."""
import collections
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]


def _norm(kernel_name: str) -> str:
    """'void hcir::simtopk_kernel<(int)0, (int)1>(...)' and 'void simtopk_kernel<0, 1>(...)' -> 'simtopk_kernel<0,1>'"""
    s = kernel_name.split("(CU")[0].split("(const")[0]
    for junk in ("void ", "hcir::", "(int)", " "):
        s = s.replace(junk, "")
    return s.split(">(")[0] + (">" if "<" in s and not s.split(">(")[0].endswith(">") else "")


def stall_shares(rep):
    """kernel name -> {stall reason: percent of all warp-stall samples}, from the SASS source page."""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    out, name, hdr, tot = {}, None, None, None
    for r in csv.reader(txt.splitlines()):
        if len(r) >= 2 and r[0] == "Kernel Name":
            if name is not None and tot:
                out.setdefault(name, tot)
            name, hdr, tot = r[1], None, collections.Counter()
        elif r and r[0] == "Address":
            hdr = r
        elif hdr is not None and len(r) == len(hdr):
            for k, v in zip(hdr, r):
                if k.startswith("stall_") and "Not Issued" not in k and v:
                    tot[k] += int(v)
    if name is not None and tot:
        out.setdefault(name, tot)
    res = {}
    for k, c in out.items():
        s = sum(c.values()) or 1
        res[k] = {r: round(100.0 * v / s, 1) for r, v in c.most_common(8)}
    return res


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    stalls = stall_shares(rep)
    res = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {"Kernel Name": name}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        for full, sh in stalls.items():
            if _norm(full) == _norm(name):
                d["warp_stall_share_pct"] = sh
                break
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["Kernel Name"][:60], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"),
              d.get("dram__bytes_write.sum"), d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
              d.get("warp_stall_share_pct"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
