"""Summarise an .ncu-rep (raw page) into the metrics DESIGN.md / profiles/ quote.  python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_register_spilling", "smsp__pcsamp_warps_issue_stalled_barrier", "l1tex__throughput.avg.pct_of_peak_sustained_active",
]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"### {d['Kernel Name'][:60]} (id {d['ID']})")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
        stalls = {k.split("issue_stalled_")[1].split("_per")[0]: float(d[k]) for k in hdr
                  if k.startswith("smsp__average_warp") and "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and "not_issued" not in k and d[k]}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:10]
        print("| warp stall cycles per issued instruction | " + ", ".join(f"{k} {v:.2f}" for k, v in top) + " | |")
        print()


main()
