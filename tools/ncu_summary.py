#!/usr/bin/env python
"""Dev helper: key metrics of every launch in an ncu report (pipes, issue, stalls, DRAM bytes).

    python tools/ncu_summary.py prof.ncu-rep
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["math_pipe_throttle", "wait", "not_selected", "long_scoreboard", "short_scoreboard", "dispatch_stall", "no_instruction",
               "branch_resolving", "lg_throttle", "mio_throttle", "barrier", "membar", "sleeping"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]][:60]
        print(f"== {name}")
        for k in KEYS:
            if k in ix:
                print(f"   {k:75s} {r[ix[k]]}")
        st = []
        for s in STALL_NAMES:
            k = STALLS % s
            if k in ix:
                st.append((float(r[ix[k]] or 0), s))
        print("   stalls per issue: " + ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    main()
