#!/usr/bin/env python
"""Condenses `ncu --page raw --csv` exports (gpurun_out/<tag>_<name>_raw.csv) into a tracked text summary:

    python tools/ncu_summary.py <tag> name1 name2 ...   ->   profiles/<tag>_ncu_<name>.txt
"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg")


def main():
    tag, names = sys.argv[1], sys.argv[2:]
    for name in names:
        src = os.path.join(ROOT, "gpurun_out", "%s_%s_raw.csv" % (tag, name))
        rows = list(csv.reader(open(src)))
        hdr, units = rows[0], rows[1]
        lines = ["ncu --set full --clock-control none (one launch; numbers taken under the profiler: shares and ratios, not "
                 "bench values); source gpurun_out/%s_%s.ncu-rep" % (tag, name)]
        for r in rows[2:]:
            lines.append("kernel: " + r[hdr.index("Kernel Name")])
            for i, h in enumerate(hdr):
                if h in KEYS:
                    lines.append("  %-70s %s %s" % (h, r[i], units[i]))
            stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
            for v, h in sorted(stalls, reverse=True)[:6]:
                lines.append("  stall %-64s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
        out = os.path.join(ROOT, "profiles", "%s_ncu_%s.txt" % (tag, name))
        open(out, "w").write("\n".join(lines) + "\n")
        print(out)


if __name__ == "__main__":
    main()
