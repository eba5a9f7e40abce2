#!/usr/bin/env python
"""Condenses one GPU-box visit (tools/gpu_round.sh <tag>) from gpurun_out/ into tracked files under profiles/.

    python tools/summarize_profiles.py <tag>

Writes profiles/<tag>_launches.csv (the ncu launch list), profiles/<tag>_launch_summary.txt (per-kernel
totals and shares), profiles/<tag>_ncu_full.txt (selected metrics of the `ncu --set full` capture) and copies
the bench JSON lines / logs of the same visit.
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

KEYS = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__inst_executed_pipe_fp64.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "sm__inst_executed_pipe_adu.avg.pct", "sm__inst_executed_pipe_cbu.avg.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg",
        "smsp__average_warps_issue_stalled", "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum",
        "gcc__cache_requests_type_instruction.sum", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")


def launch_list(tag, lines):
    src = os.path.join(G, tag + "_launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, ni = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    with open(os.path.join(P, tag + "_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "stream", "block", "grid", "gpu__time_duration.sum[ns]"])
        for r in rows[1:]:
            if r[ni] != "gpu__time_duration.sum":
                continue
            name = r[ki].split("(")[0].replace("void ", "")
            v = float(r[mi].replace(",", ""))
            w.writerow([r[0], name, r[hdr.index("Stream")], r[hdr.index("Block Size")], r[hdr.index("Grid Size")],
                        int(v)])
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values())
    probe = sum(a[1] for k, a in agg.items() if "fp64_probe" in k)
    vit = sum(a[1] for k, a in agg.items() if "viterbi" in k)
    lines.append("ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` over "
                 "`bench.py --steps 2 --warmup 1 --loci 20000` (serialised, cold cache: shares, not absolutes)" % tag)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-50s launches=%3d total_ms=%10.3f share=%.4f" % (k[-50:], n, t / 1e6, t / tot))
    lines.append("Viterbi kernels (all row classes, both modes): share of all launches %.4f; share of the step "
                 "(roofline probe excluded) %.4f" % (vit / tot, vit / (tot - probe)))
    open(os.path.join(P, tag + "_launch_summary.txt"), "w").write("\n".join(lines) + "\n")


def full_capture(tag, which="full", dest="ncu_full"):
    """Selected metrics of a full capture: from the report itself, or from its raw page exported on the box
    (<tag>_<which>_raw.csv) when the report was too large to bring back."""
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, which))
    raw_csv = os.path.join(G, "%s_%s_raw.csv" % (tag, which))
    if os.path.exists(raw_csv) and os.path.getsize(raw_csv) > 0:
        raw = open(raw_csv).read()
    elif os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["`ncu --set full --clock-control none --import-source on` (%s, %s), selected metrics" % (tag, which)]
    for r in rows[2:]:
        out.append("")
        out.append("KERNEL %s grid=%s block=%s" % (r[4].split("(")[0], r[8], r[7]))
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS):
                try:
                    if float(v.replace(",", "")) != 0:
                        out.append("  %-95s %s %s" % (h, v, u))
                except ValueError:
                    pass
    open(os.path.join(P, "%s_%s.txt" % (tag, dest)), "w").write("\n".join(out) + "\n")


def main():
    tag = sys.argv[1]
    os.makedirs(P, exist_ok=True)
    lines = []
    launch_list(tag, lines)
    full_capture(tag)
    full_capture(tag, "stutter", "ncu_stutter")   # kernel 2 (config 5)
    full_capture(tag, "c4stream", "ncu_c4_stream")  # full-matrix stream kernel on config 4
    for extra in ("memcheck.log", "racecheck.log", "dropin_timing.txt", "latency.txt"):
        src = os.path.join(G, "%s_%s" % (tag, extra))
        if os.path.exists(src) and os.path.getsize(src):
            if extra == "racecheck.log":  # keep the verdict lines, not the progress dots
                keep = [l for l in open(src, errors="ignore").read().splitlines()
                        if "RACECHECK SUMMARY" in l or "passed" in l or "rc=" in l or l.startswith("real")]
                hz = [l for l in open(src, errors="ignore").read().splitlines() if "Race reported" in l]
                open(os.path.join(P, "%s_%s" % (tag, extra)), "w").write("\n".join(sorted(set(hz))[:20] + keep) + "\n")
            else:
                shutil.copy(src, os.path.join(P, "%s_%s" % (tag, extra)))
    for fn in sorted(os.listdir(G)):
        if fn.startswith(tag + "_") and fn.endswith((".json", "probe.log", "pytest.log")):
            s = os.path.join(G, fn)
            if os.path.getsize(s):
                shutil.copy(s, os.path.join(P, fn))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
