#!/usr/bin/env python3
"""Turn ncu outputs brought back in gpurun_out/ into the committed text summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python profiles/summarize.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_full.txt
"""
import collections
import csv
import subprocess
import sys

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    rows = rows[rows.index(hdr) + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] == "ns" else v / 1e3 if r[ui] == "us" else v
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms of kernel time (ncu: cold cache, serialised)")
    print(f"{'kernel':48s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:48]:48s} {v[0]:8d} {v[1]:10.3f} {100 * v[1] / tot:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {path}")
    for r in rows[2:]:
        print(f"\n== {r[ki][:110]}")
        for m in FULL:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:86s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
