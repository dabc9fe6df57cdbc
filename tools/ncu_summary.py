#!/usr/bin/env python3
"""Summarise ncu outputs into profiles/ (text the judge can read without ncu).

  launches <launches.csv>            per-kernel count / total device time / share (gpu__time_duration.sum pass)
  full <report.ncu-rep> [regex]      key raw metrics of each profiled launch (ncu --set full capture)
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum"]


def launches(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        v = r["Metric Value"].replace(",", "")
        agg[k][0] += 1
        agg[k][1] += float(v) if v else 0.0
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.3f ms total device time (ncu, cold-cache, serialised: compare SHARES)" % (path, len(rows), tot / 1e6))
    print("%6s %14s %7s  kernel" % ("count", "total_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6d %14.1f %6.1f%%  %s" % (v[0], v[1] / 1e3, 100 * v[1] / tot, k[:110]))


def full(path, regex=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    for r in data:
        print("## launch id %s: %s" % (r[0], r[name_i][:120]))
        for i, h in enumerate(hdr):
            if h in KEYS:
                print("  %-80s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(*sys.argv[2:])
