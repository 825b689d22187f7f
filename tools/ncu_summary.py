#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("kernel:", name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-82s %s %s" % (k, r[i], units[i]))
    allsrc = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    starts = [i for i, r in enumerate(allsrc) if r and r[0] == "Kernel Name"] + [len(allsrc)]
    for si in range(len(starts) - 1):
        src = allsrc[starts[si]:starts[si + 1]]
        if len(src) <= 2:
            continue
        print("source page of:", src[0][1][:60])
        h = src[1]
        data = [r for r in src[2:] if len(r) == len(h)]
        iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        i0, i1 = h.index("stall_barrier"), h.index("stall_wait")
        names = h[i0:i1 + 1]
        tot = sum(int(r[iS]) for r in data)
        print("warp-state samples: %d; warp instructions executed: %d; top instructions (SASS, samples, dominant stall):"
              % (tot, sum(int(r[iEx]) for r in data)))
        for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:14]):
            r = data[i]
            st = sorted(((int(r[i0 + k]), names[k]) for k in range(len(names))), reverse=True)
            print("  %5d  %-56s %8s (%4.1f%%)  %s" % (i, r[iSrc].strip()[:56], r[iS], 100.0 * int(r[iS]) / max(tot, 1), st[0][1]))


if __name__ == "__main__":
    main()
