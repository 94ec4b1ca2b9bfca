#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export (one line per kernel launch) and, optionally, the hottest SASS lines
of an `ncu --page source --csv` export.  Usage: ncu_summary.py raw.csv [source.csv ...]"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        print(f"--- {name}")
        for w in WANT:
            if w in hdr:
                print(f"   {w:85s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")


def source(path, top=28):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    ia, isamp, iex, ith = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[hi + 1:] if len(r) > max(stall) and r[isamp].isdigit()]
    # the CSV repeats the listing per view: keep the first copy
    seen, uniq = set(), []
    for r in data:
        if r[0] in seen:
            break
        seen.add(r[0])
        uniq.append(r)
    data = uniq
    tot = sum(int(r[isamp]) for r in data)
    print(f"=== {path}: {len(data)} SASS lines, {tot} samples, {sum(int(r[iex]) for r in data)} warp-instructions")
    agg = {}
    for r in data:
        for i in stall:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("   stalls:", ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for idx, r in sorted(sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:top]):
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
        print(f"   {idx:4d} {r[ia].strip()[:58]:58s} samp {int(r[isamp]):6d} ({100 * int(r[isamp]) / max(tot, 1):4.1f}%) exec {r[iex]:>9s} thr {r[ith]:>3s}  {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")


if __name__ == "__main__":
    raw(sys.argv[1])
    for p in sys.argv[2:]:
        source(p)
