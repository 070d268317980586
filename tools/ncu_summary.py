"""compact per-kernel summary of an `ncu --page raw --csv` export: for every kernel name the launch with the longest
duration, with the metrics the round notes quote.  usage: ncu_summary.py raw.csv [raw2.csv ...] > summary.txt"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

best = {}
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    ti = hdr.index("gpu__time_duration.sum")
    for r in rows[2:]:
        t = float(r[ti].replace(",", ""))
        if r[ki] not in best or t > best[r[ki]][0]:
            best[r[ki]] = (t, {w: (r[hdr.index(w)], units[hdr.index(w)]) for w in WANT if w in hdr})
for name, (t, m) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    print("----", name[:150])
    for w in WANT:
        if w in m:
            print(f"  {w:88s} {m[w][0]:>18s} {m[w][1]}")
