"""Condense `ncu -i <rep> --page raw --csv` of tools/ncu_kernels.py into the few columns profiles/ keeps: one row per
hot-kernel launch (the conv / trunk / GRU kernels only, in launch order, labelled with the layer they stand for).
    python tools/ncu_sections_csv.py gpurun_out/r2c_kernels_ncu_raw.csv profiles/r2c_kernels_ncu_sections.csv"""
import csv
import sys

COLS = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
# launch order of tools/ncu_kernels.py: every case runs 1 warm-up (inside exe.run of `timed`, reps = 2) -> 2 launches
LABELS = ["L2 conv1 C128 k5 prelu", "L3 conv1 C256 k5 prelu", "dec.3.up 128->64 x4 +skip", "enc.1 trunk C64",
          "dec.3 trunk C64 +sc +up tail", "enc.0 trunk C32 +down tail", "dec.4 trunk C32 +sc +out tail",
          "bottleneck BiGRU H=256 T=801"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, body = rows[0], rows[1], rows[2:]
idx = [hdr.index(c) for c in COLS if c in hdr]
hot = [r for r in body if any(k in r[hdr.index("Kernel Name")] for k in ("conv1d_tc_kernel", "trunk_kernel", "gru_cluster"))]
per = max(1, len(hot) // len(LABELS))
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["layer (B=32, 8 s clips)"] + [hdr[i] for i in idx])
    w.writerow([""] + [units[i] for i in idx])
    for n, r in enumerate(hot):
        if n % per != per - 1:
            continue                      # keep the last (warm) launch of each case
        w.writerow([LABELS[min(n // per, len(LABELS) - 1)]] + [r[i] for i in idx])
print(f"{len(hot)} hot launches -> {len(hot) // per} rows")
