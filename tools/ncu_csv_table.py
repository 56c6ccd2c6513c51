#!/usr/bin/env python
"""Key metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` dump as a markdown table (the .ncu-rep files of
the row kernels are ~45 MB each with SASS and do not fit gpurun's 64 MiB return channel; the csv pages do)."""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp-inst"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue act %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n) for k, n in KEYS if k in hdr]
    print("| kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        vals = []
        for i, _ in cols:
            v = r[i]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append(v + (" " + units[i] if units[i] not in ("", "%") else ""))
        print("| `%s` | " % r[ki][:48] + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
