#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (from `ncu --set full`) as a markdown table."""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp-inst"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue act %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
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
        print("| `%s` | " % r[ki][:60] + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
