#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of the first kernel in an .ncu-rep captured with
`--set full --import-source on` (binary built with -lineinfo).  Usage: ncu_source_lines.py rep [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fn, hdr, lines, nfn = None, None, [], 0
    for r in rows:
        if r and r[0] == "Function Name":
            nfn += 1
            if nfn > 1:
                break
            fn = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            lines.append(r)
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in lines) or 1
    ts = sum(int(r[isamp]) for r in lines) or 1
    print(fn)
    print("total inst %d samples %d" % (tot, ts))
    for r in sorted(lines, key=lambda r: -int(r[ie]))[:top]:
        print("%9d %5.1f%% samp %5.1f%%  L%s: %s" % (int(r[ie]), 100 * int(r[ie]) / tot, 100 * int(r[isamp]) / ts, r[0],
                                                  r[1].strip()[:110]))


if __name__ == "__main__":
    main()
